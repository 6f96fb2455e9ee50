"""Outer training / testing loops and optimizer construction around the hot path.

Mirrors of the reference's `train` (n_best_asr_bert.py:391-439), `test` (:442-473) and the optimizer block of its
`__main__` (:522-569), with the same arguments, log lines, per-epoch prediction files and best-on-valid-F1 model
selection — plus what the reference lacks: a full resumable checkpoint next to `model.pt` (checkpoint.py) and rank-0-only
file output under data parallelism. Logging plumbing (`make_logger`, the observability CSVs of `observability_lens`) is
outside the hot path: a standard `logging` logger is used and the CSV export is left to the caller, who
receives the same EpochInfo objects the reference would hand to it.
"""
import logging
import os
import time
from datetime import timedelta

import torch

from . import checkpoint as ckpt
from .epoch import eval_epoch, train_epoch
from .optim import Adam, AdamW, BertAdam, get_linear_schedule_with_warmup

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")          # n_best_asr_bert.py:540


def grouped_parameters(model, lr, bert_lr):
    """One param group PER TENSOR (n_best_asr_bert.py:535-550): weight_decay 0.01 unless the name contains a no-decay
    pattern; lr = bert_lr for `bert_encoder.*`, lr otherwise."""
    groups = []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        groups.append(dict(params=p, weight_decay=0.0 if any(nd in n for nd in NO_DECAY) else 0.01,
                           lr=bert_lr if "bert_encoder" in n else lr))
    return groups


def build_optimizer(opt, model, n_train):
    """The `--optim_choice` block (n_best_asr_bert.py:522,553-569). Sets opt.n_accum_steps (4 iff --n_layers 12),
    opt.optimizer and, for adamw, opt.scheduler; returns the number of optimisation steps the schedules are built for,
    `(n_train // batchSize + 1) * max_epoch` (:556)."""
    if getattr(opt, "n_accum_steps", None) is None:
        opt.n_accum_steps = 4 if getattr(opt, "n_layers", 6) == 12 else 1
    choice = opt.optim_choice.lower()
    steps = (n_train // opt.batchSize + 1) * opt.max_epoch
    if choice == "adam":
        params = [p for p in model.parameters() if p.requires_grad]
        opt.optimizer = Adam(params, lr=opt.lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=getattr(opt, "l2", 0.0))
    elif choice == "bertadam":
        opt.optimizer = BertAdam(grouped_parameters(model, opt.lr, opt.bert_lr), lr=opt.lr, warmup=opt.warmup_proportion,
                                 t_total=steps)
    elif choice == "adamw":
        opt.optimizer = AdamW(grouped_parameters(model, opt.lr, opt.bert_lr), lr=opt.lr, correct_bias=False)
        opt.scheduler = get_linear_schedule_with_warmup(opt.optimizer, num_warmup_steps=int(opt.warmup_proportion * steps),
                                                        num_training_steps=steps)
    else:
        raise ValueError("optim_choice must be adam | adamw | bertadam, got %r" % opt.optim_choice)
    return steps


def _logger(path, name):
    lg = logging.getLogger("nbest_b200.%s.%s" % (name, path))
    lg.setLevel(logging.INFO)
    if not lg.handlers:
        h = logging.FileHandler(path)
        h.setFormatter(logging.Formatter("%(asctime)s %(message)s"))
        lg.addHandler(h)
    return lg


def _rank():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class _Null:
    def write(self, *_):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _open(path, rank):
    return open(path, "w") if rank == 0 else _Null()


def train(model, train_dataloader, valid_dataloader, test_dataloader, opt, memory, resume=True):
    """Reference `train` (n_best_asr_bert.py:391-439). Every epoch: train_epoch, eval_epoch on valid and test with the
    prediction dumps `valid.iter%d[.err]` / `test.iter%d[.err]`, and `model.save_model(exp_dir/model.pt)` whenever the
    valid F1 improves (:427-436). Additionally `exp_dir/last.ckpt` holds the full state after every epoch; with
    resume=True an existing one is continued from (epoch counter and `best` included). Returns the `best` dict."""
    rank = _rank()
    os.makedirs(opt.exp_dir, exist_ok=True)
    logger = _logger(os.path.join(opt.exp_dir, "log.train"), "train")
    t0 = time.time()
    logger.info("Training starts at %s" % time.asctime(time.localtime(time.time())))
    best = {"epoch": 0, "vf": 0.0, "tef": 0.0}
    first = 0
    last = os.path.join(opt.exp_dir, "last.ckpt")
    if resume and os.path.exists(last):
        cur = ckpt.load_checkpoint(last, model, opt.optimizer)
        first = int(cur.get("epoch", -1)) + 1
        best = dict(cur.get("best", best))
        if getattr(opt, "scheduler", None) is not None and "scheduler" in cur:
            opt.scheduler.load_state_dict(cur["scheduler"])
        logger.info("Resumed from %s at epoch %d" % (last, first))
    for i in range(first, opt.max_epoch):
        start = time.time()
        train_loss, (trp, trr, trf), tr_acc = train_epoch(model, train_dataloader, opt, memory)
        logger.info("[Train]\tEpoch: %02d\tTime: %.2f\tLoss: %.2f\t(p/r/f): (%.2f/%.2f/%.2f)\tAcc: %.2f" %
                    (i, time.time() - start, train_loss, trp, trr, trf, tr_acc))
        res = {}
        for name, loader in (("valid", valid_dataloader), ("test", test_dataloader)):
            with _open(os.path.join(opt.exp_dir, "%s.iter%d" % (name, i)), rank) as fp, \
                    _open(os.path.join(opt.exp_dir, "%s.iter%d.err" % (name, i)), rank) as efp:
                start = time.time()
                out = eval_epoch(model, loader, opt, memory, fp, efp)
                loss, (p, r, f), acc = out[0], out[1], out[2]
                logger.info("[%s]\tEpoch: %02d\tTime: %.2f\tLoss: %.2f\t(p/r/f): (%.2f/%.2f/%.2f)\tAcc: %.2f" %
                            (name.capitalize(), i, time.time() - start, loss, p, r, f, acc))
                res[name] = (f, acc, out[-1])
        vf, v_acc, _ = res["valid"]
        tef, te_acc, _ = res["test"]
        if vf > best["vf"]:
            best.update(epoch=i, vf=vf, tef=tef, v_acc=v_acc, te_acc=te_acc)
            if rank == 0:
                model.save_model(os.path.join(opt.exp_dir, "model.pt"))
            logger.info("NEW BEST:\tEpoch: %02d\tvalid F1/Acc: %.2f/%.2f\ttest F1/Acc: %.2f/%.2f" % (i, vf, v_acc, tef, te_acc))
        if rank == 0:
            cur = dict(epoch=i, best=best)
            if getattr(opt, "scheduler", None) is not None:
                cur["scheduler"] = opt.scheduler.state_dict()
            ckpt.save_checkpoint(last, model, opt.optimizer, cursor=cur)
    logger.info("Done training. Elapsed time: %s" % timedelta(seconds=time.time() - t0))
    if "v_acc" in best:
        logger.info("BEST RESULT:\tEpoch: %02d\tBest valid F1/Acc: %.2f/%.2f\ttest F1/Acc: %.2f/%.2f" % (
            best["epoch"], best["vf"], best["v_acc"], best["tef"], best["te_acc"]))
    return best


def test(model, train_dataloader, valid_dataloader, test_dataloader, opt, memory):
    """Reference `test` (n_best_asr_bert.py:442-473): eval_epoch over the three splits with `<split>.eval[.err]` dumps.
    The caller loads `exp_dir/model.pt` first, as the reference's __main__ does (:578). Returns {split: (loss, prf, acc)}."""
    rank = _rank()
    logger = _logger(os.path.join(opt.exp_dir, "log.test"), "test")
    t0 = time.time()
    logger.info("Testing starts at %s" % time.asctime(time.localtime(time.time())))
    out = {}
    for name, loader in (("train", train_dataloader), ("valid", valid_dataloader), ("test", test_dataloader)):
        with _open(os.path.join(opt.exp_dir, "%s.eval" % name), rank) as fp, \
                _open(os.path.join(opt.exp_dir, "%s.eval.err" % name), rank) as efp:
            start = time.time()
            r = eval_epoch(model, loader, opt, memory, fp, efp)
            loss, (p, rr, f), acc = r[0], r[1], r[2]
            logger.info("[%s]\tTime: %.2f\tLoss: %.2f\t(p/r/f): (%.2f/%.2f/%.2f)\tAcc: %.2f" %
                        (name.capitalize(), time.time() - start, loss, p, rr, f, acc))
            out[name] = (loss, (p, rr, f), acc)
    logger.info("Done testing. Elapsed time: %s" % timedelta(seconds=time.time() - t0))
    return out
