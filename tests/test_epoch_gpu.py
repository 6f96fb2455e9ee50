"""GPU side of SURVEY §8(f): device metrics (nbest_stc_metrics) against the reference's update_f1 counters, the epoch
loops (train_epoch / eval_epoch mirrors) against the oracle driven step by step, checkpoint -> resume continuity, and the
pinned prefetcher. Golden vectors: tests/golden/epoch_valid48.npz (made by the live reference, oracle/make_golden.py)."""
import json
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hier():
    from oracle import stc_oracle as O
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    t2b = {int(k): v for k, v in hj["top2bottom"].items()}
    return O.Hierarchy(t2b, hj["none_bottoms"]), hj, t2b


def test_device_metrics_equal_reference_counters():
    from nbest_b200 import epoch as E, ops
    from oracle import stc_oracle as O
    fx = np.load(os.path.join(GOLD, "epoch_valid48.npz"))
    meta = json.loads(str(fx["meta"]))
    hier, hj, t2b = _hier()
    idx2label = {int(k): v for k, v in hj["idx2label"].items()}
    top, flat = torch.from_numpy(fx["top"]), torch.from_numpy(fx["bottom"])
    bottoms, c = {}, 0
    for k in sorted(t2b):
        if len(t2b[k]) > 1:
            bottoms["lin_%d" % k] = flat[:, c:c + len(t2b[k])]
            c += len(t2b[k])
    decode = torch.from_numpy(np.asarray(O.decode(hier, top, bottoms)).astype(np.uint8)).cuda()
    labels = torch.from_numpy(fx["labels"]).cuda()
    m = E.EpochMetrics("cuda")
    m.update(decode[:20].contiguous(), labels[:20].contiguous())            # accumulates across calls / ragged batch sizes
    m.update(decode[20:].contiguous(), labels[20:].contiguous())
    assert m.counters.tolist() == fx["counts"].tolist() + [48]
    _, prf, acc = m.result()
    assert np.allclose(prf, fx["prf"]) and abs(acc - 100.0 * fx["counts"][3] / 48) < 1e-9
    mask = E.informative_mask(idx2label, meta["ontology"], hier.n_bottom)
    mf = E.EpochMetrics("cuda", mask)
    mf.update(decode, E.collate_labels(meta["label_lists"], meta["label2idx"], "cuda", ontology=meta["ontology"]))
    assert mf.counters.tolist()[:4] == fx["counts_filtered"].tolist()
    assert np.allclose(mf.result()[1], fx["prf_filtered"])


class _Tok:
    """Deterministic stand-in tokenizer (no vocabularies exist offline): one id per distinct word, 2 pieces for long words."""
    cls_token, sep_token, pad_token_id = "[CLS]", "[SEP]", 0

    def tokenize(self, w):
        return [w] if len(w) < 7 else [w[:4], "##" + w[4:]]

    def convert_tokens_to_ids(self, toks):
        special = {"[CLS]": 101, "[SEP]": 102, "[SYS]": 103, "[USR]": 104}
        return [special.get(t, 1000 + (hash_str(t) % 1800)) for t in toks]


def hash_str(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def _setup(layers=2, dropout=0.0):
    from oracle import stc_oracle as O
    from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC
    from nbest_b200.optim import BertAdam
    hier, hj, t2b = _hier()
    fx = np.load(os.path.join(GOLD, "epoch_valid48.npz"))
    meta = json.loads(str(fx["meta"]))
    cfg = O.EncoderConfig.bert_base(layers=layers, vocab_size=3000, max_position=512)
    params = O.init_params(cfg, hier, seed=71, style="perturbed")
    spec = EncoderSpec(kind="bert", vocab_size=cfg.vocab_size, layers=layers, max_position=cfg.max_position,
                       hidden_dropout=dropout, attn_dropout=dropout)
    model = TOD_ASR_Transformer_STC(spec=spec, top2bottom=t2b, dropout=dropout, device="cuda", none_bottoms=hj["none_bottoms"])
    model.load_state_dict(params)
    groups = [dict(params=p, lr=2e-4, weight_decay=0.0 if "bias" in n or "LayerNorm" in n else 0.01)
              for n, p in model.named_parameters()]
    optim = BertAdam(groups, lr=2e-4, warmup=0.1, t_total=40)
    memory = dict(idx2label={int(k): v for k, v in hj["idx2label"].items()}, label2idx=meta["label2idx"], top2bottom_dict=t2b)
    opt = Namespace(tokenizer=_Tok(), device="cuda", add_segment_ids=True, add_l2_loss=False, optimizer=optim,
                    pre_trained_model="bert", tod_pre_trained_model=None, without_system_act=False, ontology=None, testing=False)
    raw_in = [s.split(" ") for s in meta["raw_in"]]
    raw_trans = [s.split(" ") for s in meta["raw_trans"]]
    return model, optim, opt, memory, meta, raw_in, raw_trans, cfg, params, hier


def _batches(E, meta, raw_in, raw_trans, bs):
    out = []
    for s in range(0, len(raw_in), bs):
        ll = meta["label_lists"][s:s + bs]
        out.append((E.collate_labels(ll, meta["label2idx"]), raw_in[s:s + bs], raw_trans[s:s + bs], ll))
    return out


def test_eval_epoch_matches_oracle_forward_and_reference_metric_code():
    """eval_epoch mirror: mean loss = mean over batches of (sum of loss terms / B) like n_best_asr_bert.py:331-332; F1 and
    accuracy recomputed on the host from the dumped label strings with the reference's update_f1 restatement."""
    from nbest_b200 import epoch as E
    from nbest_b200.inputs import prepare_inputs_for_roberta
    from oracle import stc_oracle as O
    import io
    model, optim, opt, memory, meta, raw_in, raw_trans, cfg, params, hier = _setup()
    data = _batches(E, meta, raw_in, raw_trans, 16)
    fp = io.StringIO()
    mean_loss, (p, r, f), acc, eic = E.eval_epoch(model, data, opt, memory, fp=fp)
    cases = list(zip([x.split(" ") for x in eic.raw_inputs], eic.whole_pred_classes, eic.true_golds))
    assert len(cases) == 48 and fp.getvalue().count("\n") == 48
    # oracle, batch by batch
    losses, tp, fpn, fn, corr = [], 0, 0, 0, 0
    for labels, rin, rtr, ll in data:
        ids, seg, _ = prepare_inputs_for_roberta(rin, opt.tokenizer, opt, "cpu")
        tids, tseg, _ = prepare_inputs_for_roberta(rtr, opt.tokenizer, opt, "cpu")
        top, bottoms, final, a, t = O.model_forward(params, cfg, hier, ids, tids, seg, tseg)
        total, rec = O.total_loss(hier, top, bottoms, final, labels, a, t, False)
        losses.append(float(total) / len(rin))
        preds = E.decode_to_labels(np.asarray(O.decode(hier, top, bottoms)), memory["top2bottom_dict"], memory["idx2label"])
        for pr, go in zip(preds, ll):
            tp, fpn, fn = E.update_f1(pr, go, tp, fpn, fn)
            corr += int(set(pr) == set(go))
    assert abs(mean_loss - np.mean(losses)) / abs(np.mean(losses)) < 1e-2
    # a prediction can flip where a score sits within bf16 noise of 0.5 / of the runner-up; allow a couple of labels
    ours = [sum(x) for x in zip(*[E.update_f1(pr, go, 0, 0, 0) for _, pr, go in cases])]
    assert abs(ours[0] - tp) <= 2 and abs(ours[1] - fpn) <= 2 and abs(ours[2] - fn) <= 2
    assert np.allclose((p, r, f), E.compute_f1(*ours)) and abs(acc - 100.0 * sum(set(pr) == set(go) for _, pr, go in cases) / 48) < 1e-9


def test_train_epoch_checkpoint_and_resume_are_continuous(tmp_path):
    """Two epochs in one go == one epoch, save_checkpoint, fresh model + optimizer, load_checkpoint, second epoch: identical
    weights, moments and step counters (dropout off: the only nondeterminism left is the fp32 atomic order of wgrad)."""
    from nbest_b200 import epoch as E
    from nbest_b200.checkpoint import load_checkpoint, save_checkpoint
    model, optim, opt, memory, meta, raw_in, raw_trans, *_ = _setup()
    data = _batches(E, meta, raw_in, raw_trans, 12)
    l1, prf1, acc1 = E.train_epoch(model, data, opt, memory)
    assert np.isfinite(l1) and 0 <= acc1 <= 100
    path = save_checkpoint(str(tmp_path / "ck.pt"), model, optim, cursor=dict(epoch=1, best_f=prf1[2]))
    l2, _, _ = E.train_epoch(model, data, opt, memory)
    assert l2 < l1                                        # it learns something on 48 utterances
    ref_params, ref_m, ref_steps = model.flat.params.clone(), optim.flat.m.clone(), list(optim._steps)

    model_b, optim_b, opt_b, *_ = _setup()
    cursor = load_checkpoint(path, model_b, optim_b)
    assert cursor["epoch"] == 1 and optim_b._steps == [s - len(data) if s else 0 for s in ref_steps]
    l2b, _, _ = E.train_epoch(model_b, data, opt_b, memory)
    assert abs(l2b - l2) / abs(l2) < 1e-3 and optim_b._steps == ref_steps
    cos = lambda a, b: float((a.double() @ b.double()) / (a.double().norm() * b.double().norm()))
    assert cos(model_b.flat.params, ref_params) > 1 - 1e-6 and cos(optim_b.flat.m, ref_m) > 0.999    # fp32 atomic order of wgrad
    # weights-only part is the reference's save_model format: a bare state_dict loads too
    torch.save({k: v.cpu() for k, v in model.state_dict().items()}, str(tmp_path / "bare.pt"))
    model_c, *_ = _setup()
    assert load_checkpoint(str(tmp_path / "bare.pt"), model_c) == {}
    assert torch.equal(model_c.flat.params, model.flat.params)


def test_train_epoch_with_cuda_graphs_matches_the_eager_epoch():
    """opt.cuda_graphs: the epoch's optimizer steps replayed as CUDA graphs over filler-bucketed batches give the epoch the
    eager loop gives — mean loss, P / R / F counters, accuracy (dropout off; fillers and graph replay change rounding only),
    and the second epoch replays the graphs the first one captured (the fixture's four 12-utterance batches differ in
    token count and two of them hold sequences above 128 tokens: four shapes)."""
    from nbest_b200 import epoch as E
    model, optim, opt, memory, meta, raw_in, raw_trans, *_ = _setup()
    model_g, optim_g, opt_g, *_ = _setup()
    data = _batches(E, meta, raw_in, raw_trans, 12)
    opt_g.cuda_graphs, opt_g.graph_bucket, opt_g.graph_width = True, (4, 256), 256
    for ep in range(2):
        l, prf, acc = E.train_epoch(model, data, opt, memory)
        lg, prfg, accg = E.train_epoch(model_g, data, opt_g, memory)
        assert abs(l - lg) / abs(l) < 2e-3, (ep, l, lg)
        assert abs(acc - accg) <= 100.0 / 48 * 2 and max(abs(a - b) for a, b in zip(prf, prfg)) <= 5.0, (prf, prfg, acc, accg)
    gt = opt_g._nbest_graphed
    assert gt.capture_error is None and gt.eager_steps == 1 and gt.replays == 2 * len(data) - 1
    assert gt.captures <= len(data), gt.captures
    assert optim_g._steps == optim._steps
    cos = lambda a, b: float((a.double() @ b.double()) / (a.double().norm() * b.double().norm()))
    assert cos(model_g.flat.params, model.flat.params) > 1 - 2e-5


def test_prefetcher_feeds_the_trainer_from_pretokenized_data(tmp_path):
    from nbest_b200 import data as D, epoch as E
    from nbest_b200.trainer import DataParallelTrainer
    model, optim, opt, memory, meta, raw_in, raw_trans, *_ = _setup()
    out = D.pretokenize(raw_in, raw_trans, meta["label_lists"], opt.tokenizer, opt, meta["label2idx"], str(tmp_path / "pt"))
    ds = D.PretokenizedDataset(out)
    trainer = DataParallelTrainer(model, optim)
    model.train()
    metrics = E.EpochMetrics("cuda")
    n = 0
    for b in D.Prefetcher(ds, D.epoch_order(len(ds), 12, True, 3, 0), "cuda", depth=3):
        assert b["ids"].is_cuda and b["labels"].is_cuda
        ref = ds.batch(b["index"], pinned=False)
        assert torch.equal(b["ids"].cpu(), ref["ids"]) and torch.equal(b["trans_seg"].cpu(), ref["trans_seg"])
        losses = trainer.step(b["ids"], b["labels"], b["trans_ids"], b["seg"], b["trans_seg"], b["lens"], b["trans_lens"])
        metrics.update(trainer.last_head.decode, b["labels"], losses)
        n += len(b["lens"])
    mean_loss, _, _ = metrics.result()
    assert n == 48 and np.isfinite(mean_loss) and metrics.counters[4].item() == 48


def test_bucketed_adam_on_side_stream_equals_single_step_update():
    """DataParallelTrainer(overlap_optimizer=True) steps BertAdam bucket by bucket on a side stream as soon as a bucket's
    gradients are final (the data-parallel default); it must land on the same weights, moments and step counters as the
    single optimizer.step() after the backward (clipping is per tensor, so the grouping cannot matter)."""
    from nbest_b200 import epoch as E
    from nbest_b200.trainer import DataParallelTrainer
    res = []
    for overlap in (False, True):
        model, optim, opt, memory, meta, raw_in, raw_trans, *_ = _setup(layers=3)
        tr = DataParallelTrainer(model, optim, overlap_optimizer=overlap)
        assert tr.overlap_optimizer == overlap
        model.train()
        for labels, rin, rtr, ll in _batches(E, meta, raw_in, raw_trans, 16):
            from nbest_b200.inputs import prepare_inputs_for_roberta as prep
            ids, seg, lens = prep(rin, opt.tokenizer, opt, "cuda")
            tids, tseg, tlens = prep(rtr, opt.tokenizer, opt, "cuda")
            losses = tr.step(ids, labels.cuda(), tids, seg, tseg, lens, tlens)
        # a step without the transcript stream (the reference never uses it when --add_l2_loss is off)
        losses = tr.step(ids, labels.cuda(), None, seg, None, lens, None)
        torch.cuda.synchronize()
        assert torch.isfinite(losses).all()
        res.append((model.flat.params.clone(), optim.flat.m.clone(), optim.flat.v.clone(), list(optim._steps), losses.clone()))
    a, b = res
    cos = lambda x, y: float((x.double() @ y.double()) / (x.double().norm() * y.double().norm()))
    assert a[3] == b[3] and max(a[3]) == 4
    assert cos(a[0], b[0]) > 1 - 1e-6 and cos(a[1], b[1]) > 0.999 and cos(a[2], b[2]) > 0.999    # fp32 atomic order of wgrad
    assert torch.allclose(a[4], b[4], rtol=2e-3)
