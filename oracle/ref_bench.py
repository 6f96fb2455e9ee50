"""Times THE REFERENCE'S OWN training / inference step (TEST + BENCH INFRASTRUCTURE; bench.py's `--impl reference`,
`cpu_baseline` and `gpu_eager_baseline` legs are the only callers).

What runs is the unmodified reference staged under oracle/_ref/ (oracle/vendor_ref.py), driven exactly as its train_epoch
does (n_best_asr_bert.py:242-280): `models.model.make_model(opt)` around a random-init HuggingFace encoder,
`model(opt, ids, trans_ids, seg_ids=..., trans_seg_ids=...)`, `cal_total_loss`, `total_loss.backward()`,
`models.optimization.BertAdam.step()` over the 221 one-tensor groups of :535-550, `zero_grad()`. Inputs are the same
synthetic token-id batches the CUDA path gets (nbest_b200.synth — data generation only), so string tokenisation is
excluded on both sides.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _hier_json():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "dstc2_hierarchy.json")) as f:
        return json.load(f)


def build(device, kind="bert", add_l2_loss=False, seed=999, layers=None, lr=3e-5, t_total=2300):
    """-> (ref namespace, model, opt, memory). Hyper-parameters of run/train_eval_N_Best_ASR_Transformer_STC.sh:31-44."""
    from oracle import ref_loader as R
    ref = R.load()
    torch.manual_seed(seed)
    kw = {} if layers is None else dict(num_hidden_layers=layers)
    enc = R.hf_encoder(kind, eager=False, **kw)                     # attn_implementation default (sdpa), BASELINE.md §3
    mem = R.memory(device)
    opt = R.make_opt(enc, mem, device, pre_trained_model=kind, dropout=0.3, add_l2_loss=add_l2_loss)
    model = ref.make_model(opt).to(opt.device)
    groups = []
    for n, p in model.named_parameters():                           # n_best_asr_bert.py:535-550
        no_decay = any(nd in n for nd in ("bias", "LayerNorm.bias", "LayerNorm.weight"))
        groups.append(dict(params=p, weight_decay=0.0 if no_decay else 0.01, lr=lr))
    opt.optimizer = ref.BertAdam(groups, lr=lr, warmup=0.1, t_total=t_total)
    return ref, model, opt, mem


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


def time_train(device, n_utt, steps, warmup, kind="bert", add_l2_loss=False, n_hyps=5, max_len=128, threads=None,
               autocast_encoder=False, seed=999, layers=None):
    """Median step time of the reference training step on `device` + its fwd / bwd / optimizer split.
    autocast_encoder: run the HF encoder alone under torch.autocast(bf16) (the STC head's BCELoss is not autocast-safe,
    SURVEY §4) — the informational "stock PyTorch, reduced precision" point of SURVEY §8(d)."""
    from nbest_b200.synth import synth_batch
    from oracle import stc_oracle as O
    dev = torch.device(device)
    if dev.type == "cpu":
        threads = threads or os.cpu_count()
        torch.set_num_threads(threads)
    ref, model, opt, mem = build(device, kind, add_l2_loss, seed, layers)
    hj = _hier_json()
    hier = O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"])
    vocab = model.bert_encoder.config.vocab_size
    model.train()
    if autocast_encoder:
        inner = model.bert_encoder.forward

        def wrapped(*a, **k):
            with torch.autocast(dev.type, dtype=torch.bfloat16):
                out = inner(*a, **k)
            return (out[0].float(),)                                # models/model.py:46,57 only read outputs[0]
        model.bert_encoder.forward = wrapped
    split, times = [], []
    sink = io.StringIO()
    for i in range(warmup + steps):
        b = synth_batch("bert" if kind == "bert" else "xlm-roberta", vocab, hier, n_utt, n_hyps, max_len, seed + i)
        d = {k: b[k].to(dev) for k in ("ids", "seg", "trans_ids", "trans_seg", "labels")}
        seg = d["seg"] if kind == "bert" else None
        tseg = d["trans_seg"] if kind == "bert" else None
        _sync(dev)
        t0 = time.perf_counter()
        top, bottoms, final, asr, trans = model(opt, d["ids"], d["trans_ids"], seg_ids=seg, trans_seg_ids=tseg,
                                                classifier_input_type="asr")
        with contextlib.redirect_stdout(sink):                      # cal_total_loss prints the MSE term
            rec, total = ref.nb.cal_total_loss(top, bottoms, final, d["labels"], mem, opt, asr, trans)
        _sync(dev)
        t1 = time.perf_counter()
        total.backward()
        _sync(dev)
        t2 = time.perf_counter()
        opt.optimizer.step()
        opt.optimizer.zero_grad()
        _sync(dev)
        t3 = time.perf_counter()
        if i >= warmup:
            times.append(t3 - t0)
            split.append((t1 - t0, t2 - t1, t3 - t2))
    sp = np.median(np.asarray(split), axis=0)
    return dict(median_s=float(np.median(times)), total_s=float(np.sum(times)), fwd_s=float(sp[0]), bwd_s=float(sp[1]),
                opt_s=float(sp[2]), threads=threads if dev.type == "cpu" else None, n_utt=n_utt)


def time_infer(device, n_utt, steps, warmup, kind="bert", n_hyps=10, max_len=512, threads=None, seed=999):
    """Median time of the reference inference step (eval_epoch body, n_best_asr_bert.py:316-350: forward in eval mode +
    pred_one_sample per utterance; both encoder streams as the reference runs them)."""
    from nbest_b200.synth import synth_batch
    from oracle import stc_oracle as O
    dev = torch.device(device)
    if dev.type == "cpu":
        threads = threads or os.cpu_count()
        torch.set_num_threads(threads)
    ref, model, opt, mem = build(device, kind, False, seed)
    hj = _hier_json()
    hier = O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"])
    vocab = model.bert_encoder.config.vocab_size
    model.eval()
    times = []
    for i in range(warmup + steps):
        b = synth_batch("bert" if kind == "bert" else "xlm-roberta", vocab, hier, n_utt, n_hyps, max_len, seed + i)
        d = {k: b[k].to(dev) for k in ("ids", "seg", "trans_ids", "trans_seg")}
        seg = d["seg"] if kind == "bert" else None
        tseg = d["trans_seg"] if kind == "bert" else None
        _sync(dev)
        t0 = time.perf_counter()
        with torch.no_grad():
            top, bottoms, final, asr, trans = model(opt, d["ids"], d["trans_ids"], seg_ids=seg, trans_seg_ids=tseg)
        for j, ts in enumerate(top.tolist()):
            ref.nb.pred_one_sample(j, ts, bottoms, mem, opt)
        _sync(dev)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return dict(median_s=float(np.median(times)), total_s=float(np.sum(times)), threads=threads if dev.type == "cpu" else None,
                n_utt=n_utt)
