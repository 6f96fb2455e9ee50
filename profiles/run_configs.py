"""Throughput of the other BASELINE.json configurations on one B200 (bench.py measures configs[1], the headline).

    python profiles/run_configs.py [xlmr] [l2] [infer10] [dense]

  xlmr     configs[2] shape: XLM-RoBERTa-base (250,002 vocab), mask_mode=reference (pads attendable, <s> masked), B = 256
  l2       configs[3] shape: BERT-base with --add_l2_loss (both streams carry gradients + MSE on the CLS vectors), B = 256
  infer10  configs[4] shape: BERT-base 10-hypothesis n-best, max_len 512, B = 512, forward + decode only
  dense    configs[1] worst case: every ASR sequence exactly 128 tokens (roofline figure)
Synthetic DSTC2-shaped ids, random-init weights, dropout on for training. CUDA-event timing, 3 warm-up + 10 timed steps.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC  # noqa: E402
from nbest_b200.optim import BertAdam  # noqa: E402
from nbest_b200.synth import synth_batch  # noqa: E402
from nbest_b200.trainer import DataParallelTrainer  # noqa: E402

hj = json.load(open(os.path.join(ROOT, "tests", "golden", "dstc2_hierarchy.json")))
T2B = {int(k): v for k, v in hj["top2bottom"].items()}


def timed(fn, warm=4, steps=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_cfg(name, spec, kind, B, hyps, max_len, l2, dense=False):
    model = TOD_ASR_Transformer_STC(spec=spec, top2bottom=T2B, dropout=0.3, device="cuda", none_bottoms=hj["none_bottoms"])
    model.train()
    groups = [dict(params=p, lr=3e-5, weight_decay=0.0 if ("bias" in n or "LayerNorm" in n) else 0.01)
              for n, p in model.named_parameters()]
    opt = BertAdam(groups, lr=3e-5, warmup=0.1, t_total=2300)
    tr = DataParallelTrainer(model, opt, add_l2_loss=l2)
    batches = []
    for i in range(4):
        b = synth_batch(kind, spec.vocab_size, model.hier, B, hyps, max_len, seed=10 + i, dense=dense)
        batches.append(({k: b[k].cuda() for k in ("ids", "seg", "trans_ids", "trans_seg", "labels")}, b["lens"], b["trans_lens"]))

    def step(i):
        d, lens, tl = batches[i % 4]
        return tr.step(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], lens, tl)

    ms = timed(step)
    losses = step(0)
    assert bool(torch.isfinite(losses).all())
    toks = np.mean([sum(b[1]) for b in batches])
    print(json.dumps(dict(config=name, utterances_per_s=B / (ms * 1e-3), ms_per_step=ms, batch=B, asr_tokens_per_step=float(toks),
                          params=sum(p.numel() for p in model.parameters()))))
    del model, opt, tr
    torch.cuda.empty_cache()


def infer_cfg():
    spec = EncoderSpec.bert_base()
    model = TOD_ASR_Transformer_STC(spec=spec, top2bottom=T2B, dropout=0.3, device="cuda", none_bottoms=hj["none_bottoms"])
    model.eval()
    B = 512
    batches = []
    for i in range(4):
        b = synth_batch("bert", spec.vocab_size, model.hier, B, 10, 512, seed=20 + i, with_trans=False)
        batches.append((b["ids"].cuda(), b["seg"].cuda(), b["lens"]))

    def step(i):
        ids, seg, lens = batches[i % 4]
        return model.infer(ids, seg, lens)

    ms = timed(step)
    toks = np.mean([sum(b[2]) for b in batches])
    print(json.dumps(dict(config="infer10: BERT-base 10-best, max_len 512, B=512, forward + decode", utterances_per_s=B / (ms * 1e-3),
                          ms_per_step=ms, tokens_per_step=float(toks), max_len=int(max(max(b[2]) for b in batches)))))


if __name__ == "__main__":
    which = sys.argv[1:] or ["xlmr", "l2", "infer10", "dense"]
    if "l2" in which:
        train_cfg("l2: BERT-base --add_l2_loss, B=256", EncoderSpec.bert_base(), "bert", 256, 5, 128, True)
    if "dense" in which:
        train_cfg("dense: BERT-base, every sequence 128 tokens, B=256", EncoderSpec.bert_base(), "bert", 256, 5, 128, False, dense=True)
    if "xlmr" in which:
        train_cfg("xlmr: XLM-R-base 250k vocab, mask_mode=reference, B=256", EncoderSpec.xlmr_base(), "xlm-roberta", 256, 5, 128, False)
    if "infer10" in which:
        infer_cfg()
