"""Every kernel of libnbest_sm100.so once, on unit shapes — the workload the compute-sanitizer logs under profiles/ are taken on:

    compute-sanitizer --tool memcheck  python profiles/sanitizer_unit.py
    compute-sanitizer --tool racecheck python profiles/sanitizer_unit.py
    compute-sanitizer --tool synccheck python profiles/sanitizer_unit.py

Small 2-layer encoders (BERT and XLM-R layouts), dropout on, --add_l2_loss on, one batch with a sequence longer than 128
tokens (block-loop attention kernels) next to short ones (tcgen05 tile kernels), training step + BertAdam / AdamW / Adam
updates, inference + device metrics, dynamic and static GEMM scheduling. Prints the distinct kernels it launched."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbest_b200 import _lib, ops                                   # noqa: E402
from nbest_b200.epoch import EpochMetrics                          # noqa: E402
from nbest_b200.model import EncoderSpec, TOD_ASR_Transformer_STC  # noqa: E402
from nbest_b200.optim import Adam, AdamW, BertAdam, clip_grad_norm_  # noqa: E402
from nbest_b200.synth import synth_batch                           # noqa: E402

hj = json.load(open(os.path.join(ROOT, "tests", "golden", "dstc2_hierarchy.json")))
t2b = {int(k): v for k, v in hj["top2bottom"].items()}
dev = torch.device("cuda", 0)


def run(kind, dynamic):
    _lib.context(0).set_gemm_dynamic(dynamic)
    mk = EncoderSpec.xlmr_base if kind == "xlm-roberta" else EncoderSpec.bert_base
    spec = mk(layers=2, vocab_size=3000, max_position=320)
    model = TOD_ASR_Transformer_STC(spec=spec, top2bottom=t2b, dropout=0.3, device=dev, none_bottoms=hj["none_bottoms"], seed=5)
    model.train()
    b = synth_batch(kind, 3000, model.hier, B=12, n_hyps=10, max_len=300, seed=3)        # some sequences > 128 tokens
    d = {k: b[k].to(dev) for k in ("ids", "seg", "trans_ids", "trans_seg", "labels")}
    for Opt, kw in ((BertAdam, dict(lr=1e-3, warmup=0.1, t_total=10)), (AdamW, dict(lr=1e-3)), (Adam, dict(lr=1e-3, weight_decay=0.01))):
        opt = Opt([dict(params=p, lr=1e-3, weight_decay=0.01) for p in model.parameters()], **kw)
        opt.zero_grad()
        losses, head = model.forward_loss_backward(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], add_l2_loss=True)
        if Opt is not BertAdam:
            clip_grad_norm_(list(model.parameters()), 5.0, optimizer=opt)
        opt.step()
        assert bool(torch.isfinite(losses).all())
    # autograd drop-in path (stc_scores_bwd) and the full last layer (no CLS-only shortcut)
    model.cls_only_last_layer = False
    opt_ns = type("Opt", (), dict(pre_trained_model=kind))()
    top, bottoms, final, asr, trans = model(opt_ns, d["ids"], d["trans_ids"], seg_ids=d["seg"], trans_seg_ids=d["trans_seg"])
    (final.sum() + top.sum() + asr.sum() + trans.sum()).backward()
    model.cls_only_last_layer = True
    # inference + device metrics + hypothesis-id map, block-loop kernels only (NBEST_ATTN_TC=0 equivalent)
    m = EpochMetrics(dev)
    head = model.infer(d["ids"], d["seg"])
    m.update(head.decode, d["labels"])
    model.attn_tensor_path = False
    losses, head = model.forward_loss_backward(d["ids"], d["labels"], d["trans_ids"], d["seg"], d["trans_seg"], add_l2_loss=False)
    model.attn_tensor_path = True
    pk = ops.pack_batch(d["ids"], d["seg"], kind)
    ops.pack_hyp_ids(pk, 2 if kind == "xlm-roberta" else 102)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(model.flat.params).all())
    return m.result()


for kind in ("bert", "xlm-roberta"):
    for dynamic in (False, True):
        print(kind, "dynamic GEMM scheduling" if dynamic else "static GEMM scheduling", run(kind, dynamic))
print("launches:", _lib.context(0).launches())
