"""CPU, world_size 2 over gloo: the data-parallel exchange (nbest_b200.trainer.GradBucketer) and its parity rules.

Each rank computes the oracle's gradients on its half of the batch, the flat gradient buffer is all-reduced bucket by
bucket (SUM, asynchronously, in the order backward finishes the buckets), and the result must equal the full-batch
gradient of the reference loss: BCE / CE terms are sum-reduced so they need NO 1/R; the mean-reduced MSE term is scaled
by 1/R on each rank."""
import json
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    from oracle import stc_oracle as O
    from nbest_b200.synth import synth_batch
    hj = json.load(open(os.path.join(GOLD, "dstc2_hierarchy.json")))
    hier = O.Hierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"])
    cfg = O.EncoderConfig.bert_base(layers=1, vocab_size=600, max_position=64)
    params = O.init_params(cfg, hier, seed=5, style="perturbed")
    batch = synth_batch("bert", cfg.vocab_size, hier, B=4, n_hyps=2, max_len=48, seed=9)
    return O, hier, cfg, params, batch


def _grads(O, hier, cfg, params, batch, rows, mse_scale):
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    sub = {k: (v[rows] if torch.is_tensor(v) else v) for k, v in batch.items()}
    top, bottoms, final, asr, trans = O.model_forward(leaves, cfg, hier, sub["ids"], sub["trans_ids"], sub["seg"], sub["trans_seg"])
    _, terms = O.total_loss(hier, top, bottoms, final, sub["labels"], asr, trans, add_l2_loss=True)
    total = terms["bce_final"] + terms["bce_top"] + terms["ce"] + mse_scale * terms["mse"]
    total.backward()
    return {k: v.grad for k, v in leaves.items()}


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, ROOT)
    torch.set_num_threads(2)
    from nbest_b200.optim import FlatBuffers
    from nbest_b200.trainer import GradBucketer, init_distributed
    r, _, w = init_distributed()
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    O, hier, cfg, params, batch = _setup()
    names = [n for n in params if "pooler" not in n]                       # pooler: grad None, not in any bucket
    fb = FlatBuffers([tuple(params[n].shape) for n in names], "cpu", with_bf16=False)
    rows = torch.arange(rank * 2, rank * 2 + 2)
    g = _grads(O, hier, cfg, params, batch, rows, mse_scale=1.0 / world)
    for i, n in enumerate(names):
        fb.view(fb.grads, i).copy_(g[n])
    # buckets in backward order: head | layer 0 | embeddings
    first = {tag: min(i for i, n in enumerate(names) if n.startswith(pre)) for tag, pre in
             (("emb", "bert_encoder.embeddings"), ("layer0", "bert_encoder.encoder.layer.0"), ("head", "clf."))}
    bounds = sorted((fb.offsets[i], tag) for tag, i in first.items())
    segs = {tag: (off, bounds[k + 1][0] if k + 1 < len(bounds) else fb.total) for k, (off, tag) in enumerate(bounds)}
    bucketer = GradBucketer(fb.grads, [(t,) + segs[t] for t in ("head", "layer0", "emb")])
    assert bucketer.world == 2
    for tag in ("head", "layer0", "emb"):
        bucketer.reduce(tag)                                                 # asynchronous, like the overlap with backward
    bucketer.wait()
    torch.save({n: fb.view(fb.grads, i).clone() for i, n in enumerate(names)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_sum_allreduce_equals_full_batch_gradient(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    O, hier, cfg, params, batch = _setup()
    full = _grads(O, hier, cfg, params, batch, torch.arange(4), mse_scale=1.0)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for n, gref in full.items():
        if "pooler" in n:
            assert gref is None
            continue
        assert torch.equal(r0[n], r1[n]), n                                  # both ranks hold the same reduced gradient
        if "attention.self.key.bias" in n:
            continue
        err = float((r0[n] - gref).abs().max() / gref.abs().max().clamp_min(1e-30))
        assert err < 1e-4, (n, err)


def test_single_process_bucketer_is_a_noop():
    from nbest_b200.trainer import GradBucketer
    flat = torch.arange(10, dtype=torch.float32)
    b = GradBucketer(flat, [("a", 0, 4), ("b", 4, 10)])
    b.reduce("a")
    b.wait()
    assert torch.equal(flat, torch.arange(10, dtype=torch.float32))
