"""Attention micro-benchmark on the bench workload's sequence lengths (B = 256 ASR + 256 transcript sequences), dropout 0.1."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbest_b200 import ops
from nbest_b200.synth import synth_batch

hj = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dstc2_hierarchy.json")))
hier = ops.DeviceHierarchy({int(k): v for k, v in hj["top2bottom"].items()}, hj["none_bottoms"], device="cuda")
hyps = int(os.environ.get("HYPS", "5")); max_len = int(os.environ.get("MAXLEN", "128")); Bq = int(os.environ.get("BATCH", "256"))
b = synth_batch("bert", 30522, hier, B=Bq, n_hyps=hyps, max_len=max_len, seed=999)
la = (b["ids"] > 0).sum(1).numpy(); lt = (b["trans_ids"] > 0).sum(1).numpy()
heads = 12

def run(lens, name, bwd=True, reps=30):
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32); T = int(cu[-1]); B = len(lens)
    g = torch.Generator(device="cuda").manual_seed(1)
    qkvs = [torch.randn(T, 3 * heads * 64, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3)]
    cu_d = torch.from_numpy(cu).cuda()
    out = torch.empty(T, heads * 64, device="cuda", dtype=torch.bfloat16); lse = torch.empty(heads, T, device="cuda")
    dout = torch.randn(T, heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    dqkv = torch.empty_like(qkvs[0]); delta = torch.empty(heads, T, device="cuda")
    def f(i): ops.attn_fwd(qkvs[i % 3], cu_d, None, B, int(max(lens)), heads, T, out, lse, p_drop=0.1, seed=i)
    def bw(i): ops.attn_bwd(qkvs[i % 3], cu_d, None, B, int(max(lens)), heads, T, out, dout, lse, dqkv, delta, p_drop=0.1, seed=i)
    seq_of = torch.repeat_interleave(torch.arange(B, dtype=torch.int32), torch.from_numpy(np.asarray(lens)).long()).cuda()
    plan = ops.attn_plan(cu_d, seq_of, B, T)
    long_ = int(max(lens)) > 128
    def ft(i):
        ops.attn_tiles_fwd(qkvs[i % 3], plan, 0, None, heads, T, out, lse, p_drop=0.1, seed=i)
        if long_: ops.attn_fwd(qkvs[i % 3], cu_d, None, B, int(max(lens)), heads, T, out, lse, p_drop=0.1, seed=i, min_len=129)
    def bt(i):
        ops.attn_tiles_bwd(qkvs[i % 3], plan, 0, None, heads, T, T, dout, lse, delta, T, dqkv, p_drop=0.1, seed=i)
        if long_: ops.attn_bwd(qkvs[i % 3], cu_d, None, B, int(max(lens)), heads, T, None, dout, lse, dqkv, delta, p_drop=0.1, seed=i, min_len=129)
    res = []
    for fn in ([f, bw, ft, bt] if bwd else [f, ft]):
        for i in range(3): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps): fn(i)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / reps * 1e3)
    hbm_f, hbm_b = T * 768 * 2 * 4 / 6545.6e3, T * 768 * 2 * 7 / 6545.6e3      # us at the measured HBM peak
    if bwd:
        print("%-28s B=%4d T=%6d max=%3d  mma.sync fwd %7.1f bwd %7.1f us | tcgen05 tiles fwd %7.1f bwd %7.1f us | HBM floor %5.1f / %5.1f us (n_tiles %d)" % (
            name, B, T, max(lens), res[0], res[1], res[2], res[3], hbm_f, hbm_b, int(plan.counts[0])))
    else:
        print("%-28s B=%4d T=%6d max=%3d  mma.sync fwd %7.1f us | tcgen05 tiles fwd %7.1f us | HBM floor %5.1f us (n_tiles %d)" % (
            name, B, T, max(lens), res[0], res[1], hbm_f, int(plan.counts[0])))

run(np.concatenate([la, lt]), "fwd shape (asr+transcript)", bwd=False)
run(la, "bwd shape (asr only)")
run(np.minimum(la, 64), "asr clipped to 64")
