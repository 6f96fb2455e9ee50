"""LayerNorm / column-sum micro-benchmark at the bench workload's token counts (CUDA events, buffers rotated)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbest_b200 import ops
H = 768
def timeit(fn, reps=40):
    for i in range(3): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for T in (17920, 12160):
    xs = [torch.randn(T, H, device="cuda").to(torch.bfloat16) for _ in range(6)]
    dys = [torch.randn(T, H, device="cuda").to(torch.bfloat16) for _ in range(6)]
    g, b = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    y = torch.empty_like(xs[0]); mean = torch.empty(T, device="cuda"); rstd = torch.empty(T, device="cuda")
    us = timeit(lambda i: ops.ln_fwd(xs[i % 6], g, b, 1e-12, y, mean, rstd))
    print("T=%d ln_fwd  %6.1f us  %5.0f GB/s (4 B/elem)" % (T, us, T * H * 4 / us / 1e3))
    part = torch.zeros(T, H // 64, 2, device="cuda")
    part[:, :, 0] = xs[0].float().view(T, H // 64, 64).sum(-1); part[:, :, 1] = (xs[0].float() ** 2).view(T, H // 64, 64).sum(-1)
    us = timeit(lambda i: ops.ln_fwd(xs[i % 6], g, b, 1e-12, y, mean, rstd, row_partials=part))
    print("T=%d ln_fwd with GEMM-epilogue row statistics  %6.1f us  %5.0f GB/s" % (T, us, T * H * 4 / us / 1e3))
    dx, dxm = torch.empty_like(y), torch.empty_like(y)
    dg, db, dbi = (torch.zeros(H, device="cuda") for _ in range(3))
    us = timeit(lambda i: ops.ln_bwd(dys[i % 6], xs[i % 6], mean, rstd, g, dx, dg, db, dx_masked=dxm, dbias=dbi, p_drop=0.1, seed=i))
    print("T=%d ln_bwd  %6.1f us  %5.0f GB/s (8 B/elem)" % (T, us, T * H * 8 / us / 1e3))
    for N in (3072, 2304):
        ws = [torch.randn(T, N, device="cuda").to(torch.bfloat16) for _ in range(3)]
        o = torch.zeros(N, device="cuda")
        us = timeit(lambda i: ops.colsum(ws[i % 3], o))
        print("T=%d colsum N=%d %6.1f us  %5.0f GB/s" % (T, N, us, T * N * 2 / us / 1e3))
        del ws
for T in (17920, 12160):
    xs = [torch.randn(T, H, device="cuda").to(torch.bfloat16) for _ in range(6)]
    y = torch.empty_like(xs[0])
    us = timeit(lambda i: y.copy_(xs[i % 6]))
    print("T=%d torch copy_ (same bytes as ln_fwd) %6.1f us  %5.0f GB/s" % (T, us, T * H * 4 / us / 1e3))
