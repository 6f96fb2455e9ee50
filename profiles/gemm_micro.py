"""GEMM micro-benchmark: TFLOP/s of nbest_gemm_bf16 for the encoder's shapes (CUDA events, operands rotated)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbest_b200 import ops

def bench(M, N, K, epi=ops.EPI_NONE, a_mn=False, b_mn=False, reps=20, **kw):
    nb = 4
    if a_mn:
        A = [torch.randn(K, M, device="cuda").to(torch.bfloat16) for _ in range(nb)]
    else:
        A = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nb)]
    Bm = [torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16) * 0.02 for _ in range(nb)]
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi == ops.EPI_ACCUM_F32 else torch.bfloat16)
    for i in range(3):
        ops.gemm(A[i % nb], Bm[i % nb], a_mn_major=a_mn, b_mn_major=b_mn, epilogue=epi, out=out, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.gemm(A[i % nb], Bm[i % nb], a_mn_major=a_mn, b_mn_major=b_mn, epilogue=epi, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12

if __name__ == "__main__":
    T = 17920
    for name, args in [("big K=3072 N=3072", (16384, 3072, 3072)), ("ffn2 fwd", (T, 768, 3072)), ("qkv fwd", (T, 2304, 768)),
                       ("ffn1 nobias", (T, 3072, 768)), ("outproj", (T, 768, 768))]:
        us, tf = bench(*args)
        print("%-22s stages=%s  %8.1f us  %7.1f TFLOP/s" % (name, os.environ.get("NBEST_GEMM_STAGES", "max"), us, tf))
    bias = torch.randn(3072, device="cuda")
    u = torch.empty(T, 3072, device="cuda", dtype=torch.bfloat16)
    print("ffn1 bias+gelu(+u)     %8.1f us  %7.1f TFLOP/s" % bench(T, 3072, 768, ops.EPI_BIAS_GELU, bias=bias, out2=u))
    print("ffn1 bias+gelu (no u)  %8.1f us  %7.1f TFLOP/s" % bench(T, 3072, 768, ops.EPI_BIAS_GELU, bias=bias))
    print("ffn1 bias only         %8.1f us  %7.1f TFLOP/s" % bench(T, 3072, 768, ops.EPI_BIAS, bias=bias))
    T2 = 12160
    uu = torch.randn(T2, 3072, device="cuda").to(torch.bfloat16)
    print("ffn2 dgrad dgelu       %8.1f us  %7.1f TFLOP/s" % bench(T2, 3072, 768, ops.EPI_DGELU, b_mn=True, aux=uu))
    print("ffn2 dgrad none        %8.1f us  %7.1f TFLOP/s" % bench(T2, 3072, 768, ops.EPI_NONE, b_mn=True))
    r2 = torch.randn(T2, 768, device="cuda").to(torch.bfloat16)
    print("ffn1 dgrad +res        %8.1f us  %7.1f TFLOP/s" % bench(T2, 768, 3072, ops.EPI_ADD, b_mn=True, aux=r2))
    print("outproj dgrad          %8.1f us  %7.1f TFLOP/s" % bench(T2, 768, 768, ops.EPI_NONE, b_mn=True))
    print("qkv dgrad +res         %8.1f us  %7.1f TFLOP/s" % bench(T2, 768, 2304, ops.EPI_ADD, b_mn=True, aux=r2))
    r = torch.randn(T, 768, device="cuda").to(torch.bfloat16)
    b768 = torch.randn(768, device="cuda")
    print("outproj bias+drop+res  %8.1f us  %7.1f TFLOP/s" % bench(T, 768, 768, ops.EPI_BIAS_DROP_RES, bias=b768, aux=r, p_drop=0.1, seed=1))
    print("outproj bias+res p=0   %8.1f us  %7.1f TFLOP/s" % bench(T, 768, 768, ops.EPI_BIAS_DROP_RES, bias=b768, aux=r))
    print("ffn2 bias+drop+res     %8.1f us  %7.1f TFLOP/s" % bench(T, 768, 3072, ops.EPI_BIAS_DROP_RES, bias=b768, aux=r, p_drop=0.1, seed=1))
    for (no, ki) in [(768, 3072), (3072, 768), (2304, 768), (768, 768)]:
        print("wgrad %4dx%4d         %8.1f us  %7.1f TFLOP/s" % ((no, ki) + bench(no, ki, T2, ops.EPI_ACCUM_F32, a_mn=True, b_mn=True)))
