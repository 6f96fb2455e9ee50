"""One GEMM shape for ncu captures: FFN-in with the fused bias + erf-GELU epilogue (two outputs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from profiles.gemm_micro import bench
from nbest_b200 import ops
T = 17920
which = sys.argv[1] if len(sys.argv) > 1 else "gelu"
if which == "gelu":
    bias = torch.randn(3072, device="cuda")
    u = torch.empty(T, 3072, device="cuda", dtype=torch.bfloat16)
    print("ffn1 bias+gelu(+u) %8.1f us %7.1f TFLOP/s" % bench(T, 3072, 768, ops.EPI_BIAS_GELU, bias=bias, out2=u, reps=5))
elif which == "dgelu":
    uu = torch.randn(12160, 3072, device="cuda").to(torch.bfloat16)
    print("ffn2 dgrad dgelu %8.1f us %7.1f TFLOP/s" % bench(12160, 3072, 768, ops.EPI_DGELU, b_mn=True, aux=uu, reps=5))
