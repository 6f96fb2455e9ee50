"""tcgen05 GEMM (nbest_gemm_bf16) against a plain fp32 torch reference of the same op, through the C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _dgelu(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-12)


def _mk(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (300, 768, 768), (1000, 2304, 768),
                                   (77, 3072, 768), (515, 768, 3072), (4096, 384, 256)])
def test_gemm_nt_bias(M, N, K):
    from nbest_b200 import ops
    a, w = _mk((M, K), 1.0, 1), _mk((N, K), 0.05, 2)
    bias = torch.randn(N, device="cuda")
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS, bias=bias)
    torch.cuda.synchronize()
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert _rel(out, ref) < 1e-2
    out0 = ops.gemm(a, w, epilogue=ops.EPI_NONE)
    assert _rel(out0, a.float() @ w.float().t()) < 1e-2


def test_gemm_nt_gelu_two_outputs():
    from nbest_b200 import ops
    M, N, K = 333, 3072, 768
    a, w = _mk((M, K), 1.0, 3), _mk((N, K), 0.04, 4)
    bias = torch.randn(N, device="cuda") * 0.1
    u_ref = a.float() @ w.float().t() + bias
    d = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    g = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, out2=d)
    # out2 = gelu'(u): the derivative (not u) is what the forward saves for the backward's multiply-only epilogue
    assert _rel(d, _dgelu(u_ref)) < 1e-2
    assert _rel(g, _gelu(u_ref)) < 1e-2
    g_only = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias)          # inference: no second output
    assert torch.equal(g, g_only)
    # gelu computed from the unrounded accumulator must match torch's exact-erf gelu tightly at small magnitudes too
    assert (g.float() - _gelu(u_ref)).abs().max().item() < 2e-2


def test_gemm_nt_residual_and_dropout():
    from nbest_b200 import ops
    M, N, K = 450, 768, 3072
    a, w, r = _mk((M, K), 1.0, 5), _mk((N, K), 0.02, 6), _mk((M, N), 1.0, 7)
    bias = torch.randn(N, device="cuda") * 0.1
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, bias=bias, aux=r)
    assert _rel(out, ref + r.float()) < 1e-2
    p = 0.25
    outd = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, bias=bias, aux=r, p_drop=p, seed=1234)
    d = outd.float() - r.float()                      # = keep ? ref/(1-p) : 0   (up to bf16 rounding)
    kept = (d.abs() > 0.5 * (ref.abs() / (1 - p))) & (ref.abs() > 0.05)
    dropped = (d.abs() <= 0.02 + 0.01 * r.float().abs()) & (ref.abs() > 0.05)
    considered = (ref.abs() > 0.05)
    assert ((kept | dropped) | ~considered).all()
    frac = dropped.sum().item() / considered.sum().item()
    assert abs(frac - p) < 0.01
    # same seed -> same mask, different seed -> different mask
    outd2 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, bias=bias, aux=r, p_drop=p, seed=1234)
    assert torch.equal(outd, outd2)
    outd3 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, bias=bias, aux=r, p_drop=p, seed=99)
    assert not torch.equal(outd, outd3)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 768, 768), (1000, 768, 3072), (515, 3072, 768), (64, 768, 2304)])
def test_gemm_dgrad_b_mn_major(M, N, K):
    """dx[M,N] = dy[M,K] @ W[K,N]  (W is an nn.Linear weight [out=K, in=N] read MN-major)."""
    from nbest_b200 import ops
    dy, w, r = _mk((M, K), 1.0, 8), _mk((K, N), 0.05, 9), _mk((M, N), 1.0, 10)
    ref = dy.float() @ w.float()
    out = ops.gemm(dy, w, b_mn_major=True, epilogue=ops.EPI_NONE)
    assert _rel(out, ref) < 1e-2
    out = ops.gemm(dy, w, b_mn_major=True, epilogue=ops.EPI_ADD, aux=r)
    assert _rel(out, ref + r.float()) < 1e-2
    d = _dgelu(_mk((M, N), 1.5, 11).float()).to(torch.bfloat16)            # what EPI_BIAS_GELU saved in the forward
    out = ops.gemm(dy, w, b_mn_major=True, epilogue=ops.EPI_DGELU, aux=d)
    assert _rel(out, ref * d.float()) < 1e-2


@pytest.mark.parametrize("T,NO,KI", [(64, 128, 128), (512, 128, 256), (1000, 768, 768), (3001, 3072, 768),
                                     (2500, 768, 3072), (777, 2304, 768)])
def test_gemm_wgrad_both_mn_major(T, NO, KI):
    """dW[NO,KI] += dy[T,NO]^T @ x[T,KI], fp32 accumulation with split-K over T."""
    from nbest_b200 import ops
    dy, x = _mk((T, NO), 1.0, 12), _mk((T, KI), 1.0, 13)
    ref = dy.float().t() @ x.float()
    acc = torch.zeros((NO, KI), device="cuda")
    ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_ACCUM_F32, out=acc)
    assert _rel(acc, ref) < 2e-3
    ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_ACCUM_F32, out=acc)   # accumulates
    assert _rel(acc, 2 * ref) < 2e-3


def test_gemm_rejects_bad_shapes():
    from nbest_b200 import ops
    from nbest_b200._lib import NbestError
    a, w = _mk((64, 64)), _mk((100, 64))
    with pytest.raises(NbestError):
        ops.gemm(a, w)


@pytest.mark.parametrize("M", [200, 1000, 12167])
def test_gemm_dgrad_with_fused_attention_delta(M):
    """NBEST_EPI_DELTA: C = dy W (dgrad, B MN-major) and, per 64-column unit (= attention head), delta[h, m] =
    sum_c C[m, 64h + c] * O[m, 64h + c] — the flash-attention backward preprocess fused into the out-projection dgrad."""
    from nbest_b200 import ops
    N, K = 768, 768
    dy = _mk((M, K), 1.0, 1)
    w = _mk((K, N), 0.05, 2)                      # [K, N] row-major = MN-major B
    o = _mk((M, N), 1.0, 3)
    delta = torch.full((N // 64, M), float("nan"), device="cuda")
    out = ops.gemm(dy, w, b_mn_major=True, epilogue=ops.EPI_DELTA, aux=o, out2=delta)
    ref = dy.float() @ w.float()
    assert _rel(out, ref) < 1e-2
    ref_delta = (ref * o.float()).view(M, N // 64, 64).sum(-1).t()
    assert torch.isfinite(delta).all()
    assert _rel(delta, ref_delta) < 5e-3


@pytest.mark.parametrize("M", [130, 1000, 12167])
def test_gemm_dgelu_with_fused_bias_gradient(M):
    """NBEST_EPI_DGELU with out2: C = (dy W) * aux (aux = gelu'(u)) and out2[n] += sum_m C[m, n] (accumulating, fp32)."""
    from nbest_b200 import ops
    N, K = 3072, 768
    dy = _mk((M, K), 1.0, 4)
    w = _mk((K, N), 0.05, 5)
    u = _dgelu(_mk((M, N), 1.5, 6).float()).to(torch.bfloat16)
    acc = torch.full((N,), 0.25, device="cuda")
    out = ops.gemm(dy, w, b_mn_major=True, epilogue=ops.EPI_DGELU, aux=u, out2=acc)
    ref = (dy.float() @ w.float()) * u.float()
    assert _rel(out, ref) < 1.5e-2
    assert _rel(acc - 0.25, out.float().sum(0)) < 2e-3          # exactly the column sums of what was stored (bf16)
    out_b = ops.gemm(dy, w, b_mn_major=True, epilogue=ops.EPI_DGELU, aux=u)      # without out2: unchanged result
    assert torch.equal(out, out_b)
