"""Torch-tensor wrappers over the C ABI (include/nbest_sm100.h).

Each function passes raw device pointers and the current CUDA stream to libnbest_sm100.so. There is no fallback:
a missing library or a non-CUDA tensor raises.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ACCUM_F32, EPI_ADD, EPI_BIAS, EPI_BIAS_DROP_RES, EPI_BIAS_GELU, EPI_DGELU, EPI_NONE,  # noqa: F401
                   AdamTensor, Hierarchy)


def _ctx(t):
    if not t.is_cuda:
        raise RuntimeError("nbest_b200 ops need CUDA tensors (no CPU fallback)")
    return _lib.context(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_bf16(t, name):
    if t.dtype != torch.bfloat16 or t.stride(-1) != 1:
        raise ValueError("%s must be a bf16 tensor with unit inner stride" % name)


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, epilogue=EPI_NONE, bias=None, aux=None, out=None, out2=None,
         p_drop=0.0, seed=0, M=None, N=None, K=None):
    """C[M,N] = epilogue(sum_k A(m,k) B(n,k)) on the tcgen05 GEMM. See nbest_gemm_bf16 in the header.

    a: [M,K] (or [K,M] if a_mn_major); b: [N,K] (or [K,N] if b_mn_major); bf16, 2-D, unit inner stride.
    EPI_ACCUM_F32 accumulates into `out` (fp32), every other epilogue writes bf16.
    """
    _check_bf16(a, "a")
    _check_bf16(b, "b")
    if M is None:
        M = a.shape[1] if a_mn_major else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn_major else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn_major else b.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device,
                          dtype=torch.float32 if epilogue == EPI_ACCUM_F32 else torch.bfloat16)
    ctx = _ctx(a)
    rc = _lib.lib().nbest_gemm_bf16(
        ctx.handle, _p(a), a.stride(0), int(a_mn_major), _p(b), b.stride(0), int(b_mn_major), _p(out), out.stride(0),
        M, N, K, int(epilogue), _p(bias), _p(aux), aux.stride(0) if aux is not None else 0, _p(out2), float(p_drop),
        int(seed) & 0xFFFFFFFF, _stream())
    ctx.check(rc)
    return out
