"""Torch-tensor wrappers over the C ABI (include/nbest_sm100.h).

Each function passes raw device pointers and the current CUDA stream to libnbest_sm100.so. There is no fallback:
a missing library or a non-CUDA tensor raises.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ACCUM_F32, EPI_ADD, EPI_BIAS, EPI_BIAS_DROP_RES, EPI_BIAS_GELU, EPI_DELTA, EPI_DGELU, EPI_NONE,  # noqa: F401
                   AdamTensor, Hierarchy)


# Optional per-call timing (bench.py's roofline leg): ops.profile_start() makes every wrapper bracket its launch with
# CUDA events on the launching stream; ops.profile_stop() returns [(name, ms, flops, bytes, meta)].
_PROF = None


def profile_start():
    global _PROF
    _PROF = []


def profile_stop():
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    return [(n, e0.elapsed_time(e1), fl, by, meta) for n, e0, e1, fl, by, meta in rec]


class _Timed:
    __slots__ = ("name", "flops", "bytes", "meta", "e0")

    def __init__(self, name, flops=0.0, nbytes=0.0, meta=None):
        self.name, self.flops, self.bytes, self.meta = name, flops, nbytes, meta

    def __enter__(self):
        if _PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _PROF is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROF.append((self.name, self.e0, e1, self.flops, self.bytes, self.meta))
        return False


class _NoTimed:
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NOTIMED = _NoTimed()


def _timed(name, flops=0.0, nbytes=0.0, meta=None):
    """Per-launch CUDA-event bracket while ops.profile_start() is active, a shared no-op object otherwise (the wrappers
    run ~250 times per step: at the reference's batch 16 the step is bound by this host code, not by the GPU)."""
    return _NOTIMED if _PROF is None else _Timed(name, flops, nbytes, meta)


def _ctx(t):
    if not t.is_cuda:
        raise RuntimeError("nbest_b200 ops need CUDA tensors (no CPU fallback)")
    return _lib.context(t.device.index if t.device.index is not None else torch.cuda.current_device())


import os as _os
_CHECK_LENS = _os.environ.get("NBEST_CHECK_LENS", "0") not in ("", "0")

_STREAM = None   # cached cudaStream_t of the stream the current step launches on (torch.cuda.current_stream() costs ~10 us)


def bind_stream():
    """Cache torch's current stream for the launches that follow. The model / trainer entry points call this once per
    step; code that switches streams in between must call it again (or `unbind_stream()` to query torch every time)."""
    global _STREAM
    _STREAM = torch.cuda.current_stream().cuda_stream


def unbind_stream():
    global _STREAM
    _STREAM = None


class on_stream:
    """Context manager: launches made inside go to `stream` (a torch.cuda.Stream), whatever stream the step is bound to."""

    def __init__(self, stream):
        self.handle = stream.cuda_stream

    def __enter__(self):
        global _STREAM
        self.prev = _STREAM
        _STREAM = self.handle

    def __exit__(self, *exc):
        global _STREAM
        _STREAM = self.prev
        return False


def with_bound_stream(fn):
    """Decorator for the public entry points: bind the current stream once for all launches made inside."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if _STREAM is not None:
            return fn(*a, **k)
        bind_stream()
        try:
            return fn(*a, **k)
        finally:
            unbind_stream()
    return wrapper


def _stream():
    return _STREAM if _STREAM is not None else torch.cuda.current_stream().cuda_stream


def _p(t):
    return t.data_ptr() if t is not None else None      # ctypes converts int / None for c_void_p parameters


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
    kind = "gemm_wgrad" if a_mn_major else ("gemm_dgrad" if b_mn_major else "gemm_fwd")
    with _timed(kind, 2.0 * M * N * K, 0.0, (M, N, K, int(epilogue))):
        rc = _lib.lib().nbest_gemm_bf16(
            ctx.handle, _p(a), a.stride(0), int(a_mn_major), _p(b), b.stride(0), int(b_mn_major), _p(out), out.stride(0),
            M, N, K, int(epilogue), _p(bias), _p(aux), aux.stride(0) if aux is not None else 0, _p(out2), float(p_drop),
            _seed(seed), _stream())
    ctx.check(rc)
    return out


def _seed(s):
    """32-bit seed handed to the kernels, avalanche-mixed (murmur3 finaliser): callers pass small structured values
    (step * 1000003 + layer * 16 + site) and the device-side quad hash xors the seed in after its first multiply."""
    s = int(s) & 0xFFFFFFFF
    s ^= s >> 16
    s = (s * 0x85EBCA6B) & 0xFFFFFFFF
    s ^= s >> 13
    s = (s * 0xC2B2AE35) & 0xFFFFFFFF
    s ^= s >> 16
    return s


# ---------------------------------------------------------------------------------------------------------- packing
class Packed:
    """Packed variable-length batch (device tensors). T is a host int (one D2H read of cu_seqlens[-1] unless given)."""
    __slots__ = ("B", "S", "T", "max_len", "lens", "cu_seqlens", "tokens", "seg", "pos", "seq_of", "key_valid",
                 "B_asr", "T_asr", "max_len_asr", "sum_l2", "sum_l2_asr", "plan", "plan_asr")


def pack_batch(ids, seg_ids=None, kind="bert", lens_host=None):
    """ids [B,S] int64 CUDA, right padded -> Packed. See nbest_pack_batch."""
    if ids.dtype != torch.int64 or ids.dim() != 2:
        raise ValueError("ids must be an int64 [B,S] tensor")
    ids = ids.contiguous()
    if seg_ids is not None:
        seg_ids = seg_ids.contiguous()
    B, S = ids.shape
    dev = ids.device
    pk = Packed()
    pk.B, pk.S = B, S
    pk.sum_l2 = pk.sum_l2_asr = 0.0
    pk.plan = pk.plan_asr = None
    pk.lens = torch.empty(B, dtype=torch.int32, device=dev)
    pk.cu_seqlens = torch.empty(B + 1, dtype=torch.int32, device=dev)
    cap = B * S
    pk.tokens = torch.empty(cap, dtype=torch.int32, device=dev)
    pk.seg = torch.empty(cap, dtype=torch.uint8, device=dev)
    pk.pos = torch.empty(cap, dtype=torch.int32, device=dev)
    pk.seq_of = torch.empty(cap, dtype=torch.int32, device=dev)
    pk.key_valid = torch.empty(cap, dtype=torch.uint8, device=dev)
    ctx = _ctx(ids)
    with _timed('pack_batch', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_pack_batch(ctx.handle, _p(ids), _p(seg_ids), B, S, 1 if kind == "xlm-roberta" else 0,
                                              _p(pk.lens), _p(pk.cu_seqlens), _p(pk.tokens), _p(pk.seg), _p(pk.pos),
                                              _p(pk.seq_of), _p(pk.key_valid), _stream()))
    if seg_ids is not None and (seg_ids.dtype != torch.int64 or seg_ids.shape != ids.shape or not seg_ids.is_cuda):
        raise ValueError("seg_ids must be an int64 CUDA tensor shaped like ids")
    if lens_host is not None and kind != "xlm-roberta":
        # the kernel derives the true lengths from the ids; the host list only saves the D2H read of T / max_len, so it
        # must describe THIS batch (input_lens of prepare_inputs_for_roberta): cheap structural checks here, the
        # device-side truth with NBEST_CHECK_LENS=1
        if len(lens_host) != B or min(lens_host) < 0 or max(lens_host) > S:
            raise ValueError("input_lens does not describe ids [%d, %d]: %d entries, max %s" % (B, S, len(lens_host), max(lens_host)))
        pk.T = int(sum(lens_host))
        pk.max_len = int(max(lens_host))
        pk.sum_l2 = float(sum(int(x) * int(x) for x in lens_host)) if _PROF is not None else 0.0
        if _CHECK_LENS and pk.lens.cpu().tolist() != [int(x) for x in lens_host]:
            raise ValueError("input_lens disagrees with the lengths derived from ids")
    else:
        lens_cpu = pk.lens.cpu()
        pk.T = int(lens_cpu.sum())
        pk.max_len = int(lens_cpu.max())
        pk.sum_l2 = float((lens_cpu.double() ** 2).sum())
    return pk


def pack_batch_dual(ids_a, seg_a, lens_a, ids_t, seg_t, lens_t, kind="bert"):
    """Both streams of a step -> ONE Packed batch (ASR sequences first), written in place by nbest_pack_batch_dual."""
    for t in (ids_a, ids_t):
        if t.dtype != torch.int64 or t.dim() != 2 or not t.is_cuda:
            raise ValueError("ids must be int64 [B,S] CUDA tensors")
    ids_a, ids_t = ids_a.contiguous(), ids_t.contiguous()
    for sg, ref in ((seg_a, ids_a), (seg_t, ids_t)):
        if sg is not None and (sg.dtype != torch.int64 or sg.shape != ref.shape or not sg.is_cuda):
            raise ValueError("seg_ids must be an int64 CUDA tensor shaped like ids")
    seg_a = seg_a.contiguous() if seg_a is not None else None
    seg_t = seg_t.contiguous() if seg_t is not None else None
    (Ba, Sa), (Bt, St) = ids_a.shape, ids_t.shape
    dev = ids_a.device
    pk = Packed()
    pk.B, pk.S = Ba + Bt, max(Sa, St)
    pk.plan = pk.plan_asr = None
    cap = Ba * Sa + Bt * St
    ws = torch.empty(3 * cap + 2 * (Ba + Bt) + 1, dtype=torch.int32, device=dev)      # one allocation for the int32 arrays
    pk.tokens, pk.pos, pk.seq_of = ws[:cap], ws[cap:2 * cap], ws[2 * cap:3 * cap]
    pk.lens = ws[3 * cap:3 * cap + Ba + Bt]
    pk.cu_seqlens = ws[3 * cap + Ba + Bt:]
    w8 = torch.empty(2 * cap, dtype=torch.uint8, device=dev)
    pk.seg, pk.key_valid = w8[:cap], w8[cap:]
    ctx = _ctx(ids_a)
    with _timed('pack_batch', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_pack_batch_dual(ctx.handle, _p(ids_a), _p(seg_a), Ba, Sa, _p(ids_t), _p(seg_t), Bt, St,
                                                   1 if kind == "xlm-roberta" else 0, _p(pk.lens), _p(pk.cu_seqlens),
                                                   _p(pk.tokens), _p(pk.seg), _p(pk.pos), _p(pk.seq_of), _p(pk.key_valid),
                                                   _stream()))
    if lens_a is not None and lens_t is not None and kind != "xlm-roberta":
        if len(lens_a) != Ba or len(lens_t) != Bt or max(lens_a) > Sa or max(lens_t) > St or min(lens_a) < 0 or min(lens_t) < 0:
            raise ValueError("input_lens do not describe the id tensors")
        la, lt = [int(x) for x in lens_a], [int(x) for x in lens_t]
        if _CHECK_LENS and pk.lens.cpu().tolist() != la + lt:
            raise ValueError("input_lens disagrees with the lengths derived from ids")
    else:
        lc = pk.lens.cpu().tolist()
        la, lt = lc[:Ba], lc[Ba:]
    pk.T_asr, pk.max_len_asr, pk.B_asr = sum(la), max(la), Ba
    pk.T, pk.max_len = pk.T_asr + sum(lt), max(pk.max_len_asr, max(lt))
    if _PROF is not None or lens_a is None:
        pk.sum_l2_asr = float(sum(x * x for x in la))
        pk.sum_l2 = pk.sum_l2_asr + float(sum(x * x for x in lt))
    else:
        pk.sum_l2 = pk.sum_l2_asr = 0.0
    return pk


def rows_gather(src, row_idx, n, out):
    """out[i] = src[row_idx[i]] for [*, 768] bf16 rows; out may be bf16 or fp32 (see nbest_rows_gather)."""
    ctx = _ctx(src)
    _check_bf16(src, "src")
    with _timed('rows_gather', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_rows_gather(ctx.handle, _p(src), _p(row_idx), int(n), src.shape[1], _p(out),
                                               int(out.dtype == torch.float32), _stream()))
    return out


def rows_scatter(src, row_idx, n, T, out):
    """out [T, 768] bf16 = 0 except out[row_idx[i]] = src[i] (see nbest_rows_scatter)."""
    ctx = _ctx(src)
    _check_bf16(src, "src")
    with _timed('rows_scatter', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_rows_scatter(ctx.handle, _p(src), _p(row_idx), int(n), int(T), src.shape[1], _p(out), _stream()))
    return out


def zero_(t):
    """Asynchronous zero-fill of a contiguous tensor on the bound stream (gradient / accumulator buffers)."""
    ctx = _ctx(t)
    ctx.check(_lib.lib().nbest_zero(ctx.handle, _p(t), t.numel() * t.element_size(), _stream()))
    return t


def rows_mark(tokens, T, flags):
    """flags[:] = 0; flags[tokens[t]] = 1 for t < T (int32 flags over the embedding rows; see nbest_rows_touched)."""
    ctx = _ctx(flags)
    ctx.check(_lib.lib().nbest_rows_touched(ctx.handle, _p(tokens), int(T), flags.numel(), -1, _p(flags), 0, None, None, _stream()))


def rows_compact(flags, skip_row, rows, count):
    """rows[0..count) = ascending indices of set flags, skip_row excluded; count is a 1-element int32 device tensor."""
    ctx = _ctx(flags)
    ctx.check(_lib.lib().nbest_rows_touched(ctx.handle, None, 0, flags.numel(), int(skip_row), _p(flags), 1, _p(rows), _p(count), _stream()))


def rows_move_f32(src, rows, n, dst, scatter):
    ctx = _ctx(src)
    ctx.check(_lib.lib().nbest_rows_move_f32(ctx.handle, _p(src), _p(rows), int(n), src.shape[1], _p(dst), int(bool(scatter)), _stream()))


def pack_hyp_ids(pk, sep_id, B=None):
    """uint8 [T] hypothesis index of every packed token (see nbest_pack_hyp_ids): 0 = [CLS] + system turn + first [SEP]."""
    B = pk.B if B is None else B
    out = torch.empty(max(pk.T, 1), dtype=torch.uint8, device=pk.tokens.device)
    ctx = _ctx(pk.tokens)
    with _timed('pack_hyp_ids', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_pack_hyp_ids(ctx.handle, _p(pk.tokens), _p(pk.cu_seqlens), B, int(sep_id), _p(out), _stream()))
    return out[:pk.T]


# ---------------------------------------------------------------------------------------------------------- embed / LN
def embed_ln_fwd(pk, word, posemb, type_emb, gamma, beta, eps, y, mean, rstd, p_drop=0.0, seed=0):
    ctx = _ctx(word)
    with _timed('embed_ln_fwd', 0.0, pk.T * (3 * 768 * 4 + 768 * 2.0)):
        ctx.check(_lib.lib().nbest_embed_ln_fwd(ctx.handle, _p(pk.tokens), _p(pk.seg), _p(pk.pos), pk.T, _p(word), _p(posemb),
                                                _p(type_emb), _p(gamma), _p(beta), float(eps), word.shape[1], _p(y), _p(mean),
                                                _p(rstd), float(p_drop), _seed(seed), _stream()))


def embed_ln_bwd(pk, word, posemb, type_emb, gamma, mean, rstd, dy, dword, dpos, dtype, dgamma, dbeta, p_drop=0.0, seed=0,
                 word_pad_row=-1, pos_pad_row=-1, T=None):
    ctx = _ctx(word)
    with _timed('embed_ln_bwd', 0.0, (pk.T if T is None else T) * (3 * 768 * 4 + 768 * 2 + 3 * 768 * 4.0)):
        ctx.check(_lib.lib().nbest_embed_ln_bwd(ctx.handle, _p(pk.tokens), _p(pk.seg), _p(pk.pos), pk.T if T is None else T,
                                                _p(word), _p(posemb), _p(type_emb), _p(gamma), _p(mean), _p(rstd), word.shape[1],
                                                _p(dy), float(p_drop), _seed(seed), _p(dword), _p(dpos), _p(dtype), _p(dgamma),
                                                _p(dbeta), int(word_pad_row), int(pos_pad_row), _stream()))


def ln_fwd(x, gamma, beta, eps, y, mean=None, rstd=None, T=None, row_partials=None):
    """row_partials: fp32 [T, n, 2] per-row partial {sum, sum of squares} written by the producing GEMM (EPI_BIAS_DROP_RES
    with out2=...): the LayerNorm then needs one pass over the row."""
    ctx = _ctx(x)
    n_part = 0
    if row_partials is not None:
        if row_partials.dtype != torch.float32 or not row_partials.is_contiguous() or row_partials.dim() != 3:
            raise ValueError("row_partials must be a contiguous fp32 [T, n, 2] tensor")
        n_part = row_partials.shape[1]
    with _timed('ln_fwd', 0.0, (x.shape[0] if T is None else T) * (2 * 768 * 2.0)):
        ctx.check(_lib.lib().nbest_ln_fwd_stats(ctx.handle, _p(x), _p(gamma), _p(beta), float(eps), x.shape[0] if T is None else T,
                                                x.shape[1], _p(row_partials), n_part, _p(y), _p(mean), _p(rstd), _stream()))


def ln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, dx_masked=None, dbias=None, p_drop=0.0, seed=0, T=None):
    ctx = _ctx(x)
    with _timed('ln_bwd', 0.0, (x.shape[0] if T is None else T) * ((3 + (dx_masked is not None)) * 768 * 2.0)):
        ctx.check(_lib.lib().nbest_ln_bwd(ctx.handle, _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma),
                                          x.shape[0] if T is None else T, x.shape[1], _p(dx), _p(dx_masked), float(p_drop),
                                          _seed(seed), _p(dgamma), _p(dbeta), _p(dbias), _stream()))


def colsum(x, out, T=None):
    ctx = _ctx(x)
    with _timed('colsum_bf16', 0.0, (x.shape[0] if T is None else T) * x.shape[1] * 2.0):
        ctx.check(_lib.lib().nbest_colsum_bf16(ctx.handle, _p(x), x.shape[0] if T is None else T, x.shape[1], _p(out), _stream()))


def cast_f32_bf16(src, dst):
    ctx = _ctx(src)
    with _timed('cast_f32_bf16', 0.0, src.numel() * 6.0):
        ctx.check(_lib.lib().nbest_cast_f32_bf16(ctx.handle, _p(src), _p(dst), src.numel(), _stream()))


# ---------------------------------------------------------------------------------------------------------- attention
def _attn_cost(T, sum_l2, heads, bwd=False, T_active=None):
    """Algorithmic FLOPs / bytes of one attention call: forward 4 * L^2 * 64 per (sequence, head), backward
    twice that; bytes = QKV rows read + context rows written (+ dO read, dQKV written in the backward) + lse / delta."""
    hd = heads * 64
    Ta = T if T_active is None else T_active
    fl = 4.0 * sum_l2 * hd * (2.0 if bwd else 1.0)
    by = (Ta * (3 * hd + hd + 3 * hd) * 2.0 + 2 * heads * Ta * 4.0) if bwd else (T * (3 * hd + hd) * 2.0 + heads * T * 4.0)
    return fl, by


def attn_fwd(qkv, cu_seqlens, key_valid, B, max_len, heads, T, out, lse, p_drop=0.0, seed=0, min_len=0, sum_l2=0.0):
    """min_len > 0: only sequences of at least that many tokens (the rest belongs to attn_tiles_fwd)."""
    ctx = _ctx(qkv)
    fl, by = _attn_cost(T, sum_l2, heads) if (sum_l2 and not min_len) else (0.0, 0.0)
    with _timed('attn_varlen_fwd', fl, by):
        ctx.check(_lib.lib().nbest_attn_varlen_fwd2(ctx.handle, _p(qkv), _p(cu_seqlens), _p(key_valid), B, max_len, heads, T,
                                                    _p(out), _p(lse), float(p_drop), _seed(seed), int(min_len), _stream()))


def attn_bwd(qkv, cu_seqlens, key_valid, B, max_len, heads, T, out, dout, lse, dqkv, delta_ws, p_drop=0.0, seed=0,
             T_active=None, min_len=0, sum_l2=0.0):
    ctx = _ctx(qkv)
    fl, by = _attn_cost(T, sum_l2, heads, True, T_active) if (sum_l2 and not min_len) else (0.0, 0.0)
    with _timed('attn_varlen_bwd', fl, by):
        ctx.check(_lib.lib().nbest_attn_varlen_bwd2(ctx.handle, _p(qkv), _p(cu_seqlens), _p(key_valid), B, max_len, heads, T,
                                                    T if T_active is None else T_active, _p(out), _p(dout), _p(lse), _p(dqkv), _p(delta_ws), float(p_drop), _seed(seed),
                                                    int(min_len), _stream()))


class AttnPlan:
    """Tile plan of a packed batch for the tcgen05 attention kernels (see nbest_attn_plan)."""
    __slots__ = ("tiles", "counts", "row_bounds", "max_tiles")


def attn_plan(cu_seqlens, seq_of, B, T, break_at=None):
    """Greedy packing of the batch's sequences (<= 128 tokens each) into 128-row tiles + per-token sequence bounds.
    break_at: no tile straddles this sequence index (the ASR / transcript boundary of the merged batch)."""
    dev = cu_seqlens.device
    pl = AttnPlan()
    pl.tiles = torch.empty(2 * B, dtype=torch.int32, device=dev)
    pl.counts = torch.empty(4, dtype=torch.int32, device=dev)
    pl.row_bounds = torch.empty(2 * max(T, 1), dtype=torch.int32, device=dev)
    pl.max_tiles = B
    ctx = _ctx(cu_seqlens)
    with _timed('attn_plan', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_attn_plan(ctx.handle, _p(cu_seqlens), _p(seq_of), B, T, B if break_at is None else break_at,
                                             _p(pl.tiles), _p(pl.counts), _p(pl.row_bounds), _stream()))
    return pl


def attn_tiles_fwd(qkv, plan, count_idx, key_valid, heads, T, out, lse, p_drop=0.0, seed=0, sum_l2=0.0):
    """tcgen05 / TMEM / TMA attention forward over the plan's tiles (sequences of <= 128 tokens)."""
    ctx = _ctx(qkv)
    _check_bf16(qkv, "qkv")
    fl, by = _attn_cost(T, sum_l2, heads) if sum_l2 else (0.0, 0.0)
    with _timed('attn_tiles_fwd', fl, by):
        ctx.check(_lib.lib().nbest_attn_tiles_fwd(ctx.handle, _p(qkv), _p(plan.tiles), _p(plan.counts), int(count_idx),
                                                  plan.max_tiles, _p(plan.row_bounds), _p(key_valid), heads, T, _p(out), _p(lse),
                                                  float(p_drop), _seed(seed), _stream()))


def attn_tiles_bwd(qkv, plan, count_idx, key_valid, heads, T, T_active, dout, lse, delta, delta_pitch, dqkv, p_drop=0.0, seed=0,
                   sum_l2=0.0):
    ctx = _ctx(qkv)
    _check_bf16(qkv, "qkv")
    _check_bf16(dout, "dout")
    fl, by = _attn_cost(T, sum_l2, heads, True, T_active) if sum_l2 else (0.0, 0.0)
    with _timed('attn_tiles_bwd', fl, by):
        ctx.check(_lib.lib().nbest_attn_tiles_bwd(ctx.handle, _p(qkv), _p(plan.tiles), _p(plan.counts), int(count_idx),
                                                  plan.max_tiles, _p(plan.row_bounds), _p(key_valid), heads, T, T_active, _p(dout),
                                                  _p(lse), _p(delta), int(delta_pitch), _p(dqkv), float(p_drop), _seed(seed),
                                                  _stream()))


def attn_cls_fwd(qkv, cu_seqlens, key_valid, B, max_len, heads, T, out_cls, lse_cls, p_drop=0.0, seed=0):
    """Last-layer attention for the CLS query row of each of the B sequences (see nbest_attn_cls_fwd)."""
    ctx = _ctx(qkv)
    with _timed('attn_cls_fwd', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_attn_cls_fwd(ctx.handle, _p(qkv), _p(cu_seqlens), _p(key_valid), B, max_len, heads, T,
                                                _p(out_cls), _p(lse_cls), float(p_drop), _seed(seed), _stream()))


def attn_cls_bwd(qkv, cu_seqlens, key_valid, B, max_len, heads, T, out_cls, dout_cls, lse_cls, lse_stride, dqkv, p_drop=0.0,
                 seed=0):
    ctx = _ctx(qkv)
    with _timed('attn_cls_bwd', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_attn_cls_bwd(ctx.handle, _p(qkv), _p(cu_seqlens), _p(key_valid), B, max_len, heads, T,
                                                _p(out_cls), _p(dout_cls), _p(lse_cls), lse_stride, _p(dqkv), float(p_drop),
                                                _seed(seed), _stream()))


# ---------------------------------------------------------------------------------------------------------- STC head / loss
class DeviceHierarchy:
    """Label hierarchy tables on the device + the ctypes struct the C ABI takes (nbest_hierarchy)."""

    def __init__(self, top2bottom, none_bottoms=(), device="cuda"):
        t2b = {int(k): [int(x) for x in v] for k, v in top2bottom.items()}
        self.top2bottom = t2b
        self.n_top = len(t2b)
        self.n_bottom = sum(len(v) for v in t2b.values())
        self.group_tops = [k for k in sorted(t2b) if len(t2b[k]) >= 2]
        self.n_groups = len(self.group_tops)
        grp_off = [self.n_top]
        for k in self.group_tops:
            grp_off.append(grp_off[-1] + len(t2b[k]))
        self.n_cols = grp_off[-1]
        col_group = [0] * self.n_top
        col_bottom = [t2b[i][0] if len(t2b[i]) == 1 else -1 for i in range(self.n_top)]
        for g, k in enumerate(self.group_tops, start=1):
            col_group += [g] * len(t2b[k])
            col_bottom += t2b[k]
        none_set = set(int(x) for x in none_bottoms)
        self.none_bottoms = none_set
        none_col = [1 if (c >= self.n_top and col_bottom[c] in none_set) else 0 for c in range(self.n_cols)]
        self.grp_off_host, self.col_bottom_host = grp_off, col_bottom
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=device)
        self.col_group = i32(col_group)
        self.col_bottom = i32(col_bottom)
        self.grp_off = i32(grp_off)
        self.grp_top = i32(self.group_tops if self.group_tops else [0])
        self.none_col = torch.tensor(none_col, dtype=torch.uint8, device=device)
        self.struct = Hierarchy(self.n_top, self.n_bottom, self.n_groups, self.n_cols, self.col_group.data_ptr(),
                                self.col_bottom.data_ptr(), self.grp_off.data_ptr(), self.grp_top.data_ptr())

    def ref(self):
        return C.byref(self.struct)


def stc_head_fwd(x, cu_seqlens, B, W, bias, hier, cls, logits, top, bottom, final, decode=None, p_drop=0.0, seed=0):
    ctx = _ctx(x)
    with _timed('stc_head_fwd', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_stc_head_fwd(ctx.handle, _p(x), _p(cu_seqlens), B, x.shape[1], _p(W), _p(bias), hier.ref(),
                                                _p(hier.none_col), float(p_drop), _seed(seed), _p(cls), _p(logits), _p(top),
                                                _p(bottom), _p(final), _p(decode), _stream()))


def stc_loss_fwd_bwd(logits, labels, hier, losses, dlogits, asr_cls=None, trans_cls=None, mse_scale=1.0, d_asr=None,
                     d_trans=None):
    ctx = _ctx(logits)
    with _timed('stc_loss_fwd_bwd', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_stc_loss_fwd_bwd(ctx.handle, _p(logits), _p(labels), logits.shape[0], hier.ref(), _p(asr_cls),
                                                    _p(trans_cls), asr_cls.shape[1] if asr_cls is not None else 768,
                                                    float(mse_scale), _p(losses), _p(dlogits), _p(d_asr), _p(d_trans), _stream()))


def stc_scores_bwd(top, bottom, d_top, d_bottom, d_final, hier, dlogits):
    ctx = _ctx(top)
    with _timed('stc_scores_bwd', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_stc_scores_bwd(ctx.handle, _p(top), _p(bottom), _p(d_top), _p(d_bottom), _p(d_final),
                                                  top.shape[0], hier.ref(), _p(dlogits), _stream()))


def stc_head_bwd(dlogits, cls, W, hier, dW, dbias, dcls, accumulate_dcls=False, p_drop=0.0, seed=0):
    ctx = _ctx(cls)
    with _timed('stc_head_bwd', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_stc_head_bwd(ctx.handle, _p(dlogits), _p(cls), _p(W), cls.shape[0], cls.shape[1], hier.ref(),
                                                float(p_drop), _seed(seed), _p(dW), _p(dbias), _p(dcls), int(accumulate_dcls),
                                                _stream()))


def stc_metrics(decode, labels, counters, col_mask=None):
    """counters[4] (int64, device) += TP, FP, FN, exact matches of this batch (see nbest_stc_metrics)."""
    ctx = _ctx(decode)
    assert decode.dtype == torch.uint8 and labels.dtype == torch.float32 and counters.dtype == torch.int64
    assert decode.is_contiguous() and labels.is_contiguous() and decode.shape == labels.shape
    with _timed('stc_metrics', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_stc_metrics(ctx.handle, _p(decode), _p(labels), _p(col_mask), decode.shape[0],
                                               decode.shape[1], _p(counters), _stream()))


def cls_scatter(dcls, cu_seqlens, B, T, dx):
    ctx = _ctx(dcls)
    with _timed('cls_scatter', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_cls_scatter(ctx.handle, _p(dcls), _p(cu_seqlens), B, T, dcls.shape[1], _p(dx), _stream()))


# ---------------------------------------------------------------------------------------------------------- BertAdam
def bertadam_step(p, g, m, v, p_bf16, tensors_dev, n_tensors, chunks_dev, n_chunks, norms_ws, sched, b1=0.9, b2=0.999,
                  eps=1e-6, max_grad_norm=1.0, mode=_lib.ADAM_BERT, global_clip=False, step=0):
    """mode ADAM_BERT: nbest_bertadam_step semantics; ADAM_HF_ADAMW / ADAM_TORCH: see nbest_adam_step."""
    ctx = _ctx(p)
    with _timed('bertadam_step', 0.0, 0.0):
        ctx.check(_lib.lib().nbest_adam_step(ctx.handle, int(mode), _p(p), _p(g), _p(m), _p(v), _p(p_bf16), _p(tensors_dev),
                                             n_tensors, _p(chunks_dev), n_chunks, _p(norms_ws), float(sched), float(b1),
                                             float(b2), float(eps), float(max_grad_norm), int(bool(global_clip)), int(step),
                                             _stream()))
