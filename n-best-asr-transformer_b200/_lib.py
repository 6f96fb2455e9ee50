"""ctypes binding of libnbest_sm100.so (include/nbest_sm100.h).

The product path has no CPU fallback: `lib()` raises if the shared library is missing and `Context()` raises if no
sm_100 GPU is present. Tensors cross the boundary as raw device pointers (`tensor.data_ptr()`), never as torch types.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbest_sm100.so")

NBEST_OK = 0
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_DGELU, EPI_ADD, EPI_ACCUM_F32, EPI_DELTA = range(8)
ADAM_BERT, ADAM_HF_ADAMW, ADAM_TORCH = range(3)

_vp, _i32, _i64, _u32, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_double


class Hierarchy(C.Structure):
    _fields_ = [("n_top", _i32), ("n_bottom", _i32), ("n_groups", _i32), ("n_cols", _i32),
                ("col_group", _vp), ("col_bottom", _vp), ("grp_off", _vp), ("grp_top", _vp)]


class AdamTensor(C.Structure):
    _fields_ = [("offset", _i64), ("numel", _i64), ("lr", _f64), ("weight_decay", _f32), ("active", _i32)]


_SIGNATURES = {
    "nbest_abi_version": (C.c_int, []),
    "nbest_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "nbest_ctx_destroy": (None, [_vp]),
    "nbest_last_error": (C.c_char_p, [_vp]),
    "nbest_launch_count": (_u64, [_vp]),
    "nbest_ctx_set_sm_reserve": (C.c_int, [_vp, C.c_int]),
    "nbest_tmap_cache_hits": (_u64, [_vp]),
    "nbest_ctx_set_gemm_dynamic": (C.c_int, [_vp, C.c_int]),
    "nbest_ctx_set_step_state": (C.c_int, [_vp, _u32, _f64, _f32, _f32, _vp]),
    "nbest_ctx_set_step_indirect": (C.c_int, [_vp, C.c_int]),
    "nbest_pack_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nbest_pack_hyp_ids": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "nbest_pack_batch_dual": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp]),
    "nbest_rows_gather": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "nbest_rows_scatter": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "nbest_zero": (C.c_int, [_vp, _vp, _i64, _vp]),
    "nbest_rows_touched": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "nbest_rows_move_f32": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "nbest_embed_ln_fwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _f32, C.c_int, _vp, _vp, _vp,
                                     _f32, _u32, _vp]),
    "nbest_embed_ln_bwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _f32, _u32,
                                     _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "nbest_ln_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _f32, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nbest_ln_fwd_stats": (C.c_int, [_vp, _vp, _vp, _vp, _f32, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "nbest_ln_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _f32, _u32, _vp, _vp, _vp, _vp]),
    "nbest_colsum_bf16": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "nbest_cast_f32_bf16": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "nbest_gemm_bf16": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _i64, C.c_int, _vp, _i64, C.c_int, C.c_int, C.c_int,
                                  C.c_int, _vp, _vp, _i64, _vp, _f32, _u32, _vp]),
    "nbest_attn_varlen_fwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _f32, _u32, _vp]),
    "nbest_attn_varlen_bwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                        _f32, _u32, _vp]),
    "nbest_attn_plan": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nbest_attn_tiles_fwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _f32, _u32, _vp]),
    "nbest_attn_tiles_bwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp,
                                       C.c_int, _vp, _f32, _u32, _vp]),
    "nbest_attn_varlen_fwd2": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _f32, _u32, C.c_int, _vp]),
    "nbest_attn_varlen_bwd2": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                         _f32, _u32, C.c_int, _vp]),
    "nbest_attn_cls_fwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _f32, _u32, _vp]),
    "nbest_attn_cls_bwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp,
                                     _f32, _u32, _vp]),
    "nbest_stc_head_fwd": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.POINTER(Hierarchy), _vp, _f32, _u32,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nbest_stc_loss_fwd_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int, C.POINTER(Hierarchy), _vp, _vp, C.c_int, _f32, _vp, _vp,
                                         _vp, _vp, _vp]),
    "nbest_stc_scores_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(Hierarchy), _vp, _vp]),
    "nbest_stc_head_bwd": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(Hierarchy), _f32, _u32, _vp, _vp,
                                     _vp, C.c_int, _vp]),
    "nbest_cls_scatter": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "nbest_stc_metrics": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "nbest_bertadam_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _f64, _f32, _f32,
                                      _f32, _f32, _vp]),
    "nbest_adam_step": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _f64, _f32, _f32,
                                  _f32, _f32, C.c_int, C.c_int, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Load the shared library (once). Raises if it has not been built — there is no fallback implementation."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libnbest_sm100.so is missing (%s). Build it with `python n-best-asr-transformer_b200/build.py`; "
                    "this package has no CPU or PyTorch fallback path." % LIB_PATH)
            l = C.CDLL(LIB_PATH)
            missing = []
            for name, (res, args) in _SIGNATURES.items():
                try:
                    fn = getattr(l, name)
                except AttributeError:
                    missing.append(name)
                    continue
                fn.restype = res
                fn.argtypes = args
            if missing and not os.environ.get("NBEST_ALLOW_MISSING_SYMBOLS"):
                raise RuntimeError("libnbest_sm100.so does not export: %s (stale build?)" % ", ".join(missing))
            _lib = l
    return _lib


def exported_symbols():
    return list(_SIGNATURES.keys())


class NbestError(RuntimeError):
    pass


class Context:
    """One nbest_ctx per process / GPU rank."""

    def __init__(self, device=0):
        self._l = lib()
        h = _vp()
        rc = self._l.nbest_ctx_create(C.byref(h), int(device))
        if rc != NBEST_OK:
            msg = self._l.nbest_last_error(None)
            raise NbestError("nbest_ctx_create(device=%d) failed (%d): %s — an sm_100 GPU is required, there is no "
                             "CPU fallback" % (device, rc, msg.decode() if msg else ""))
        self.handle = h
        self.device = device

    def check(self, rc):
        if rc != NBEST_OK:
            msg = self._l.nbest_last_error(self.handle)
            raise NbestError("libnbest_sm100 call failed (%d): %s" % (rc, msg.decode() if msg else ""))

    def launches(self):
        return int(self._l.nbest_launch_count(self.handle))

    def set_sm_reserve(self, n_sms):
        self.check(self._l.nbest_ctx_set_sm_reserve(self.handle, int(n_sms)))

    def set_gemm_dynamic(self, on):
        self.check(self._l.nbest_ctx_set_gemm_dynamic(self.handle, int(bool(on))))

    def set_step_state(self, salt, sched=1.0, inv_bc1=1.0, inv_sqrt_bc2=1.0, stream=None):
        """Per-step scalars a replayed CUDA graph of the training step reads from device memory (include/nbest_sm100.h)."""
        self.check(self._l.nbest_ctx_set_step_state(self.handle, int(salt) & 0xFFFFFFFF, float(sched), float(inv_bc1),
                                                    float(inv_sqrt_bc2), stream))

    def set_step_indirect(self, on):
        self.check(self._l.nbest_ctx_set_step_indirect(self.handle, int(bool(on))))

    def tmap_cache_hits(self):
        return int(self._l.nbest_tmap_cache_hits(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._l.nbest_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_ctx_by_device = {}


def context(device=0):
    c = _ctx_by_device.get(device)
    if c is None:
        c = Context(device)
        _ctx_by_device[device] = c
    return c
