// Context, error reporting and TMA descriptor encoding for libnbest_sm100.so.
#include <stdarg.h>
#include <stdlib.h>

#include "common.h"

int g_nbest_pdl = 1;

void nbest_set_error(nbest_ctx* ctx, const char* fmt, ...) {
  if (!ctx) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
  va_end(ap);
}

extern "C" int nbest_abi_version(void) { return NBEST_ABI_VERSION; }

static __thread char g_create_err[256];

extern "C" int nbest_ctx_create(nbest_ctx** out, int device) {
  if (!out) return NBEST_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    snprintf(g_create_err, sizeof(g_create_err), "no CUDA device: %s", cudaGetErrorString(e));
    return NBEST_ECUDA;  // no CPU fallback: the hot path exists only on the GPU
  }
  if (device < 0 || device >= ndev) return NBEST_EINVAL;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NBEST_ECUDA;
  if (prop.major != 10) {
    snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d, this library is sm_100a only", device, prop.major,
             prop.minor);
    return NBEST_EUNSUPPORTED;
  }
  nbest_ctx* ctx = (nbest_ctx*)calloc(1, sizeof(nbest_ctx));
  if (!ctx) return NBEST_EINVAL;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  if (cudaSetDevice(device) != cudaSuccess) {
    free(ctx);
    return NBEST_ECUDA;
  }
  cudaFree(0);  // make sure the primary context exists before asking for driver entry points
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    snprintf(g_create_err, sizeof(g_create_err), "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    free(ctx);
    return NBEST_ECUDA;
  }
  ctx->encode_tiled = (nbest_encode_tiled_fn)fn;
  auto env_int = [](const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
  };
  g_nbest_pdl = env_int("NBEST_PDL", 1) != 0;
  ctx->knobs.gemm_cta_group = env_int("NBEST_GEMM_CTA_GROUP", 2) == 1 ? 1 : 2;
  ctx->knobs.gemm_force_bn = env_int("NBEST_GEMM_BN", 0);
  ctx->knobs.wgrad_splits = env_int("NBEST_WGRAD_SPLITS", 0);
  ctx->knobs.gemm_debug = env_int("NBEST_GEMM_DEBUG", 0);
  ctx->knobs.gemm_stages = env_int("NBEST_GEMM_STAGES", 0);
  ctx->knobs.attn_no_fused_bwd = getenv("NBEST_ATTN_NO_FUSED_BWD") != nullptr;
  ctx->reserve_sms = env_int("NBEST_GEMM_RESERVE_SMS", 0);
  ctx->tmap_cache = (nbest_tmap_entry*)calloc(kTmapCacheSize, sizeof(nbest_tmap_entry));
  if (!ctx->tmap_cache) {
    free(ctx);
    return NBEST_EINVAL;
  }
  ctx->gemm_dynamic = env_int("NBEST_GEMM_DYNAMIC", 0) != 0;
  // one device allocation: the GEMM scheduler ring, then the per-step state record (salt 0 = eager behaviour)
  const size_t dev_bytes = kSchedRing * 2 * sizeof(uint32_t) + sizeof(nbest_step_state);
  if (cudaMalloc(&ctx->sched_buf, dev_bytes) != cudaSuccess || cudaMemset(ctx->sched_buf, 0, dev_bytes) != cudaSuccess) {
    snprintf(g_create_err, sizeof(g_create_err), "cannot allocate the %zu-byte GEMM scheduler buffer", kSchedRing * 2 * sizeof(uint32_t));
    free(ctx->tmap_cache);
    free(ctx);
    return NBEST_ECUDA;
  }
  ctx->step_state = reinterpret_cast<nbest_step_state*>(ctx->sched_buf + kSchedRing * 2);
  ctx->step_indirect = 0;
  ctx->err[0] = 0;
  *out = ctx;
  return NBEST_OK;
}

extern "C" void nbest_ctx_destroy(nbest_ctx* ctx) {
  if (ctx) {
    free(ctx->tmap_cache);
    cudaFree(ctx->sched_buf);
  }
  free(ctx);
}

extern "C" int nbest_ctx_set_gemm_dynamic(nbest_ctx* ctx, int on) {
  if (!ctx) return NBEST_EINVAL;
  ctx->gemm_dynamic = on != 0;
  return NBEST_OK;
}

extern "C" int nbest_ctx_set_sm_reserve(nbest_ctx* ctx, int n_sms) {
  if (!ctx || n_sms < 0 || n_sms > ctx->num_sms - 2) return NBEST_EINVAL;
  ctx->reserve_sms = n_sms & ~1;      // CTA pairs: whole TPCs
  return NBEST_OK;
}

namespace {
__global__ void set_step_state_kernel(nbest_step_state* st, uint32_t salt, double sched, float inv_bc1, float inv_sqrt_bc2) {
  st->salt = salt;
  st->indirect = 1;
  st->sched = sched;
  st->inv_bc1 = inv_bc1;
  st->inv_sqrt_bc2 = inv_sqrt_bc2;
}
}  // namespace

extern "C" int nbest_ctx_set_step_state(nbest_ctx* ctx, uint32_t salt, double sched, float inv_bc1, float inv_sqrt_bc2,
                                        void* stream) {
  if (!ctx) return NBEST_EINVAL;
  // a plain launch (no programmatic overlap): the kernels behind it read the record after their griddepcontrol.wait
  set_step_state_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ctx->step_state, salt, sched, inv_bc1, inv_sqrt_bc2);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_ctx_set_step_indirect(nbest_ctx* ctx, int on) {
  if (!ctx) return NBEST_EINVAL;
  ctx->step_indirect = on != 0;
  return NBEST_OK;
}

extern "C" uint64_t nbest_tmap_cache_hits(nbest_ctx* ctx) { return ctx ? ctx->tmap_hits : 0; }

extern "C" const char* nbest_last_error(nbest_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" uint64_t nbest_launch_count(nbest_ctx* ctx) { return ctx ? ctx->launches : 0; }

int nbest_make_tmap_bf16(nbest_ctx* ctx, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};  // bytes, dimension 1
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (gstride[0] & 15) != 0) {
    nbest_set_error(ctx, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    return NBEST_EINVAL;
  }
  uint64_t h = reinterpret_cast<uintptr_t>(base) * 0x9E3779B97F4A7C15ull;
  h ^= (rows + 0x632BE59BD9B4E019ull) * 0xC2B2AE3D27D4EB4Full;
  h ^= (cols * 0x165667B19E3779F9ull) ^ (ld * 0x27D4EB2F165667C5ull) ^ ((uint64_t)box_rows << 48);
  nbest_tmap_entry* e = &ctx->tmap_cache[(h ^ (h >> 29)) & (kTmapCacheSize - 1)];
  if (e->valid && e->base == base && e->rows == rows && e->cols == cols && e->ld == ld && e->box_rows == box_rows) {
    *out = e->map;
    ++ctx->tmap_hits;
    return NBEST_OK;
  }
  ++ctx->tmap_misses;
  CUresult r = ctx->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    nbest_set_error(ctx, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", (int)r,
                    (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return NBEST_ECUDA;
  }
  e->base = base;
  e->rows = rows;
  e->cols = cols;
  e->ld = ld;
  e->box_rows = box_rows;
  e->map = *out;
  e->valid = 1;
  return NBEST_OK;
}
