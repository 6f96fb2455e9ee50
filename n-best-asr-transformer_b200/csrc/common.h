// Internal host-side definitions shared by the translation units behind the C ABI (include/nbest_sm100.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nbest_sm100.h"

typedef CUresult (*nbest_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Environment knobs, read ONCE at nbest_ctx_create (a getenv per GEMM launch showed up in the host profile).
struct nbest_knobs {
  int gemm_cta_group;   // NBEST_GEMM_CTA_GROUP: 1 = single-CTA tiles, else CTA pairs
  int gemm_force_bn;    // NBEST_GEMM_BN: 128 / 256, 0 = automatic
  int wgrad_splits;     // NBEST_WGRAD_SPLITS: >= 1 forces the split count, 0 = automatic
  int gemm_debug;       // NBEST_GEMM_DEBUG
  int gemm_stages;      // NBEST_GEMM_STAGES: cap on the smem ring depth, 0 = full
  int attn_no_fused_bwd;   // NBEST_ATTN_NO_FUSED_BWD
};

// A tensor map is a pure function of (base, rows, cols, pitch, box): the caching allocator hands the same buffers back
// every step, so encoded maps are kept in a small direct-mapped cache instead of calling the driver 3-4 times per GEMM.
struct nbest_tmap_entry {
  const void* base;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  uint32_t valid;
  CUtensorMap map;
};
constexpr int kTmapCacheSize = 1024;
constexpr int kSchedRing = 64;        // launches in flight never share a scheduler slot

// Per-step scalars a replayed CUDA graph of the training step reads from device memory (ptx.cuh step_salt; bertadam.cu).
struct nbest_step_state {
  uint32_t salt;        // XORed into every dropout seed
  uint32_t indirect;    // (host mirror only) != 0: the optimizer kernels take sched / bias corrections from here
  double sched;         // learning-rate schedule multiplier of this step
  float inv_bc1, inv_sqrt_bc2;   // torch.optim.Adam bias corrections of this step
};

struct nbest_ctx {
  int device;
  int num_sms;
  int cc_major, cc_minor;
  nbest_encode_tiled_fn encode_tiled;  // resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency)
  uint64_t launches;                   // kernels launched through this context (bench.py reports it)
  nbest_knobs knobs;
  int reserve_sms;                     // SMs the persistent GEMMs leave free (nbest_ctx_set_sm_reserve)
  int gemm_dynamic;                    // GEMM work items drawn from a global counter (nbest_ctx_set_gemm_dynamic)
  uint32_t* sched_buf;                 // device: kSchedRing x {next item, finished units}; the ONE allocation the library owns
  uint32_t sched_seq;
  nbest_step_state* step_state;        // device: per-step scalars (nbest_ctx_set_step_state); zero = eager behaviour
  int step_indirect;                   // optimizer scalars come from step_state (nbest_ctx_set_step_indirect)
  uint64_t tmap_hits, tmap_misses;
  nbest_tmap_entry* tmap_cache;
  char err[512];
};

void nbest_set_error(nbest_ctx* ctx, const char* fmt, ...);
// device address of the dropout salt (every kernel that draws dropout masks takes it next to its by-value seed)
static inline const uint32_t* nbest_salt(const nbest_ctx* ctx) { return &ctx->step_state->salt; }

#define NBEST_CHECK_ARG(ctx, cond, msg)                                        \
  do {                                                                         \
    if (!(cond)) {                                                             \
      nbest_set_error((ctx), "%s:%d: invalid argument: %s", __func__, __LINE__, (msg)); \
      return NBEST_EINVAL;                                                     \
    }                                                                          \
  } while (0)

#define NBEST_CHECK_CUDA(ctx, expr)                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      nbest_set_error((ctx), "%s:%d: CUDA error %s (%s)", __func__, __LINE__, cudaGetErrorName(e__), \
                      cudaGetErrorString(e__));                                                      \
      return NBEST_ECUDA;                                                                            \
    }                                                                                                \
  } while (0)

#define NBEST_CHECK_LAUNCH(ctx)       \
  do {                                \
    (ctx)->launches++;                \
    NBEST_CHECK_CUDA((ctx), cudaGetLastError()); \
  } while (0)

// Kernel launch with programmatic dependent launch (NBEST_PDL=0 disables): the ~250 kernels of a step are short (5-100 us),
// so the launch latency and the drain / fill bubble at every kernel boundary add up; see ptx.cuh pdl_grid_sync.
extern int g_nbest_pdl;
#ifdef __CUDACC__
template <typename K, typename... Args>
inline cudaError_t nbest_launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_nbest_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

// Encode a 2-D bf16 row-major tensor [rows, cols] (row pitch ld elements) with a {64 x box_rows} box, 128-B swizzle.
int nbest_make_tmap_bf16(nbest_ctx* ctx, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows);
