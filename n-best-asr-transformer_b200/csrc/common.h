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

struct nbest_ctx {
  int device;
  int num_sms;
  int cc_major, cc_minor;
  nbest_encode_tiled_fn encode_tiled;  // resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency)
  uint64_t launches;                   // kernels launched through this context (bench.py reports it)
  char err[512];
};

void nbest_set_error(nbest_ctx* ctx, const char* fmt, ...);

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

// Encode a 2-D bf16 row-major tensor [rows, cols] (row pitch ld elements) with a {64 x box_rows} box, 128-B swizzle.
int nbest_make_tmap_bf16(nbest_ctx* ctx, CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows);
