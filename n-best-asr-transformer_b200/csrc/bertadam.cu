// Fused multi-tensor BertAdam over one flat fp32 parameter buffer.
//
// Reference: BertAdam.step (models/optimization.py:237-302) run over 221 one-tensor param groups
// (n_best_asr_bert.py:535-550): per tensor clip_grad_norm_(p, 1.0) (:270-271), m/v moments (:275-276),
// update = m / (sqrt(v) + e) + wd * p (:277-287), p -= lr * schedule(step/t_total) * update (:289-293), no bias
// correction. The reference launches ~14 kernels per tensor (~3 k per step); here it is two launches per step:
//   pass 1: per-tensor sum of squares of the gradient                       (reads g:           4 B / param)
//   pass 2: clip + moments + decay + update (+ bf16 working copy refresh)   (reads p,g,m,v, writes p,m,v[,bf16]: 28-30 B)
// The norm is reduced in a FIXED order (per-chunk partials, then one block per tensor folds its partials): data-parallel
// replicas that received bit-identical all-reduced gradients compute bit-identical clip coefficients and stay
// bit-identical (an atomicAdd reduction made them drift apart by ~1e-7 per step; tests/dp_parity_2gpu.py).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

using namespace nbest;

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
adam_norm_kernel(const float* __restrict__ g, const nbest_adam_tensor* __restrict__ tensors, const int32_t* __restrict__ chunks,
                 float* __restrict__ partials) {
  pdl_grid_sync();
  const int ti = chunks[3 * blockIdx.x], start = chunks[3 * blockIdx.x + 1], len = chunks[3 * blockIdx.x + 2];
  const int64_t base = tensors[ti].offset + start;
  const float* gp = g + base;
  float acc = 0.f;
  const int n4 = (base & 3) == 0 ? (len >> 2) : 0;   // float4 path only for 16-byte aligned chunks (tiny head biases are not)
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(gp) + i);
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  for (int i = (n4 << 2) + threadIdx.x; i < len; i += kThreads) acc += gp[i] * gp[i];
  acc = warp_sum(acc);
  __shared__ float sh[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;      // fixed-order reduction: no atomics
  }
}

// norms[ti] = sum of the chunk partials of tensor ti, folded in a fixed order (one block per tensor; chunks are sorted by
// tensor, so the tensor's range is found by bisection)
__global__ void __launch_bounds__(kThreads)
adam_norm_finish_kernel(const int32_t* __restrict__ chunks, int n_chunks, const float* __restrict__ partials,
                        float* __restrict__ norms) {
  pdl_grid_sync();
  const int ti = blockIdx.x;
  int lo = 0, hi = n_chunks;
  while (lo < hi) {                      // first chunk with tensor index >= ti
    const int mid = (lo + hi) >> 1;
    if (chunks[3 * mid] < ti) lo = mid + 1; else hi = mid;
  }
  const int first = lo;
  hi = n_chunks;
  while (lo < hi) {                      // first chunk with tensor index > ti
    const int mid = (lo + hi) >> 1;
    if (chunks[3 * mid] <= ti) lo = mid + 1; else hi = mid;
  }
  const int last = lo;
  float acc = 0.f;
  for (int i = first + threadIdx.x; i < last; i += kThreads) acc += partials[i];
  acc = warp_sum(acc);
  __shared__ float sh[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) norms[ti] = v;
  }
}

// One element of the three optimizers n_best_asr_bert.py:553-569 can select (MODE = nbest_adam_mode):
//   BERTADAM  models/optimization.py:275-293     u = m/(sqrt(v)+e) + wd p;  p -= lr_t u            (no bias correction)
//   ADAMW_HF  transformers 2.3.0 optimization.py AdamW.step with correct_bias=False (n_best_asr_bert.py:563):
//             p -= lr_t m/(sqrt(v)+e);  then  p -= lr_t wd p   (decoupled decay applied to the UPDATED parameter)
//   ADAM      torch.optim.Adam (n_best_asr_bert.py:554): g += wd p (L2);  p -= (lr_t/bc1) m / (sqrt(v)/sqrt(bc2) + e)
template <int MODE>
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float coef, float b1, float b2, float eps,
                                          float wd, float lr_t, float inv_bc1, float inv_sqrt_bc2) {
  g *= coef;
  if (MODE == NBEST_ADAM_TORCH && wd > 0.f) g += wd * p;
  m = m * b1 + (1.0f - b1) * g;
  v = v * b2 + (1.0f - b2) * g * g;
  if (MODE == NBEST_ADAM_BERT) {
    float upd = m / (sqrtf(v) + eps);
    if (wd > 0.f) upd += wd * p;
    p -= lr_t * upd;
  } else if (MODE == NBEST_ADAM_HF_ADAMW) {
    p -= lr_t * (m / (sqrtf(v) + eps));
    if (wd > 0.f) p -= lr_t * wd * p;
  } else {
    p -= (lr_t * inv_bc1) * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
adam_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   __nv_bfloat16* __restrict__ pb, const nbest_adam_tensor* __restrict__ tensors, int n_tensors,
                   const int32_t* __restrict__ chunks, const float* __restrict__ norms, double sched, float b1, float b2,
                   float eps, float max_grad_norm, int global_clip, float inv_bc1, float inv_sqrt_bc2,
                   const nbest_step_state* __restrict__ step_state) {
  pdl_grid_sync();
  if (step_state) {   // CUDA-graph replay: this step's schedule multiplier / bias corrections come from device memory
    sched = step_state->sched;
    inv_bc1 = step_state->inv_bc1;
    inv_sqrt_bc2 = step_state->inv_sqrt_bc2;
  }
  const int ti = chunks[3 * blockIdx.x], start = chunks[3 * blockIdx.x + 1], len = chunks[3 * blockIdx.x + 2];
  const nbest_adam_tensor t = tensors[ti];
  const int64_t base = t.offset + start;
  float coef = 1.0f;
  if (max_grad_norm > 0.f) {
    float sq = norms[ti];
    if (global_clip) {
      // torch.nn.utils.clip_grad_norm_ over ALL parameters (n_best_asr_bert.py:268-271): the 2-norm of the per-tensor
      // norms; every block folds the <= few hundred partial sums itself (inactive tensors hold 0)
      __shared__ float wsum[kThreads / 32];
      float a = 0.f;
      for (int i = threadIdx.x; i < n_tensors; i += kThreads) a += norms[i];
      a = warp_sum(a);
      if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = a;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) tot += wsum[w];     // same order in every block and on every rank
      sq = tot;
    }
    coef = fminf(max_grad_norm / (sqrtf(sq) + 1e-6f), 1.0f);   // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
  }
  const float lr_t = (float)(t.lr * sched);
  const float wd = t.weight_decay;
  const int n4 = (base & 3) == 0 ? (len >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    float4 pv = reinterpret_cast<float4*>(p + base)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + base) + i);
    float4 mv = reinterpret_cast<float4*>(m + base)[i];
    float4 vv = reinterpret_cast<float4*>(v + base)[i];
    adam_elem<MODE>(pv.x, gv.x, mv.x, vv.x, coef, b1, b2, eps, wd, lr_t, inv_bc1, inv_sqrt_bc2);
    adam_elem<MODE>(pv.y, gv.y, mv.y, vv.y, coef, b1, b2, eps, wd, lr_t, inv_bc1, inv_sqrt_bc2);
    adam_elem<MODE>(pv.z, gv.z, mv.z, vv.z, coef, b1, b2, eps, wd, lr_t, inv_bc1, inv_sqrt_bc2);
    adam_elem<MODE>(pv.w, gv.w, mv.w, vv.w, coef, b1, b2, eps, wd, lr_t, inv_bc1, inv_sqrt_bc2);
    reinterpret_cast<float4*>(p + base)[i] = pv;
    reinterpret_cast<float4*>(m + base)[i] = mv;
    reinterpret_cast<float4*>(v + base)[i] = vv;
    if (pb) *reinterpret_cast<uint2*>(pb + base + 4 * i) = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
  for (int i = (n4 << 2) + threadIdx.x; i < len; i += kThreads) {
    float pv = p[base + i], mv = m[base + i], vv = v[base + i];
    adam_elem<MODE>(pv, g[base + i], mv, vv, coef, b1, b2, eps, wd, lr_t, inv_bc1, inv_sqrt_bc2);
    p[base + i] = pv;
    m[base + i] = mv;
    v[base + i] = vv;
    if (pb) pb[base + i] = __float2bfloat16_rn(pv);
  }
}

}  // namespace

extern "C" int nbest_adam_step(nbest_ctx* ctx, int mode, float* p, const float* g, float* m, float* v, void* p_bf16,
                               const nbest_adam_tensor* tensors, int n_tensors, const int32_t* chunks, int n_chunks,
                               float* norms_ws, double sched, float b1, float b2, float eps, float max_grad_norm,
                               int global_clip, int step, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, p && g && m && v && tensors && chunks && norms_ws, "null pointer");
  NBEST_CHECK_ARG(ctx, n_tensors > 0 && n_chunks > 0, "empty tensor table");
  NBEST_CHECK_ARG(ctx, mode == NBEST_ADAM_BERT || mode == NBEST_ADAM_HF_ADAMW || mode == NBEST_ADAM_TORCH, "unknown mode");
  NBEST_CHECK_ARG(ctx, mode != NBEST_ADAM_TORCH || step >= 1, "torch Adam needs the 1-based step count (bias correction)");
  NBEST_CHECK_ARG(ctx, ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0, "flat buffers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (max_grad_norm > 0.f) {
    nbest_launch(adam_norm_kernel, dim3(n_chunks), dim3(kThreads), 0, s, g, tensors, chunks, norms_ws + n_tensors);
    NBEST_CHECK_LAUNCH(ctx);
    nbest_launch(adam_norm_finish_kernel, dim3(n_tensors), dim3(kThreads), 0, s, chunks, n_chunks, norms_ws + n_tensors, norms_ws);
    NBEST_CHECK_LAUNCH(ctx);
  }
  float inv_bc1 = 1.f, inv_sqrt_bc2 = 1.f;
  if (mode == NBEST_ADAM_TORCH) {
    inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, (double)step)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)step)));
  }
  auto* pb = reinterpret_cast<__nv_bfloat16*>(p_bf16);
  // nbest_ctx_set_step_indirect: the by-value scalars are placeholders, the kernel reads the context's step-state record
  const nbest_step_state* st = ctx->step_indirect ? ctx->step_state : nullptr;
  if (mode == NBEST_ADAM_BERT)
    nbest_launch(adam_update_kernel<NBEST_ADAM_BERT>, dim3(n_chunks), dim3(kThreads), 0, s, p, g, m, v, pb, tensors, n_tensors, chunks, norms_ws, sched,
                                                                       b1, b2, eps, max_grad_norm, global_clip, inv_bc1, inv_sqrt_bc2, st);
  else if (mode == NBEST_ADAM_HF_ADAMW)
    nbest_launch(adam_update_kernel<NBEST_ADAM_HF_ADAMW>, dim3(n_chunks), dim3(kThreads), 0, s, p, g, m, v, pb, tensors, n_tensors, chunks, norms_ws,
                                                                           sched, b1, b2, eps, max_grad_norm, global_clip, inv_bc1,
                                                                           inv_sqrt_bc2, st);
  else
    nbest_launch(adam_update_kernel<NBEST_ADAM_TORCH>, dim3(n_chunks), dim3(kThreads), 0, s, p, g, m, v, pb, tensors, n_tensors, chunks, norms_ws, sched,
                                                                        b1, b2, eps, max_grad_norm, global_clip, inv_bc1, inv_sqrt_bc2, st);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_bertadam_step(nbest_ctx* ctx, float* p, const float* g, float* m, float* v, void* p_bf16,
                                   const nbest_adam_tensor* tensors, int n_tensors, const int32_t* chunks, int n_chunks,
                                   float* norms_ws, double sched, float b1, float b2, float eps, float max_grad_norm,
                                   void* stream) {
  return nbest_adam_step(ctx, NBEST_ADAM_BERT, p, g, m, v, p_bf16, tensors, n_tensors, chunks, n_chunks, norms_ws, sched, b1, b2,
                         eps, max_grad_norm, 0, 0, stream);
}
