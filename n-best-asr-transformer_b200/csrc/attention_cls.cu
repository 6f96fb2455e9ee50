// Last-layer attention for the CLS query only.
//
// models/model.py:46-47,58 keep ONLY row 0 (the [CLS] position) of the encoder's last hidden state, so in the last
// encoder layer every query row other than [CLS] is dead work: its attention output, out-projection, FFN and both
// LayerNorms never reach the head. K and V of the last layer are still needed for all tokens (the CLS query attends to
// them), so the QKV projection stays full; everything after it runs on one row per sequence. Results are identical to
// the full computation (same arithmetic for the surviving row, same dropout mask indices).
//
// One warp per (sequence, head): lanes stride over the keys for the score / softmax pass (64-wide dot products against
// the register-resident query), then own two of the 64 output dims for the P.V pass. Backward mirrors it and writes the
// full dqkv rows of its sequence (dQ is zero except at the CLS row).
#include "common.h"
#include "ptx.cuh"

using namespace nbest;

namespace {

constexpr int D = 64;
constexpr int kWarps = 4;
constexpr int kMaxLen = 512;
constexpr float kLog2e = 1.4426950408889634f;


__device__ __forceinline__ void load_row64(const __nv_bfloat16* p, float (&v)[D]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
    v[8 * i + 0] = bf16lo(u.x);
    v[8 * i + 1] = bf16hi(u.x);
    v[8 * i + 2] = bf16lo(u.y);
    v[8 * i + 3] = bf16hi(u.y);
    v[8 * i + 4] = bf16lo(u.z);
    v[8 * i + 5] = bf16hi(u.z);
    v[8 * i + 6] = bf16lo(u.w);
    v[8 * i + 7] = bf16hi(u.w);
  }
}
__device__ __forceinline__ float dot_row64(const __nv_bfloat16* p, const float (&q)[D]) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + i);
    acc = fmaf(q[8 * i + 0], bf16lo(u.x), acc);
    acc = fmaf(q[8 * i + 1], bf16hi(u.x), acc);
    acc = fmaf(q[8 * i + 2], bf16lo(u.y), acc);
    acc = fmaf(q[8 * i + 3], bf16hi(u.y), acc);
    acc = fmaf(q[8 * i + 4], bf16lo(u.z), acc);
    acc = fmaf(q[8 * i + 5], bf16hi(u.z), acc);
    acc = fmaf(q[8 * i + 6], bf16lo(u.w), acc);
    acc = fmaf(q[8 * i + 7], bf16hi(u.w), acc);
  }
  return acc;
}

__global__ void __launch_bounds__(kWarps * 32)
attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu, const uint8_t* __restrict__ key_valid,
                    int B, int heads, int T, __nv_bfloat16* __restrict__ out_cls, float* __restrict__ lse_cls, float scale,
                    uint32_t thr, float rscale, uint32_t seed, const uint32_t* __restrict__ salt) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  __shared__ float sp[kWarps][kMaxLen];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * kWarps + warp;
  if (wid >= B * heads) return;
  const int b = wid / heads, h = wid - b * heads;
  const int s0 = cu[b], L = cu[b + 1] - s0;
  const int hd = heads * D;
  const int64_t ld = 3 * hd;
  __nv_bfloat16* orow = out_cls + (int64_t)b * hd + h * D + 2 * lane;
  if (L <= 0) {
    *reinterpret_cast<uint32_t*>(orow) = 0u;
    if (lane == 0) lse_cls[(int64_t)h * B + b] = 0.f;
    return;
  }
  float q[D];
  load_row64(qkv + (int64_t)s0 * ld + h * D, q);
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const bool ok = key_valid == nullptr || key_valid[s0 + j] != 0;
    const float s = ok ? dot_row64(qkv + (int64_t)(s0 + j) * ld + hd + h * D, q) : -INFINITY;
    sp[warp][j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  const float m_use = (mx == -INFINITY) ? 0.f : mx;
  const float sl2 = scale * kLog2e;
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float p = ex2_approx((sp[warp][j] - m_use) * sl2);
    sp[warp][j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
  for (int j = lane; j < L; j += 32) {
    float p = sp[warp][j] * inv;
    if (thr) p = attn_dropout_keep(seed, h, T, s0, j, thr) ? p * rscale : 0.f;
    sp[warp][j] = p;
  }
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  const __nv_bfloat16* vbase = qkv + (int64_t)s0 * ld + 2 * hd + h * D + 2 * lane;
  for (int j = 0; j < L; ++j) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(vbase + (int64_t)j * ld));
    const float p = sp[warp][j];
    a0 = fmaf(p, bf16lo(v), a0);
    a1 = fmaf(p, bf16hi(v), a1);
  }
  *reinterpret_cast<uint32_t*>(orow) = pack_bf16x2(a0, a1);
  if (lane == 0) lse_cls[(int64_t)h * B + b] = sum > 0.f ? mx * scale + logf(sum) : 0.f;
}

__global__ void __launch_bounds__(kWarps * 32)
attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu, const uint8_t* __restrict__ key_valid,
                    int B, int heads, int T, const __nv_bfloat16* __restrict__ out_cls, const __nv_bfloat16* __restrict__ dout_cls,
                    const float* __restrict__ lse_cls, int lse_stride, __nv_bfloat16* __restrict__ dqkv, float scale, uint32_t thr,
                    float rscale, uint32_t seed, const uint32_t* __restrict__ salt) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  __shared__ float sp[kWarps][kMaxLen];    // dropped probabilities  -> dV
  __shared__ float sds[kWarps][kMaxLen];   // dS                     -> dK, dQ
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * kWarps + warp;
  if (wid >= B * heads) return;
  const int b = wid / heads, h = wid - b * heads;
  const int s0 = cu[b], L = cu[b + 1] - s0;
  if (L <= 0) return;
  const int hd = heads * D;
  const int64_t ld = 3 * hd;
  float q[D], dO[D];
  load_row64(qkv + (int64_t)s0 * ld + h * D, q);
  load_row64(dout_cls + (int64_t)b * hd + h * D, dO);
  // delta = dO . O over the head's 64 dims
  const uint32_t o2 = __ldg(reinterpret_cast<const uint32_t*>(out_cls + (int64_t)b * hd + h * D + 2 * lane));
  const uint32_t g2 = __ldg(reinterpret_cast<const uint32_t*>(dout_cls + (int64_t)b * hd + h * D + 2 * lane));
  const uint32_t q2 = __ldg(reinterpret_cast<const uint32_t*>(qkv + (int64_t)s0 * ld + h * D + 2 * lane));
  const float delta = warp_sum(bf16lo(o2) * bf16lo(g2) + bf16hi(o2) * bf16hi(g2));
  const float lse2 = lse_cls[(int64_t)h * lse_stride + b] * kLog2e;
  const float sl2 = scale * kLog2e;
  for (int j = lane; j < L; j += 32) {
    const bool ok = key_valid == nullptr || key_valid[s0 + j] != 0;
    const float s = dot_row64(qkv + (int64_t)(s0 + j) * ld + hd + h * D, q);
    const float p = ok ? ex2_approx(s * sl2 - lse2) : 0.f;
    float dp = dot_row64(qkv + (int64_t)(s0 + j) * ld + 2 * hd + h * D, dO);
    float pd = p;
    if (thr) {
      const bool keep = attn_dropout_keep(seed, h, T, s0, j, thr);
      pd = keep ? p * rscale : 0.f;
      dp = keep ? dp * rscale : 0.f;
    }
    sp[warp][j] = pd;
    sds[warp][j] = p * (dp - delta);
  }
  __syncwarp();
  const float q0 = bf16lo(q2) * scale, q1 = bf16hi(q2) * scale, g0 = bf16lo(g2), g1 = bf16hi(g2);
  float dq0 = 0.f, dq1 = 0.f;
  const __nv_bfloat16* kbase = qkv + (int64_t)s0 * ld + hd + h * D + 2 * lane;
  __nv_bfloat16* drow = dqkv + (int64_t)s0 * ld + h * D + 2 * lane;
  for (int j = 0; j < L; ++j) {
    const float ds = sds[warp][j], pd = sp[warp][j];
    const uint32_t k2 = __ldg(reinterpret_cast<const uint32_t*>(kbase + (int64_t)j * ld));
    dq0 = fmaf(ds, bf16lo(k2), dq0);
    dq1 = fmaf(ds, bf16hi(k2), dq1);
    __nv_bfloat16* r = drow + (int64_t)j * ld;
    if (j > 0) *reinterpret_cast<uint32_t*>(r) = 0u;                                     // dQ of a non-CLS row
    *reinterpret_cast<uint32_t*>(r + hd) = pack_bf16x2(ds * q0, ds * q1);                // dK_j = scale * dS_j * q
    *reinterpret_cast<uint32_t*>(r + 2 * hd) = pack_bf16x2(pd * g0, pd * g1);            // dV_j = Pdrop_j * dO
  }
  *reinterpret_cast<uint32_t*>(drow) = pack_bf16x2(dq0 * scale, dq1 * scale);            // dQ_cls = scale * sum_j dS_j k_j
}

inline uint32_t drop_threshold(float p) {
  const double t = (double)p * 65536.0 + 0.5;
  return p <= 0.f ? 0u : (t >= 65535.0 ? 65535u : (uint32_t)t);
}

}  // namespace

extern "C" int nbest_attn_cls_fwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid,
                                  int B, int max_len, int heads, int T, void* out_cls_bf16, float* lse_cls, float p_drop,
                                  uint32_t seed, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, qkv_bf16 && cu_seqlens && out_cls_bf16 && lse_cls, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0 && heads > 0, "empty batch");
  NBEST_CHECK_ARG(ctx, max_len > 0 && max_len <= kMaxLen, "need 0 < max_len <= 512 (BERT position limit)");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  const int blocks = (B * heads + kWarps - 1) / kWarps;
  nbest_launch(attn_cls_fwd_kernel, dim3(blocks), dim3(kWarps * 32), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), cu_seqlens, key_valid, B, heads, T,
      reinterpret_cast<__nv_bfloat16*>(out_cls_bf16), lse_cls, 0.125f, drop_threshold(p_drop), 1.0f / (1.0f - p_drop), seed, nbest_salt(ctx));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_attn_cls_bwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid,
                                  int B, int max_len, int heads, int T, const void* out_cls_bf16, const void* dout_cls_bf16,
                                  const float* lse_cls, int lse_stride, void* dqkv_bf16, float p_drop, uint32_t seed,
                                  void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, qkv_bf16 && cu_seqlens && out_cls_bf16 && dout_cls_bf16 && lse_cls && dqkv_bf16, "null pointer");
  NBEST_CHECK_ARG(ctx, max_len > 0 && max_len <= kMaxLen, "need 0 < max_len <= 512 (BERT position limit)");
  NBEST_CHECK_ARG(ctx, B > 0 && heads > 0 && lse_stride >= B, "bad batch / lse stride");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  const int blocks = (B * heads + kWarps - 1) / kWarps;
  nbest_launch(attn_cls_bwd_kernel, dim3(blocks), dim3(kWarps * 32), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), cu_seqlens, key_valid, B, heads, T,
      reinterpret_cast<const __nv_bfloat16*>(out_cls_bf16), reinterpret_cast<const __nv_bfloat16*>(dout_cls_bf16), lse_cls,
      lse_stride, reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), 0.125f, drop_threshold(p_drop), 1.0f / (1.0f - p_drop), seed, nbest_salt(ctx));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}
