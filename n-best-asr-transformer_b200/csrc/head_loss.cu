// CLS gather + hierarchical STC head + the four-term loss, forward and backward, as a handful of fused kernels.
//
// Reference: models/model.py:46-47 (CLS row), models/modules/hierarchical_classifier.py:35-60 (11 Linear calls, sigmoid,
// 10 softmaxes, 30 scatters into final_scores), n_best_asr_bert.py:145-195 + utils/STC_util.py:4-51 (BCE-sum on final and
// top scores, CE-sum per value group averaged over groups, optional MSE-mean between the two CLS vectors) and
// n_best_asr_bert.py:198-215 (decode). The reference spends ~60 micro-kernels and >= 14 host syncs here; this file is
// 2 launches forward (+1 for the loss) and 2-3 backward, with no host sync.
#include "common.h"
#include "ptx.cuh"

using namespace nbest;

namespace {

constexpr int H = 768;
constexpr int kMaxCols = 256;   // n_cols (171 for DSTC2) must fit
constexpr int kMaxBottom = 256; // n_bottom (161)
constexpr int kMaxGroups = 32;

struct Hier {
  int n_top, n_bottom, n_groups, n_cols;
  const int32_t *col_group, *col_bottom, *grp_off, *grp_top;
};

__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

// BCELoss(sum) element: torch clamps each log at -100 in the forward; its backward is the closed form
// (p - t) / max(p * (1 - p), 1e-12) (ATen binary_cross_entropy_backward), which is what autograd gives the reference.
__device__ __forceinline__ float bce_elem(float p, float t, float& dp) {
  const float a = fmaxf(logf(p), -100.f), b = fmaxf(logf(1.0f - p), -100.f);
  dp = (p - t) / fmaxf((1.0f - p) * p, 1e-12f);
  return -(t * a + (1.0f - t) * b);
}

// ------------------------------------------------------------------------------------------------ head forward
__global__ void __launch_bounds__(256)
stc_head_fwd_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ cu, int B, const float* __restrict__ W,
                    const float* __restrict__ bias, Hier h, const uint8_t* __restrict__ none_col, uint32_t thr, float rscale,
                    uint32_t seed, const uint32_t* __restrict__ salt, float* __restrict__ cls, float* __restrict__ logits, float* __restrict__ top_scores,
                    float* __restrict__ bottom_scores, float* __restrict__ final_scores, uint8_t* __restrict__ decode) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  __shared__ float f[H];
  __shared__ float z[kMaxCols];
  __shared__ float sc[kMaxCols];   // sigmoid / softmax of z
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const __nv_bfloat16* row = x + (int64_t)cu[b] * H;
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    const float v = __bfloat162float(row[d]);
    f[d] = v;
    cls[(int64_t)b * H + d] = v;
  }
  __syncthreads();
  for (int c = warp; c < h.n_cols; c += 8) {
    const int g = h.col_group[c];
    const float* w = W + (int64_t)c * H;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < H / 128; ++i) {
      const int d = 4 * (lane + 32 * i);
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
      float4 fv = *reinterpret_cast<const float4*>(f + d);
      if (thr) {
        const uint32_t base = ((uint32_t)g * (uint32_t)B + (uint32_t)b) * H + d;
        bool k0_, k1_, k2_, k3_;
        dropout_keep4(seed, base, thr, k0_, k1_, k2_, k3_);
        fv.x = k0_ ? fv.x * rscale : 0.f;
        fv.y = k1_ ? fv.y * rscale : 0.f;
        fv.z = k2_ ? fv.z * rscale : 0.f;
        fv.w = k3_ ? fv.w * rscale : 0.f;
      }
      acc += wv.x * fv.x + wv.y * fv.y + wv.z * fv.z + wv.w * fv.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) z[c] = acc + bias[c];
  }
  __syncthreads();
  // sigmoid on act-slot columns, softmax per value group (one warp per group)
  for (int c = threadIdx.x; c < h.n_top; c += blockDim.x) sc[c] = sigmoidf_(z[c]);
  for (int g = warp; g < h.n_groups; g += 8) {
    const int c0 = h.grp_off[g], c1 = h.grp_off[g + 1];
    float mx = -INFINITY;
    for (int c = c0 + lane; c < c1; c += 32) mx = fmaxf(mx, z[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = c0 + lane; c < c1; c += 32) {
      const float e = expf(z[c] - mx);
      sc[c] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int c = c0 + lane; c < c1; c += 32) sc[c] = sc[c] / s;
  }
  __syncthreads();
  const int nb = h.n_cols - h.n_top;
  for (int c = threadIdx.x; c < h.n_cols; c += blockDim.x) {
    logits[(int64_t)b * h.n_cols + c] = z[c];
    if (c < h.n_top) {
      top_scores[(int64_t)b * h.n_top + c] = sc[c];
      const int bot = h.col_bottom[c];
      if (bot >= 0) {
        final_scores[(int64_t)b * h.n_bottom + bot] = sc[c];
        if (decode) decode[(int64_t)b * h.n_bottom + bot] = sc[c] > 0.5f ? 1 : 0;
      }
    } else {
      bottom_scores[(int64_t)b * nb + (c - h.n_top)] = sc[c];
      const int g = h.col_group[c] - 1;
      final_scores[(int64_t)b * h.n_bottom + h.col_bottom[c]] = sc[h.grp_top[g]] * sc[c];
    }
  }
  // decode (pred_one_sample): act-slot > 0.5 -> argmax value unless it is the NONE value
  if (decode) {
    for (int c = h.n_top + threadIdx.x; c < h.n_cols; c += blockDim.x) decode[(int64_t)b * h.n_bottom + h.col_bottom[c]] = 0;
    __syncthreads();
    for (int g = warp; g < h.n_groups; g += 8) {
      if (lane == 0 && sc[h.grp_top[g]] > 0.5f) {
        const int c0 = h.grp_off[g], c1 = h.grp_off[g + 1];
        int best = c0;
        for (int c = c0 + 1; c < c1; ++c)
          if (sc[c] > sc[best]) best = c;   // first maximum, like numpy argmax
        if (!(none_col && none_col[best])) decode[(int64_t)b * h.n_bottom + h.col_bottom[best]] = 1;
      }
    }
  }
}

// Shared tail: upstream gradients w.r.t. (top s, bottom q, final) -> dlogits, one block per batch row.
// ds[c] for c < n_top holds dL/ds_i (direct part), dq[c] for c >= n_top holds dL/dq_j (direct part),
// dfin[c] holds dL/dfinal of the bottom label scored by column c (0 for act-slots that own a group).
__device__ __forceinline__ void scores_jacobian(const Hier& h, const float* sc, float* ds, float* dq, const float* dfin,
                                                float* __restrict__ dlogits_row) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // fold final = s_i * q_j (groups) and final = s_i (singletons) into ds / dq
  for (int c = threadIdx.x; c < h.n_top; c += blockDim.x)
    if (h.col_bottom[c] >= 0) ds[c] += dfin[c];
  __syncthreads();
  for (int g = warp; g < h.n_groups; g += 8) {
    const int c0 = h.grp_off[g], c1 = h.grp_off[g + 1], ti = h.grp_top[g];
    const float s = sc[ti];
    float acc_s = 0.f, dot = 0.f;
    for (int c = c0 + lane; c < c1; c += 32) {
      acc_s += dfin[c] * sc[c];
      const float d = dq[c] + dfin[c] * s;
      dq[c] = d;
      dot += d * sc[c];
    }
    acc_s = warp_sum(acc_s);
    dot = warp_sum(dot);
    if (lane == 0) ds[ti] += acc_s;
    for (int c = c0 + lane; c < c1; c += 32) dlogits_row[c] = sc[c] * (dq[c] - dot);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h.n_top; c += blockDim.x) dlogits_row[c] = ds[c] * sc[c] * (1.0f - sc[c]);
}

__device__ __forceinline__ void recompute_scores(const Hier& h, const float* z, float* sc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < h.n_top; c += blockDim.x) sc[c] = sigmoidf_(z[c]);
  for (int g = warp; g < h.n_groups; g += 8) {
    const int c0 = h.grp_off[g], c1 = h.grp_off[g + 1];
    float mx = -INFINITY;
    for (int c = c0 + lane; c < c1; c += 32) mx = fmaxf(mx, z[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = c0 + lane; c < c1; c += 32) {
      const float e = expf(z[c] - mx);
      sc[c] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int c = c0 + lane; c < c1; c += 32) sc[c] = sc[c] / s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ fused loss fwd+bwd
__global__ void __launch_bounds__(256)
stc_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int B, Hier h,
                float* __restrict__ losses, float* __restrict__ dlogits) {
  pdl_grid_sync();
  __shared__ float z[kMaxCols], sc[kMaxCols], ds[kMaxCols], dq[kMaxCols], dfin[kMaxCols];
  __shared__ float y[kMaxBottom];
  __shared__ float red[3];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < h.n_cols; c += blockDim.x) {
    z[c] = logits[(int64_t)b * h.n_cols + c];
    ds[c] = dq[c] = dfin[c] = 0.f;
  }
  for (int j = threadIdx.x; j < h.n_bottom; j += blockDim.x) y[j] = labels[(int64_t)b * h.n_bottom + j];
  if (threadIdx.x < 3) red[threadIdx.x] = 0.f;
  __syncthreads();
  recompute_scores(h, z, sc);
  float l_final = 0.f, l_top = 0.f, l_ce = 0.f;
  // BCE(final, y): one element per bottom label = per column that scores a bottom label
  for (int c = threadIdx.x; c < h.n_cols; c += blockDim.x) {
    const int bot = h.col_bottom[c];
    if (bot < 0) continue;
    const float p = c < h.n_top ? sc[c] : sc[h.grp_top[h.col_group[c] - 1]] * sc[c];
    float dp;
    l_final += bce_elem(p, y[bot], dp);
    dfin[c] = dp;
  }
  // BCE(top, Y) with Y_i = sum of the labels below act-slot i  (convert_labels = labels @ b2t)
  for (int c = threadIdx.x; c < h.n_top; c += blockDim.x) {
    float Y;
    if (h.col_bottom[c] >= 0) {
      Y = y[h.col_bottom[c]];
    } else {
      Y = 0.f;
      for (int g = 0; g < h.n_groups; ++g)
        if (h.grp_top[g] == c)
          for (int cc = h.grp_off[g]; cc < h.grp_off[g + 1]; ++cc) Y += y[h.col_bottom[cc]];
    }
    float dp;
    l_top += bce_elem(sc[c], Y, dp);
    ds[c] = dp;
  }
  // CE per group: target = active label or the last (NONE) one; NLL(sum) of log(q + 1e-12), averaged over groups
  for (int g = warp; g < h.n_groups; g += 8) {
    if (lane == 0) {
      const int c0 = h.grp_off[g], c1 = h.grp_off[g + 1];
      int tgt = c0;
      float best = y[h.col_bottom[c0]], sum = 0.f;
      for (int c = c0; c < c1; ++c) {
        const float v = y[h.col_bottom[c]];
        sum += v;
        if (v > best) {
          best = v;
          tgt = c;
        }
      }
      if (sum == 0.f) tgt = c1 - 1;
      const float inv_g = 1.0f / (float)h.n_groups;
      l_ce += -logf(sc[tgt] + 1e-12f) * inv_g;
      dq[tgt] = -inv_g / (sc[tgt] + 1e-12f);
    }
  }
  l_final = warp_sum(l_final);
  l_top = warp_sum(l_top);
  l_ce = warp_sum(l_ce);
  if (lane == 0) {
    atomicAdd(&red[0], l_final);
    atomicAdd(&red[1], l_top);
    atomicAdd(&red[2], l_ce);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&losses[1], red[0]);
    atomicAdd(&losses[2], red[1]);
    atomicAdd(&losses[3], red[2]);
  }
  scores_jacobian(h, sc, ds, dq, dfin, dlogits + (int64_t)b * h.n_cols);
}

__global__ void mse_kernel(const float* __restrict__ a, const float* __restrict__ t, int64_t n, float coef /*scale/n*/,
                           float* __restrict__ losses, float* __restrict__ da, float* __restrict__ dt) {
  pdl_grid_sync();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - t[i];
    acc += d * d;
    const float g = 2.0f * d * coef;
    if (da) da[i] = g;
    if (dt) dt[i] = -g;
  }
  acc = warp_sum(acc);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(&losses[0], v * coef);
  }
}

__global__ void __launch_bounds__(256)
stc_scores_bwd_kernel(const float* __restrict__ top_scores, const float* __restrict__ bottom_scores,
                      const float* __restrict__ d_top, const float* __restrict__ d_bottom, const float* __restrict__ d_final,
                      int B, Hier h, float* __restrict__ dlogits) {
  pdl_grid_sync();
  __shared__ float sc[kMaxCols], ds[kMaxCols], dq[kMaxCols], dfin[kMaxCols];
  const int b = blockIdx.x;
  const int nb = h.n_cols - h.n_top;
  for (int c = threadIdx.x; c < h.n_cols; c += blockDim.x) {
    const int bot = h.col_bottom[c];
    ds[c] = dq[c] = 0.f;
    dfin[c] = (d_final && bot >= 0) ? d_final[(int64_t)b * h.n_bottom + bot] : 0.f;
    if (c < h.n_top) {
      sc[c] = top_scores[(int64_t)b * h.n_top + c];
      if (d_top) ds[c] = d_top[(int64_t)b * h.n_top + c];
    } else {
      sc[c] = bottom_scores[(int64_t)b * nb + (c - h.n_top)];
      if (d_bottom) dq[c] = d_bottom[(int64_t)b * nb + (c - h.n_top)];
    }
  }
  __syncthreads();
  scores_jacobian(h, sc, ds, dq, dfin, dlogits + (int64_t)b * h.n_cols);
}

// ------------------------------------------------------------------------------------------------ head backward
// dW[c,:] += sum_b dl[b,c] * drop_g(c)(f[b,:]),  dbias[c] += sum_b dl[b,c]
// grid (column, batch split): 192 threads own four consecutive features each (one dropout quad per utterance), the
// batch is cut into kWgradSplits slices that meet in fp32 red.add on the flat gradient (one column x 256 utterances in
// a single block was a 100 us serial chain of dependent loads).
constexpr int kWgradSplits = 8;
__global__ void __launch_bounds__(192)
stc_head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ cls, int B, Hier h, uint32_t thr,
                      float rscale, uint32_t seed, const uint32_t* __restrict__ salt, float* __restrict__ dW, float* __restrict__ dbias) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  const int c = blockIdx.x;
  const int g = h.col_group[c];
  const int per = (B + kWgradSplits - 1) / kWgradSplits;
  const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  const int d = 4 * threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float dl = __ldg(dlogits + (int64_t)b * h.n_cols + c);
    float4 fv = __ldg(reinterpret_cast<const float4*>(cls + (int64_t)b * H + d));
    sb += dl;
    if (thr) {
      bool k0, k1, k2, k3;
      dropout_keep4(seed, ((uint32_t)g * (uint32_t)B + (uint32_t)b) * H + d, thr, k0, k1, k2, k3);
      fv.x = k0 ? fv.x * rscale : 0.f;
      fv.y = k1 ? fv.y * rscale : 0.f;
      fv.z = k2 ? fv.z * rscale : 0.f;
      fv.w = k3 ? fv.w * rscale : 0.f;
    }
    acc.x = fmaf(dl, fv.x, acc.x);
    acc.y = fmaf(dl, fv.y, acc.y);
    acc.z = fmaf(dl, fv.z, acc.z);
    acc.w = fmaf(dl, fv.w, acc.w);
  }
  if (b1 > b0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dW + (int64_t)c * H + d), "f"(acc.x), "f"(acc.y), "f"(acc.z),
                 "f"(acc.w)
                 : "memory");
    if (threadIdx.x == 0) atomicAdd(dbias + c, sb);
  }
}

// dcls[b,:] (=|+=) sum_c dl[b,c] * W[c,:] * mask_g(c)[b,:]                          (one block per batch row)
// 192 threads own four consecutive features each. A column's dropout mask only depends on its GROUP (the 11 nn.Linear
// calls of hierarchical_classifier.py:41,46 each drew one mask of the feature), so the <= 16 group masks are hashed once
// per thread (one quad each) instead of once per column and feature (171 x 3 hashes before).
__global__ void __launch_bounds__(192)
stc_head_dgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ W, int B, Hier h, uint32_t thr,
                      float rscale, uint32_t seed, const uint32_t* __restrict__ salt, float* __restrict__ dcls, int accumulate) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  __shared__ float dl[kMaxCols];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < h.n_cols; c += blockDim.x) dl[c] = dlogits[(int64_t)b * h.n_cols + c];
  const int d = 4 * threadIdx.x;
  uint64_t keep_bits = ~0ull;          // bit 4 g + i: feature d + i survives group g's mask
  if (thr) {
    keep_bits = 0;
    const int ng = h.n_groups + 1 < 16 ? h.n_groups + 1 : 16;   // host checks n_groups + 1 <= 16 (4 keep bits per group)
    for (int g = 0; g < ng; ++g) {
      bool k0, k1, k2, k3;
      dropout_keep4(seed, ((uint32_t)g * (uint32_t)B + (uint32_t)b) * H + d, thr, k0, k1, k2, k3);
      keep_bits |= (uint64_t)((k0 ? 1u : 0u) | (k1 ? 2u : 0u) | (k2 ? 4u : 0u) | (k3 ? 8u : 0u)) << (4 * g);
    }
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int c = 0; c < h.n_cols; ++c) {
    const uint32_t m = (uint32_t)(keep_bits >> (4 * h.col_group[c])) & 15u;
    const float v = dl[c];
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + (int64_t)c * H + d));
    acc.x = fmaf(v, (m & 1u) ? w.x : 0.f, acc.x);
    acc.y = fmaf(v, (m & 2u) ? w.y : 0.f, acc.y);
    acc.z = fmaf(v, (m & 4u) ? w.z : 0.f, acc.z);
    acc.w = fmaf(v, (m & 8u) ? w.w : 0.f, acc.w);
  }
  const float sc = thr ? rscale : 1.0f;
  float4* o = reinterpret_cast<float4*>(dcls + (int64_t)b * H + d);
  float4 r = make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc);
  if (accumulate) {
    const float4 prev = *o;
    r.x += prev.x;
    r.y += prev.y;
    r.z += prev.z;
    r.w += prev.w;
  }
  *o = r;
}

__global__ void cls_scatter_kernel(const float* __restrict__ dcls, const int32_t* __restrict__ cu,
                                   __nv_bfloat16* __restrict__ dx) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  __nv_bfloat16* row = dx + (int64_t)cu[b] * H;
  for (int d = threadIdx.x; d < H; d += blockDim.x) row[d] = __float2bfloat16_rn(dcls[(int64_t)b * H + d]);
}

// One warp per utterance: TP / FP / FN over the (optionally ontology-filtered) label columns and the exact-match flag.
__global__ void __launch_bounds__(256)
stc_metrics_kernel(const uint8_t* __restrict__ decode, const float* __restrict__ labels, const uint8_t* __restrict__ col_mask,
                   int B, int nb, unsigned long long* __restrict__ counters) {
  pdl_grid_sync();
  __shared__ unsigned int part[4];
  if (threadIdx.x < 4) part[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b < B) {
    unsigned int tp = 0, fp = 0, fn = 0;
    for (int c = lane; c < nb; c += 32) {
      if (col_mask != nullptr && col_mask[c] == 0) continue;
      const bool p = decode[(int64_t)b * nb + c] != 0, g = labels[(int64_t)b * nb + c] > 0.5f;
      tp += (p && g) ? 1u : 0u;
      fp += (p && !g) ? 1u : 0u;
      fn += (!p && g) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tp += __shfl_xor_sync(0xffffffffu, tp, o);
      fp += __shfl_xor_sync(0xffffffffu, fp, o);
      fn += __shfl_xor_sync(0xffffffffu, fn, o);
    }
    if (lane == 0) {
      atomicAdd(&part[0], tp);
      atomicAdd(&part[1], fp);
      atomicAdd(&part[2], fn);
      atomicAdd(&part[3], (fp == 0 && fn == 0) ? 1u : 0u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 4 && part[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)part[threadIdx.x]);
}

inline uint32_t drop_threshold(float p) {
  const double t = (double)p * 65536.0 + 0.5;   // 16-bit threshold (ptx.cuh dropout_keep)
  return p <= 0.f ? 0u : (t >= 65535.0 ? 65535u : (uint32_t)t);
}

int check_hier(nbest_ctx* ctx, const nbest_hierarchy* h, Hier* out) {
  NBEST_CHECK_ARG(ctx, h && h->col_group && h->col_bottom && h->grp_off && h->grp_top, "null hierarchy");
  NBEST_CHECK_ARG(ctx, h->n_cols > 0 && h->n_cols <= kMaxCols && h->n_bottom > 0 && h->n_bottom <= kMaxBottom &&
                           h->n_groups >= 0 && h->n_groups <= kMaxGroups && h->n_top > 0 && h->n_top <= h->n_cols,
                  "hierarchy sizes out of range");
  out->n_top = h->n_top;
  out->n_bottom = h->n_bottom;
  out->n_groups = h->n_groups;
  out->n_cols = h->n_cols;
  out->col_group = h->col_group;
  out->col_bottom = h->col_bottom;
  out->grp_off = h->grp_off;
  out->grp_top = h->grp_top;
  return NBEST_OK;
}

}  // namespace

extern "C" int nbest_stc_head_fwd(nbest_ctx* ctx, const void* x_bf16, const int32_t* cu_seqlens, int B, int hidden,
                                  const float* W, const float* bias, const nbest_hierarchy* hh, const uint8_t* none_col_mask,
                                  float p_drop, uint32_t seed, float* cls, float* logits, float* top_scores,
                                  float* bottom_scores, float* final_scores, uint8_t* decode, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  Hier h;
  int rc = check_hier(ctx, hh, &h);
  if (rc) return rc;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, x_bf16 && cu_seqlens && W && bias && cls && logits && top_scores && bottom_scores && final_scores,
                  "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0, "empty batch");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  nbest_launch(stc_head_fwd_kernel, dim3(B), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), cu_seqlens, B, W, bias, h, none_col_mask, drop_threshold(p_drop),
      1.0f / (1.0f - p_drop), seed, nbest_salt(ctx), cls, logits, top_scores, bottom_scores, final_scores, decode);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_stc_loss_fwd_bwd(nbest_ctx* ctx, const float* logits, const float* labels, int B,
                                      const nbest_hierarchy* hh, const float* asr_cls, const float* trans_cls, int hidden,
                                      float mse_scale, float* losses, float* dlogits, float* d_asr_cls, float* d_trans_cls,
                                      void* stream) {
  if (!ctx) return NBEST_EINVAL;
  Hier h;
  int rc = check_hier(ctx, hh, &h);
  if (rc) return rc;
  NBEST_CHECK_ARG(ctx, logits && labels && losses && dlogits, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0, "empty batch");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  nbest_launch(stc_loss_kernel, dim3(B), dim3(256), 0, s, logits, labels, B, h, losses, dlogits);
  NBEST_CHECK_LAUNCH(ctx);
  if (asr_cls && trans_cls) {
    const int64_t n = (int64_t)B * hidden;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 256) blocks = 256;
    nbest_launch(mse_kernel, dim3(blocks), dim3(256), 0, s, asr_cls, trans_cls, n, mse_scale / (float)n, losses, d_asr_cls, d_trans_cls);
    NBEST_CHECK_LAUNCH(ctx);
  }
  return NBEST_OK;
}

extern "C" int nbest_stc_scores_bwd(nbest_ctx* ctx, const float* top_scores, const float* bottom_scores, const float* d_top,
                                    const float* d_bottom, const float* d_final, int B, const nbest_hierarchy* hh,
                                    float* dlogits, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  Hier h;
  int rc = check_hier(ctx, hh, &h);
  if (rc) return rc;
  NBEST_CHECK_ARG(ctx, top_scores && bottom_scores && dlogits, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0, "empty batch");
  nbest_launch(stc_scores_bwd_kernel, dim3(B), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), top_scores, bottom_scores, d_top, d_bottom,
                                                                                d_final, B, h, dlogits);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_stc_head_bwd(nbest_ctx* ctx, const float* dlogits, const float* cls, const float* W, int B, int hidden,
                                  const nbest_hierarchy* hh, float p_drop, uint32_t seed, float* dW, float* dbias,
                                  float* dcls, int accumulate_dcls, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  Hier h;
  int rc = check_hier(ctx, hh, &h);
  if (rc) return rc;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, dlogits && cls && W && dW && dbias && dcls, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0, "empty batch");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thr = drop_threshold(p_drop);
  const float rscale = 1.0f / (1.0f - p_drop);
  nbest_launch(stc_head_wgrad_kernel, dim3(dim3(h.n_cols, kWgradSplits)), dim3(192), 0, s, dlogits, cls, B, h, thr, rscale, seed, nbest_salt(ctx), dW, dbias);
  NBEST_CHECK_LAUNCH(ctx);
  NBEST_CHECK_ARG(ctx, h.n_groups + 1 <= 16, "at most 15 multi-way value groups (4 keep bits per group in one word)");
  nbest_launch(stc_head_dgrad_kernel, dim3(B), dim3(192), 0, s, dlogits, W, B, h, thr, rscale, seed, nbest_salt(ctx), dcls, accumulate_dcls);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_stc_metrics(nbest_ctx* ctx, const uint8_t* decode, const float* labels, const uint8_t* col_mask, int B,
                                 int n_bottom, long long* counters, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, decode && labels && counters, "null pointer");
  NBEST_CHECK_ARG(ctx, B >= 0 && n_bottom > 0, "bad shape");
  if (B == 0) return NBEST_OK;
  nbest_launch(stc_metrics_kernel, dim3((B + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      decode, labels, col_mask, B, n_bottom, reinterpret_cast<unsigned long long*>(counters));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_cls_scatter(nbest_ctx* ctx, const float* dcls, const int32_t* cu_seqlens, int B, int T, int hidden,
                                 void* dx_bf16, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, dcls && cu_seqlens && dx_bf16, "null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  NBEST_CHECK_CUDA(ctx, cudaMemsetAsync(dx_bf16, 0, (size_t)T * H * 2, s));
  nbest_launch(cls_scatter_kernel, dim3(B), dim3(256), 0, s, dcls, cu_seqlens, reinterpret_cast<__nv_bfloat16*>(dx_bf16));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}
