// Memory-bound kernels of the packed encoder: batch packing, embedding gather + LayerNorm, LayerNorm fwd/bwd,
// column sums and the fp32 -> bf16 weight refresh. All are one-warp-per-token (hidden = 768 = 32 lanes x 6 x float4),
// 8/16-byte vector accesses, warp-shuffle reductions, fp32 statistics.
//
// Reference counterparts: utils/bert_xlnet_inputs.py:91-102 + models/model.py:43 (packing / key mask),
// transformers modeling_bert.py:102-112 (BertEmbeddings), :294-298 and :352-356 (residual LayerNorms).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

using namespace nbest;

namespace {

constexpr int H = 768;          // models/model.py:30 hard-codes fea_dim = 768
constexpr int VPL = H / 128;    // float4 vectors per lane (6)
constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 bf16x4_to_f4(uint2 u) { return make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y)); }
__device__ __forceinline__ uint2 f4_to_bf16x4(float4 v) { return make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)); }
static inline uint32_t drop_threshold(float p) {
  const double t = (double)p * 65536.0 + 0.5;   // 16-bit threshold (ptx.cuh dropout_keep)
  return p <= 0.f ? 0u : (t >= 65535.0 ? 65535u : (uint32_t)t);
}

// ------------------------------------------------------------------------------------------------ packing
__global__ void pack_lens_kernel(const int64_t* __restrict__ ids, int B, int S, int32_t* __restrict__ lens) {
  pdl_grid_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  int last = 0;
  for (int j = lane; j < S; j += 32)
    if (ids[(int64_t)row * S + j] > 0) last = j + 1;
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) lens[row] = last;
}

__global__ void pack_scan_kernel(const int32_t* __restrict__ lens, int B, int32_t* __restrict__ cu) {
  pdl_grid_sync();
  __shared__ int32_t sh[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) {
    carry = 0;
    cu[0] = 0;
  }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    int v = i < B ? lens[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < B) cu[i + 1] = carry + sh[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
}

__global__ void pack_scatter_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ seg_ids, int B, int S,
                                    int pos_mode, const int32_t* __restrict__ cu, int32_t* __restrict__ tokens,
                                    uint8_t* __restrict__ seg, int32_t* __restrict__ pos, int32_t* __restrict__ seq_of,
                                    uint8_t* __restrict__ key_valid, int row0) {
  pdl_grid_sync();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  cu += row0;                               // this stream's sequences sit at [row0, row0 + B) of the merged batch
  const int t0 = cu[row], L = cu[row + 1] - cu[row];
  int carry = 0;
  for (int c = 0; c < L; c += 32) {
    const int j = c + lane;
    const bool in = j < L;
    const int64_t id = in ? ids[(int64_t)row * S + j] : 1;
    const bool nonpad = in && id != 1;
    const uint32_t m = __ballot_sync(0xffffffffu, nonpad);
    if (in) {
      const int t = t0 + j;
      tokens[t] = (int32_t)id;
      seg[t] = seg_ids ? (uint8_t)seg_ids[(int64_t)row * S + j] : (uint8_t)0;
      seq_of[t] = row0 + row;
      key_valid[t] = id > 0 ? 1 : 0;
      if (pos_mode == 1) {
        const int incl = carry + __popc(m & (0xffffffffu >> (31 - lane)));
        pos[t] = nonpad ? incl + 1 : 1;
      } else {
        pos[t] = j;
      }
    }
    carry += __popc(m);
  }
}

// hyp_id[t] = number of separator tokens before position t of its sequence (one warp per sequence, ballot scan)
__global__ void pack_hyp_ids_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ cu, int B, int sep_id,
                                    uint8_t* __restrict__ hyp_id) {
  pdl_grid_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const int t0 = cu[row], L = cu[row + 1] - cu[row];
  int carry = 0;
  for (int c = 0; c < L; c += 32) {
    const int j = c + lane;
    const bool is_sep = j < L && tokens[t0 + j] == sep_id;
    const uint32_t m = __ballot_sync(0xffffffffu, is_sep);
    if (j < L) hyp_id[t0 + j] = (uint8_t)min(255, carry + __popc(m & ((1u << lane) - 1u)));
    carry += __popc(m);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm core
struct RowStats {
  float mean, rstd;
};

__device__ __forceinline__ RowStats row_stats(const float4 (&x)[VPL], float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) * (1.0f / H);
  return RowStats{mean, rsqrtf(var + eps)};
}

// y = dropout(LN(x)) for one row held in registers; writes bf16.
__device__ __forceinline__ void ln_apply_store(const float4 (&x)[VPL], RowStats st, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, __nv_bfloat16* __restrict__ yrow, int lane,
                                               uint32_t thr, float scale, uint32_t seed, uint32_t row_base) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (x[i].x - st.mean) * st.rstd * g.x + b.x;
    y.y = (x[i].y - st.mean) * st.rstd * g.y + b.y;
    y.z = (x[i].z - st.mean) * st.rstd * g.z + b.z;
    y.w = (x[i].w - st.mean) * st.rstd * g.w + b.w;
    if (thr) {
      bool k0_, k1_, k2_, k3_;
      dropout_keep4(seed, row_base + c, thr, k0_, k1_, k2_, k3_);
      y.x = k0_ ? y.x * scale : 0.f;
      y.y = k1_ ? y.y * scale : 0.f;
      y.z = k2_ ? y.z * scale : 0.f;
      y.w = k3_ ? y.w * scale : 0.f;
    }
    *reinterpret_cast<uint2*>(yrow + c) = f4_to_bf16x4(y);
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_ln_fwd_kernel(const int32_t* __restrict__ tokens, const uint8_t* __restrict__ seg, const int32_t* __restrict__ pos,
                    int T, const float* __restrict__ word, const float* __restrict__ posemb, const float* __restrict__ type,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                    __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, uint32_t thr,
                    float scale, uint32_t seed, const uint32_t* __restrict__ salt) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  const int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= T) return;
  const int lane = threadIdx.x & 31;
  const float* w = word + (int64_t)tokens[t] * H;
  const float* p = posemb + (int64_t)pos[t] * H;
  const float* ty = type + (int64_t)seg[t] * H;
  float4 x[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + c));
    const float4 d = __ldg(reinterpret_cast<const float4*>(ty + c));
    x[i] = make_float4(a.x + d.x + b.x, a.y + d.y + b.y, a.z + d.z + b.z, a.w + d.w + b.w);  // word + type + pos (HF order)
  }
  const RowStats st = row_stats(x, eps);
  if (lane == 0) {
    mean[t] = st.mean;
    rstd[t] = st.rstd;
  }
  ln_apply_store(x, st, gamma, beta, y + (int64_t)t * H, lane, thr, scale, seed, (uint32_t)t * H);
}

// ---- 16-byte row chunks: lane owns columns 8 (lane + 32 i) + k, i < 3, k < 8 (a 768-wide bf16 row = 96 chunks) ----
constexpr int CPL = H / 256;   // 16-byte chunks per lane (3)

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  v[0] = bf16lo(u.x);
  v[1] = bf16hi(u.x);
  v[2] = bf16lo(u.y);
  v[3] = bf16hi(u.y);
  v[4] = bf16lo(u.z);
  v[5] = bf16hi(u.z);
  v[6] = bf16lo(u.w);
  v[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void load8f(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Residual LayerNorm forward. Each warp normalises kRowsPerWarp rows whose 16-byte loads are all issued up front (the
// kernel is pure streaming: memory-level parallelism per warp is what sets its bandwidth).
template <int kLnFwdRows>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kLnFwdRows <= 2 ? 4 : 2)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ xin, const float* __restrict__ gamma, const float* __restrict__ beta,
              float eps, int T, __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * kLnFwdRows;
  if (t0 >= T) return;
  uint4 raw[kLnFwdRows][CPL];
#pragma unroll
  for (int r = 0; r < kLnFwdRows; ++r) {
    const int t = min(t0 + r, T - 1);
#pragma unroll
    for (int i = 0; i < CPL; ++i) raw[r][i] = __ldg(reinterpret_cast<const uint4*>(xin + (int64_t)t * H) + lane + 32 * i);
  }
#pragma unroll
  for (int r = 0; r < kLnFwdRows; ++r) {
    const int t = t0 + r;
    if (t >= T) break;
    float x[CPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      unpack8(raw[r][i], x[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += x[i][k];
    }
    const float mu = warp_sum(s) * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = x[i][k] - mu;
        q = fmaf(d, d, q);
      }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
    if (lane == 0) {
      if (mean) mean[t] = mu;
      if (rstd) rstd[t] = rs;
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = 8 * (lane + 32 * i);
      float g[8], b[8], o[8];
      load8f(gamma + c, g);
      load8f(beta + c, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf((x[i][k] - mu) * rs, g[k], b[k]);
      *(reinterpret_cast<uint4*>(y + (int64_t)t * H) + lane + 32 * i) = pack8(o);
    }
  }
}

// Residual LayerNorm forward with the row statistics supplied by the producing GEMM (NBEST_EPI_BIAS_DROP_RES writes per-row
// partial {sum, sum of squares} over its 64-column units: row_part [T][n_part] float2). With mean / rstd known before the
// row arrives there is no reduction over the row data any more: every 16-byte chunk is normalised and stored as soon as
// its own load returns — the kernel streams like a copy instead of load -> two warp reductions -> store.
template <int kRows>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
ln_fwd_stats_kernel(const __nv_bfloat16* __restrict__ xin, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float eps, int T, const float2* __restrict__ row_part, int n_part, __nv_bfloat16* __restrict__ y,
                    float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * kRows;
  if (t0 >= T) return;
  uint4 raw[kRows][CPL];
  float2 part[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int t = min(t0 + r, T - 1);
    part[r] = lane < n_part ? __ldg(row_part + (int64_t)t * n_part + lane) : make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CPL; ++i) raw[r][i] = __ldg(reinterpret_cast<const uint4*>(xin + (int64_t)t * H) + lane + 32 * i);
  }
  float rs[kRows], nm[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    float s1 = part[r].x, s2 = part[r].y;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {         // n_part <= 16 partials sit in lanes 0..15
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 = __shfl_sync(0xffffffffu, s1, 0);
    s2 = __shfl_sync(0xffffffffu, s2, 0);
    const float mu = s1 * (1.0f / H);
    rs[r] = rsqrtf(fmaxf(s2 * (1.0f / H) - mu * mu, 0.f) + eps);
    nm[r] = -mu * rs[r];
    if (lane == 0 && t0 + r < T) {
      if (mean) mean[t0 + r] = mu;
      if (rstd) rstd[t0 + r] = rs[r];
    }
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = 8 * (lane + 32 * i);
    float g[8], b[8];
    load8f(gamma + c, g);
    load8f(beta + c, b);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (t0 + r >= T) break;
      float x[8], o[8];
      unpack8(raw[r][i], x);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(fmaf(x[k], rs[r], nm[r]), g[k], b[k]);
      *(reinterpret_cast<uint4*>(y + (int64_t)(t0 + r) * H) + lane + 32 * i) = pack8(o);
    }
  }
}

// Persistent variant of ln_fwd_stats_kernel: ncu shows the one-shot kernels moving 27 MB of DRAM reads in 15 us (1.8 TB/s,
// the writes stay in L2) — two waves of blocks whose warps each expose a full DRAM round trip and a block turn-around.
// Here every warp keeps kDepth rows in flight for the whole kernel (row r + kDepth * W is requested as soon as row r has
// been consumed), gamma / beta sit in shared memory in a lane-major float4 layout (conflict-free), and with the row
// statistics known up front nothing waits on a reduction over the row data.
template <int kDepth>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
ln_fwd_stream_kernel(const __nv_bfloat16* __restrict__ xin, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, int T, const float2* __restrict__ row_part, int n_part, __nv_bfloat16* __restrict__ y,
                     float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_grid_sync();
  __shared__ float4 gs[CPL][2][32], bs[CPL][2][32];
  for (int idx = threadIdx.x; idx < H / 4; idx += blockDim.x) {
    const int c = 4 * idx, i = c >> 8, within = c & 255;
    gs[i][(within & 7) >> 2][within >> 3] = __ldg(reinterpret_cast<const float4*>(gamma + c));
    bs[i][(within & 7) >> 2][within >> 3] = __ldg(reinterpret_cast<const float4*>(beta + c));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int W = gridDim.x * kWarpsPerBlock;
  const int w0 = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  uint4 raw[kDepth][CPL];
  float2 part[kDepth];
  auto fetch = [&](int row, uint4 (&r)[CPL], float2& p) {
    if (row < T) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) r[i] = __ldg(reinterpret_cast<const uint4*>(xin + (int64_t)row * H) + lane + 32 * i);
      p = lane < n_part ? __ldg(row_part + (int64_t)row * n_part + lane) : make_float2(0.f, 0.f);
    }
  };
#pragma unroll
  for (int d = 0; d < kDepth; ++d) fetch(w0 + d * W, raw[d], part[d]);
  for (int t0 = w0; t0 < T; t0 += kDepth * W) {
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      const int t = t0 + d * W;
      if (t >= T) break;
      float s1 = part[d].x, s2 = part[d].y;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 = __shfl_sync(0xffffffffu, s1, 0);
      s2 = __shfl_sync(0xffffffffu, s2, 0);
      const float mu = s1 * (1.0f / H);
      const float rs = rsqrtf(fmaxf(s2 * (1.0f / H) - mu * mu, 0.f) + eps);
      const float nm = -mu * rs;
      if (lane == 0) {
        if (mean) mean[t] = mu;
        if (rstd) rstd[t] = rs;
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        float x[8], o[8];
        unpack8(raw[d][i], x);
        const float4 g0 = gs[i][0][lane], g1 = gs[i][1][lane], b0 = bs[i][0][lane], b1 = bs[i][1][lane];
        o[0] = fmaf(fmaf(x[0], rs, nm), g0.x, b0.x);
        o[1] = fmaf(fmaf(x[1], rs, nm), g0.y, b0.y);
        o[2] = fmaf(fmaf(x[2], rs, nm), g0.z, b0.z);
        o[3] = fmaf(fmaf(x[3], rs, nm), g0.w, b0.w);
        o[4] = fmaf(fmaf(x[4], rs, nm), g1.x, b1.x);
        o[5] = fmaf(fmaf(x[5], rs, nm), g1.y, b1.y);
        o[6] = fmaf(fmaf(x[6], rs, nm), g1.z, b1.z);
        o[7] = fmaf(fmaf(x[7], rs, nm), g1.w, b1.w);
        *(reinterpret_cast<uint4*>(y + (int64_t)t * H) + lane + 32 * i) = pack8(o);
      }
      fetch(t + kDepth * W, raw[d], part[d]);
    }
  }
}

// Block-level reduction of per-warp column partials (VPL float4 per lane) followed by one atomic per column per block.
__device__ __forceinline__ void block_reduce_cols_atomic(float4 (&acc)[VPL], float* __restrict__ out, float* sh /*[warps][H]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < VPL; ++i) *reinterpret_cast<float4*>(sh + warp * H + 4 * (lane + 32 * i)) = acc[i];
  __syncthreads();
  for (int c = 4 * threadIdx.x; c < H; c += 4 * blockDim.x) {   // one 4-wide vector reduction per thread
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(sh + w * H + c);
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    red_add_v4(out + c, s);
  }
}

__device__ __forceinline__ void f4_acc(float4& a, const float4& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
}

// LayerNorm backward for a row in registers: returns dx (fp32) in `g` (overwrites), accumulates dgamma/dbeta partials.
__device__ __forceinline__ void ln_bwd_row(const float4 (&xhat)[VPL], float4 (&dy)[VPL], float rstd_v,
                                           const float* __restrict__ gamma, int lane, float4 (&dgam)[VPL],
                                           float4 (&dbet)[VPL]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * i)));
    dgam[i].x += dy[i].x * xhat[i].x;
    dgam[i].y += dy[i].y * xhat[i].y;
    dgam[i].z += dy[i].z * xhat[i].z;
    dgam[i].w += dy[i].w * xhat[i].w;
    f4_acc(dbet[i], dy[i]);
    dy[i].x *= gm.x;
    dy[i].y *= gm.y;
    dy[i].z *= gm.z;
    dy[i].w *= gm.w;
    s1 += (dy[i].x + dy[i].y) + (dy[i].z + dy[i].w);
    s2 += (dy[i].x * xhat[i].x + dy[i].y * xhat[i].y) + (dy[i].z * xhat[i].z + dy[i].w * xhat[i].w);
  }
  const float c1 = warp_sum(s1) * (1.0f / H), c2 = warp_sum(s2) * (1.0f / H);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    dy[i].x = rstd_v * (dy[i].x - c1 - xhat[i].x * c2);
    dy[i].y = rstd_v * (dy[i].y - c1 - xhat[i].y * c2);
    dy[i].z = rstd_v * (dy[i].z - c1 - xhat[i].z * c2);
    dy[i].w = rstd_v * (dy[i].w - c1 - xhat[i].w * c2);
  }
}

// Residual LayerNorm backward (+ the dropout-masked gradient and the bias gradient of the dense layer before it).
// Persistent, one block (8 warps) per SM: every lane carries 72 column accumulators (dgamma, dbeta, dbias), so occupancy
// cannot supply memory-level parallelism; instead each warp streams its rows through a private 4-deep cp.async ring in
// shared memory (x row + dy row = 3 KiB per stage, 96 KiB per block): three tokens are always in flight per warp. A lane
// reads back exactly the chunks it copied, so the ring needs no cross-lane synchronisation at all.
constexpr int kLnStages = 4;
constexpr int kLnStageBytes = 2 * H * 2;                                   // x row + dy row
constexpr int kLnBwdSmem = kWarpsPerBlock * kLnStages * kLnStageBytes;     // 96 KiB (>= the 24 KiB the final reduce needs)

__device__ __forceinline__ void block_reduce_cols8_atomic(const float (&acc)[CPL][8], float* __restrict__ out, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    float* dst = sh + warp * H + 8 * (lane + 32 * i);
    *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
  __syncthreads();
  for (int c = 4 * threadIdx.x; c < H; c += 4 * blockDim.x) {
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(sh + w * H + c);
      s4.x += v.x;
      s4.y += v.y;
      s4.z += v.z;
      s4.w += v.w;
    }
    red_add_v4(out + c, s4);
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dyin, const __nv_bfloat16* __restrict__ xin, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, int T, __nv_bfloat16* __restrict__ dx,
              __nv_bfloat16* __restrict__ dxm, uint32_t thr, float scale, uint32_t seed, const uint32_t* __restrict__ salt,
              float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dbias) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  extern __shared__ __align__(16) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(ln_smem) + warp * (kLnStages * kLnStageBytes);
  float dgam[CPL][8], dbet[CPL][8], dbia[CPL][8], gam[CPL][8];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    load8f(gamma + 8 * (lane + 32 * i), gam[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) dgam[i][k] = dbet[i][k] = dbia[i][k] = 0.f;
  }
  const int tstep = gridDim.x * kWarpsPerBlock;
  const int t0 = blockIdx.x * kWarpsPerBlock + warp;
  auto issue = [&](int t, int stage) {
    if (t < T) {
      const uint32_t dst = ring + stage * kLnStageBytes + lane * 16;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        cp_async_16(dst + i * 512, reinterpret_cast<const uint4*>(xin + (int64_t)t * H) + lane + 32 * i, true);
        cp_async_16(dst + H * 2 + i * 512, reinterpret_cast<const uint4*>(dyin + (int64_t)t * H) + lane + 32 * i, true);
      }
    }
    cp_async_commit();   // committed even when empty: the group count per iteration stays uniform
  };
#pragma unroll
  for (int s = 0; s < kLnStages - 1; ++s) issue(t0 + s * tstep, s);
  float nmu = 0.f, nrs = 0.f;
  if (t0 < T) {
    nmu = mean[t0];
    nrs = rstd[t0];
  }
  int stage = 0;
  for (int t = t0; t < T; t += tstep) {
    cp_async_wait<kLnStages - 2>();   // this token's copies (the oldest group) have landed
    const uint8_t* src = ln_smem + warp * (kLnStages * kLnStageBytes) + stage * kLnStageBytes + lane * 16;
    uint4 xr[CPL], dr[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      xr[i] = *reinterpret_cast<const uint4*>(src + i * 512);
      dr[i] = *reinterpret_cast<const uint4*>(src + H * 2 + i * 512);
    }
    // refill the stage consumed in the previous iteration with the token kLnStages - 1 steps ahead
    issue(t + (kLnStages - 1) * tstep, stage == 0 ? kLnStages - 1 : stage - 1);
    const float mu = nmu, rs = nrs;
    if (t + tstep < T) {
      nmu = mean[t + tstep];
      nrs = rstd[t + tstep];
    }
    float xhat[CPL][8], dy[CPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      unpack8(xr[i], xhat[i]);
      unpack8(dr[i], dy[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xhat[i][k] = (xhat[i][k] - mu) * rs;
        dgam[i][k] = fmaf(dy[i][k], xhat[i][k], dgam[i][k]);
        dbet[i][k] += dy[i][k];
        dy[i][k] *= gam[i][k];
        s1 += dy[i][k];
        s2 = fmaf(dy[i][k], xhat[i][k], s2);
      }
    }
    const float c1 = warp_sum(s1) * (1.0f / H), c2 = warp_sum(s2) * (1.0f / H);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = 8 * (lane + 32 * i);
#pragma unroll
      for (int k = 0; k < 8; ++k) dy[i][k] = rs * (dy[i][k] - c1 - xhat[i][k] * c2);
      *(reinterpret_cast<uint4*>(dx + (int64_t)t * H) + lane + 32 * i) = pack8(dy[i]);
      if (dxm != nullptr) {
        const uint32_t base = (uint32_t)t * H + c;
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          bool k0, k1, k2, k3;
          dropout_keep4(seed, base + 4 * hq, thr, k0, k1, k2, k3);
          dy[i][4 * hq + 0] = k0 ? dy[i][4 * hq + 0] * scale : 0.f;
          dy[i][4 * hq + 1] = k1 ? dy[i][4 * hq + 1] * scale : 0.f;
          dy[i][4 * hq + 2] = k2 ? dy[i][4 * hq + 2] * scale : 0.f;
          dy[i][4 * hq + 3] = k3 ? dy[i][4 * hq + 3] * scale : 0.f;
        }
        *(reinterpret_cast<uint4*>(dxm + (int64_t)t * H) + lane + 32 * i) = pack8(dy[i]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) dbia[i][k] += dy[i][k];
    }
    stage = stage == kLnStages - 1 ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  float* sh = reinterpret_cast<float*>(ln_smem);
  block_reduce_cols8_atomic(dgam, dgamma, sh);
  block_reduce_cols8_atomic(dbet, dbeta, sh);
  if (dbias != nullptr) block_reduce_cols8_atomic(dbia, dbias, sh);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_ln_bwd_kernel(const int32_t* __restrict__ tokens, const uint8_t* __restrict__ seg, const int32_t* __restrict__ pos,
                    int T, const float* __restrict__ word, const float* __restrict__ posemb, const float* __restrict__ type,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const __nv_bfloat16* __restrict__ dyin, uint32_t thr, float scale, uint32_t seed, const uint32_t* __restrict__ salt,
                    float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, int word_pad_row, int pos_pad_row) {
  pdl_grid_sync();
  seed ^= step_salt(salt);
  __shared__ float sh[kWarpsPerBlock * H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 dgam[VPL], dbet[VPL], dty0[VPL], dty1[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) dgam[i] = dbet[i] = dty0[i] = dty1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  bool any1 = false;
  for (int t = blockIdx.x * kWarpsPerBlock + warp; t < T; t += gridDim.x * kWarpsPerBlock) {
    const int tok = tokens[t], ps = pos[t], sg = seg[t];
    const float mu = mean[t], rs = rstd[t];
    const float* w = word + (int64_t)tok * H;
    const float* p = posemb + (int64_t)ps * H;
    const float* ty = type + (int64_t)sg * H;
    float4 xhat[VPL], dy[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = 4 * (lane + 32 * i);
      const float4 a = __ldg(reinterpret_cast<const float4*>(w + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p + c));
      const float4 d = __ldg(reinterpret_cast<const float4*>(ty + c));
      xhat[i] = make_float4((a.x + d.x + b.x - mu) * rs, (a.y + d.y + b.y - mu) * rs, (a.z + d.z + b.z - mu) * rs,
                            (a.w + d.w + b.w - mu) * rs);
      dy[i] = bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(dyin + (int64_t)t * H + c)));
      if (thr) {
        const uint32_t base = (uint32_t)t * H + c;
        bool k0_, k1_, k2_, k3_;
        dropout_keep4(seed, base, thr, k0_, k1_, k2_, k3_);
        dy[i].x = k0_ ? dy[i].x * scale : 0.f;
        dy[i].y = k1_ ? dy[i].y * scale : 0.f;
        dy[i].z = k2_ ? dy[i].z * scale : 0.f;
        dy[i].w = k3_ ? dy[i].w * scale : 0.f;
      }
    }
    ln_bwd_row(xhat, dy, rs, gamma, lane, dgam, dbet);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (tok != word_pad_row) red_add_v4(dword + (int64_t)tok * H + c, dy[i]);
      if (ps != pos_pad_row) red_add_v4(dpos + (int64_t)ps * H + c, dy[i]);
      if (sg == 0) {
        f4_acc(dty0[i], dy[i]);
      } else if (sg == 1) {
        f4_acc(dty1[i], dy[i]);
        any1 = true;
      } else {
        red_add_v4(dtype + (int64_t)sg * H + c, dy[i]);
      }
    }
  }
  block_reduce_cols_atomic(dgam, dgamma, sh);
  block_reduce_cols_atomic(dbet, dbeta, sh);
  block_reduce_cols_atomic(dty0, dtype, sh);
  if (__syncthreads_or(any1 ? 1 : 0)) block_reduce_cols_atomic(dty1, dtype + H, sh);
}

// ------------------------------------------------------------------------------------------------ column sums / cast
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, int T, int N, int rows_per_block, float* __restrict__ out) {
  pdl_grid_sync();
  const int c = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, T);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = r0;
  for (; r + 8 <= r1; r += 8) {   // 8 independent 8-byte loads in flight per thread
    uint2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const uint2*>(x + (int64_t)(r + k) * N + c));
#pragma unroll
    for (int k = 0; k < 8; ++k) f4_acc(acc, bf16x4_to_f4(v[k]));
  }
  for (; r < r1; ++r) f4_acc(acc, bf16x4_to_f4(__ldg(reinterpret_cast<const uint2*>(x + (int64_t)r * N + c))));
  red_add_v4(out + c, acc);
}

__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_grid_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
      *reinterpret_cast<uint4*>(dst + i) =
          make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
  }
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int nbest_pack_batch(nbest_ctx* ctx, const int64_t* ids, const int64_t* seg_ids, int B, int S, int pos_mode,
                                int32_t* lens, int32_t* cu_seqlens, int32_t* tokens, uint8_t* seg, int32_t* pos,
                                int32_t* seq_of, uint8_t* key_valid, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, ids && lens && cu_seqlens && tokens && seg && pos && seq_of && key_valid, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0 && S > 0, "empty batch");
  NBEST_CHECK_ARG(ctx, pos_mode == 0 || pos_mode == 1, "pos_mode must be 0 (bert) or 1 (xlm-roberta)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (B + 7) / 8;
  nbest_launch(pack_lens_kernel, dim3(blocks), dim3(256), 0, s, ids, B, S, lens);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(pack_scan_kernel, dim3(1), dim3(1024), 0, s, lens, B, cu_seqlens);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(pack_scatter_kernel, dim3(blocks), dim3(256), 0, s, ids, seg_ids, B, S, pos_mode, cu_seqlens, tokens, seg, pos, seq_of, key_valid, 0);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_pack_batch_dual(nbest_ctx* ctx, const int64_t* ids_a, const int64_t* seg_a, int B_a, int S_a,
                                     const int64_t* ids_t, const int64_t* seg_t, int B_t, int S_t, int pos_mode, int32_t* lens,
                                     int32_t* cu_seqlens, int32_t* tokens, uint8_t* seg, int32_t* pos, int32_t* seq_of,
                                     uint8_t* key_valid, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, ids_a && ids_t && lens && cu_seqlens && tokens && seg && pos && seq_of && key_valid, "null pointer");
  NBEST_CHECK_ARG(ctx, B_a > 0 && S_a > 0 && B_t > 0 && S_t > 0, "empty batch");
  NBEST_CHECK_ARG(ctx, pos_mode == 0 || pos_mode == 1, "pos_mode must be 0 (bert) or 1 (xlm-roberta)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  nbest_launch(pack_lens_kernel, dim3((B_a + 7) / 8), dim3(256), 0, s, ids_a, B_a, S_a, lens);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(pack_lens_kernel, dim3((B_t + 7) / 8), dim3(256), 0, s, ids_t, B_t, S_t, lens + B_a);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(pack_scan_kernel, dim3(1), dim3(1024), 0, s, lens, B_a + B_t, cu_seqlens);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(pack_scatter_kernel, dim3((B_a + 7) / 8), dim3(256), 0, s, ids_a, seg_a, B_a, S_a, pos_mode, cu_seqlens, tokens, seg, pos, seq_of, key_valid, 0);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(pack_scatter_kernel, dim3((B_t + 7) / 8), dim3(256), 0, s, ids_t, seg_t, B_t, S_t, pos_mode, cu_seqlens, tokens, seg, pos, seq_of, key_valid, B_a);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

// dst[i, :] = src[row_idx[i], :] (bf16 -> bf16, or bf16 -> fp32 when dst_f32); one warp per row, 16-byte chunks
__global__ void rows_gather_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ row_idx, int n,
                                   void* __restrict__ dst, int dst_f32) {
  pdl_grid_sync();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const uint4* sp = reinterpret_cast<const uint4*>(src + (int64_t)row_idx[i] * H);
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const uint4 v = __ldg(sp + lane + 32 * c);
    if (!dst_f32) {
      *(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + (int64_t)i * H) + lane + 32 * c) = v;
    } else {
      float f[8];
      unpack8(v, f);
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + (int64_t)i * H + 8 * (lane + 32 * c));
      o[0] = make_float4(f[0], f[1], f[2], f[3]);
      o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}
// dst[row_idx[i], :] = src[i, :] (bf16); the caller zero-fills dst
__global__ void rows_scatter_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ row_idx, int n,
                                    __nv_bfloat16* __restrict__ dst) {
  pdl_grid_sync();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const uint4* sp = reinterpret_cast<const uint4*>(src + (int64_t)i * H);
  uint4* dp = reinterpret_cast<uint4*>(dst + (int64_t)row_idx[i] * H);
#pragma unroll
  for (int c = 0; c < CPL; ++c) dp[lane + 32 * c] = __ldg(sp + lane + 32 * c);
}

// ---- row-sparse exchange of the word-embedding gradient (data parallel, large vocabularies) ----------------------------
// Only rows whose token occurred in this step's batch carry a non-zero gradient (<= T of XLM-R's 250,002 rows): the ranks
// agree on the union of touched rows (flags, MAX-reduced), exchange those rows as one dense [n, 768] all-reduce and
// scatter the sums back. Ascending row order on every rank -> identical summation order -> replicas stay bit-identical.
__global__ void rows_mark_kernel(const int32_t* __restrict__ tokens, int T, int32_t* __restrict__ flags) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) flags[tokens[t]] = 1;
}
// rows[0 .. count) = ascending indices r != skip_row with flags[r] != 0 (single block: chunked ballot scan)
__global__ void __launch_bounds__(1024)
rows_compact_kernel(const int32_t* __restrict__ flags, int n_rows, int skip_row, int32_t* __restrict__ rows, int32_t* __restrict__ count) {
  pdl_grid_sync();
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int r0 = 0; r0 < n_rows; r0 += 1024) {
    const int r = r0 + threadIdx.x;
    const bool on = r < n_rows && r != skip_row && flags[r] != 0;
    const uint32_t m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (on) rows[off + __popc(m & ((1u << lane) - 1u))] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
      base_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base_s;
}
// fp32 [*, 768] rows: gather dst[i] = src[rows[i]], or scatter dst[rows[i]] = src[i]; one warp per row, float4 chunks
__global__ void rows_move_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows, int n,
                                     float* __restrict__ dst, int scatter) {
  pdl_grid_sync();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = rows[i];
  const float4* sp = reinterpret_cast<const float4*>(src + (scatter ? (int64_t)i : r) * H);
  float4* dp = reinterpret_cast<float4*>(dst + (scatter ? r : (int64_t)i) * H);
#pragma unroll
  for (int c = 0; c < VPL; ++c) dp[lane + 32 * c] = __ldg(sp + lane + 32 * c);
}

extern "C" int nbest_rows_touched(nbest_ctx* ctx, const int32_t* tokens, int T, int n_rows, int skip_row, int32_t* flags,
                                  int phase, int32_t* rows, int32_t* count, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, flags && n_rows > 0, "null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (phase == 0) {           // flags = 0; flags[tokens[t]] = 1
    NBEST_CHECK_ARG(ctx, tokens && T >= 0, "null pointer");
    NBEST_CHECK_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int32_t) * (size_t)n_rows, s));
    if (T > 0) {
      nbest_launch(rows_mark_kernel, dim3((T + 255) / 256), dim3(256), 0, s, tokens, T, flags);
      NBEST_CHECK_LAUNCH(ctx);
    }
  } else {                    // rows / count from the (reduced) flags
    NBEST_CHECK_ARG(ctx, rows && count, "null pointer");
    nbest_launch(rows_compact_kernel, dim3(1), dim3(1024), 0, s, flags, n_rows, skip_row, rows, count);
    NBEST_CHECK_LAUNCH(ctx);
  }
  return NBEST_OK;
}

extern "C" int nbest_rows_move_f32(nbest_ctx* ctx, const float* src, const int32_t* rows, int n, int hidden, float* dst,
                                   int scatter, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, src && rows && dst, "null pointer");
  if (n <= 0) return NBEST_OK;
  nbest_launch(rows_move_f32_kernel, dim3((n + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), src, rows, n, dst, scatter);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_rows_gather(nbest_ctx* ctx, const void* src_bf16, const int32_t* row_idx, int n, int hidden, void* dst,
                                 int dst_is_f32, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, src_bf16 && row_idx && dst, "null pointer");
  if (n <= 0) return NBEST_OK;
  nbest_launch(rows_gather_kernel, dim3((n + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(src_bf16), row_idx, n, dst, dst_is_f32);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_rows_scatter(nbest_ctx* ctx, const void* src_bf16, const int32_t* row_idx, int n, int T, int hidden,
                                  void* dst_bf16, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, src_bf16 && row_idx && dst_bf16 && T >= 0, "null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  NBEST_CHECK_CUDA(ctx, cudaMemsetAsync(dst_bf16, 0, (size_t)T * H * 2, s));
  if (n <= 0) return NBEST_OK;
  nbest_launch(rows_scatter_kernel, dim3((n + 7) / 8), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(src_bf16), row_idx, n,
                                                  reinterpret_cast<__nv_bfloat16*>(dst_bf16));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_zero(nbest_ctx* ctx, void* ptr, int64_t nbytes, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, ptr && nbytes >= 0, "null pointer");
  NBEST_CHECK_CUDA(ctx, cudaMemsetAsync(ptr, 0, (size_t)nbytes, reinterpret_cast<cudaStream_t>(stream)));
  return NBEST_OK;
}

extern "C" int nbest_pack_hyp_ids(nbest_ctx* ctx, const int32_t* tokens, const int32_t* cu_seqlens, int B, int sep_id,
                                  uint8_t* hyp_id, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, tokens && cu_seqlens && hyp_id, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0, "empty batch");
  nbest_launch(pack_hyp_ids_kernel, dim3((B + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), tokens, cu_seqlens, B, sep_id, hyp_id);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_embed_ln_fwd(nbest_ctx* ctx, const int32_t* tokens, const uint8_t* seg, const int32_t* pos, int T,
                                  const float* word, const float* posemb, const float* type, const float* gamma,
                                  const float* beta, float eps, int hidden, void* y_bf16, float* mean, float* rstd,
                                  float p_drop, uint32_t seed, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, tokens && seg && pos && word && posemb && type && gamma && beta && y_bf16 && mean && rstd, "null pointer");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  NBEST_CHECK_ARG(ctx, p_drop <= 0.f || (int64_t)T * 768 < (1LL << 32), "dropout counter t * 768 + c would wrap 32 bits");
  if (T <= 0) return NBEST_OK;
  nbest_launch(embed_ln_fwd_kernel, dim3((T + kWarpsPerBlock - 1) / kWarpsPerBlock), dim3(kWarpsPerBlock * 32), 0, reinterpret_cast<cudaStream_t>(stream), 
      tokens, seg, pos, T, word, posemb, type, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y_bf16), mean, rstd,
      drop_threshold(p_drop), 1.0f / (1.0f - p_drop), seed, nbest_salt(ctx));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_embed_ln_bwd(nbest_ctx* ctx, const int32_t* tokens, const uint8_t* seg, const int32_t* pos, int T,
                                  const float* word, const float* posemb, const float* type, const float* gamma,
                                  const float* mean, const float* rstd, int hidden, const void* dy_bf16, float p_drop,
                                  uint32_t seed, float* dword, float* dpos, float* dtype, float* dgamma, float* dbeta,
                                  int word_pad_row, int pos_pad_row, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, tokens && seg && pos && word && posemb && type && gamma && mean && rstd && dy_bf16, "null pointer");
  NBEST_CHECK_ARG(ctx, dword && dpos && dtype && dgamma && dbeta, "null gradient pointer");
  if (T <= 0) return NBEST_OK;
  int blocks = (T + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (blocks > ctx->num_sms) blocks = ctx->num_sms;   // one block per SM (register-bound); fewer blocks = fewer column atomics
  nbest_launch(embed_ln_bwd_kernel, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, reinterpret_cast<cudaStream_t>(stream), 
      tokens, seg, pos, T, word, posemb, type, gamma, mean, rstd, reinterpret_cast<const __nv_bfloat16*>(dy_bf16),
      drop_threshold(p_drop), 1.0f / (1.0f - p_drop), seed, nbest_salt(ctx), dword, dpos, dtype, dgamma, dbeta, word_pad_row, pos_pad_row);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_ln_fwd_stats(nbest_ctx* ctx, const void* x_bf16, const float* gamma, const float* beta, float eps,
                                  int T, int hidden, const float* row_partials, int n_partials, void* y_bf16, float* mean,
                                  float* rstd, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, x_bf16 && gamma && beta && y_bf16, "null pointer");
  NBEST_CHECK_ARG(ctx, row_partials == nullptr || (n_partials >= 1 && n_partials <= 16), "n_partials must be in [1, 16]");
  if (T <= 0) return NBEST_OK;
  auto* xi = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  auto* yo = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static int rows = 0;
  if (rows == 0) {
    const char* e = getenv("NBEST_LN_FWD_ROWS");
    rows = e ? atoi(e) : 2;
    if (rows != 1 && rows != 2 && rows != 4) rows = 2;
  }
  const int rows_per_block = kWarpsPerBlock * rows;
  const int blocks = (T + rows_per_block - 1) / rows_per_block;
  static int depth = -1;
  if (depth < 0) depth = getenv("NBEST_LN_FWD_STREAM") ? atoi(getenv("NBEST_LN_FWD_STREAM")) : 1;
  if (row_partials != nullptr && depth > 0) {
    int pb = (T + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (pb > 4 * ctx->num_sms) pb = 4 * ctx->num_sms;
    const auto* rp = reinterpret_cast<const float2*>(row_partials);
    if (depth == 3) nbest_launch(ln_fwd_stream_kernel<3>, dim3(pb), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, rp, n_partials, yo, mean, rstd);
    else if (depth == 1) nbest_launch(ln_fwd_stream_kernel<1>, dim3(pb), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, rp, n_partials, yo, mean, rstd);
    else nbest_launch(ln_fwd_stream_kernel<2>, dim3(pb), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, rp, n_partials, yo, mean, rstd);
  } else if (row_partials != nullptr) {
    const auto* rp = reinterpret_cast<const float2*>(row_partials);
    if (rows == 4) nbest_launch(ln_fwd_stats_kernel<4>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, rp, n_partials, yo, mean, rstd);
    else if (rows == 1) nbest_launch(ln_fwd_stats_kernel<1>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, rp, n_partials, yo, mean, rstd);
    else nbest_launch(ln_fwd_stats_kernel<2>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, rp, n_partials, yo, mean, rstd);
  } else if (rows == 1) {
    nbest_launch(ln_fwd_kernel<1>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, yo, mean, rstd);
  } else if (rows == 2) {
    nbest_launch(ln_fwd_kernel<2>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, yo, mean, rstd);
  } else {
    nbest_launch(ln_fwd_kernel<4>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, xi, gamma, beta, eps, T, yo, mean, rstd);
  }
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_ln_fwd(nbest_ctx* ctx, const void* x_bf16, const float* gamma, const float* beta, float eps, int T,
                            int hidden, void* y_bf16, float* mean, float* rstd, void* stream) {
  return nbest_ln_fwd_stats(ctx, x_bf16, gamma, beta, eps, T, hidden, nullptr, 0, y_bf16, mean, rstd, stream);
}

extern "C" int nbest_ln_bwd(nbest_ctx* ctx, const void* dy_bf16, const void* x_bf16, const float* mean, const float* rstd,
                            const float* gamma, int T, int hidden, void* dx_bf16, void* dx_masked_bf16, float p_drop,
                            uint32_t seed, float* dgamma, float* dbeta, float* dbias, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, hidden == H, "hidden must be 768");
  NBEST_CHECK_ARG(ctx, dy_bf16 && x_bf16 && mean && rstd && gamma && dx_bf16 && dgamma && dbeta, "null pointer");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  NBEST_CHECK_ARG(ctx, p_drop <= 0.f || (int64_t)T * 768 < (1LL << 32), "dropout counter t * 768 + c would wrap 32 bits");
  NBEST_CHECK_ARG(ctx, !(p_drop > 0.f) || dx_masked_bf16, "p_drop > 0 needs dx_masked");
  if (T <= 0) return NBEST_OK;
  int blocks = (T + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (blocks > ctx->num_sms) blocks = ctx->num_sms;   // one block per SM (register-bound); fewer blocks = fewer column atomics
  static bool attr_dev[64] = {};   // per device: cudaFuncSetAttribute applies to the current device only
  bool& attr = attr_dev[ctx->device & 63];
  if (!attr) {
    NBEST_CHECK_CUDA(ctx, cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnBwdSmem));
    attr = true;
  }
  nbest_launch(ln_bwd_kernel, dim3(blocks), dim3(kWarpsPerBlock * 32), kLnBwdSmem, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(dy_bf16), reinterpret_cast<const __nv_bfloat16*>(x_bf16), mean, rstd, gamma, T,
      reinterpret_cast<__nv_bfloat16*>(dx_bf16), p_drop > 0.f ? reinterpret_cast<__nv_bfloat16*>(dx_masked_bf16) : nullptr,
      drop_threshold(p_drop), 1.0f / (1.0f - p_drop), seed, nbest_salt(ctx), dgamma, dbeta, dbias);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_colsum_bf16(nbest_ctx* ctx, const void* x_bf16, int T, int N, float* out, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, x_bf16 && out, "null pointer");
  NBEST_CHECK_ARG(ctx, N > 0 && N % 4 == 0, "N must be a multiple of 4");
  if (T <= 0) return NBEST_OK;
  const int threads = 128;
  const int gx = (N / 4 + threads - 1) / threads;
  int gy = (8 * ctx->num_sms + gx - 1) / gx;
  int rows_per_block = (T + gy - 1) / gy;
  if (rows_per_block < 32) rows_per_block = 32;
  gy = (T + rows_per_block - 1) / rows_per_block;
  nbest_launch(colsum_kernel, dim3(dim3(gx, gy)), dim3(threads), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), T, N, rows_per_block, out);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_cast_f32_bf16(nbest_ctx* ctx, const float* src, void* dst_bf16, int64_t n, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, src && dst_bf16, "null pointer");
  NBEST_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0,
                  "buffers must be 16-byte aligned");
  if (n <= 0) return NBEST_OK;
  int64_t blocks = (n / 8 + 255) / 256;
  if (blocks > 8 * ctx->num_sms) blocks = 8 * ctx->num_sms;
  if (blocks < 1) blocks = 1;
  nbest_launch(cast_kernel, dim3((int)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), n);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}
