// Masked self-attention over packed short sequences on the Blackwell tensor path: tcgen05.mma with S / O (forward) and
// S / dP / dQ / dK / dV (backward) in TMEM, Q / K / V / dO tiles by TMA, persistent warp-specialised CTAs.
//
// Replaces BertSelfAttention's scores / softmax / dropout / context (transformers modeling_bert.py:115-140,192-205) and
// its autograd under the key mask `attention_mask = input_ids > 0` of models/model.py:43,45, like attention.cu, for the
// sequences that dominate the n-best workload: <= 128 tokens (all of DSTC2 5-best training; ~90 % of 10-best inference).
//
// DSTC2 n-best sequences are 20-130 tokens, far below a tensor-core tile, so WHOLE sequences are packed greedily into
// 128-row TILES of the packed token axis (nbest_attn_plan: consecutive sequences while they fit, a tile never straddles
// the ASR / transcript boundary). One work item = (tile, head):
//     S = Q K^T            one 128 x 128 x 64 tcgen05.mma chain over the tile's tokens (queries AND keys are the tile)
//     block-diagonal mask  key column c is visible to query row r iff both belong to the same sequence (per-row [lo, hi)
//                          column range from row_bounds) and key_valid[c] (XLM-R quirk: <s> = 0 is masked as a key)
//     softmax / dropout    one thread per query row straight out of TMEM (tcgen05.ld), P to shared memory as bf16 in the
//                          K-major SWIZZLE_128B operand layout
//     O = P V              128 x 64 x 128 chain, V consumed MN-major from the very tile TMA delivered
// The masked-out part of the 128 x 128 score tile is wasted tensor work, but the tensor pipe is idle anyway: per layer the
// op moves ~110 MB (forward) / ~130 MB (backward) for ~1.7 / ~4.2 GFLOP of useful math, so HBM is the roof and the
// kernel is organised around keeping TMA loads in flight (3-stage input ring, items of different heads back to back).
// Dropout keeps the index map of attention.cu / ptx.cuh (counter = (head, global query token, key index in sequence)),
// so masks are bit-identical between the two implementations and between forward and backward.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

using namespace nbest;

namespace {

constexpr int kTile = 128;                      // tokens per tile (queries = keys)
constexpr int kD = 64;                          // head dim
constexpr uint32_t kMatBytes = kTile * kD * 2;  // one 128 x 64 bf16 operand tile = 16 KiB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------- plan
// Greedy packing of whole sequences into 128-row tiles. tiles[i] = {first packed row, number of rows owned}. Sequences
// longer than 128 tokens are left to the block-loop kernels of attention.cu (counts[2] = how many there are).
// counts[0] = tiles in total, counts[1] = tiles that cover sequences [0, break_at) (the gradient-carrying ASR prefix).
constexpr int kPlanSmemSeqs = 4096;     // sequences whose plan is built in shared memory (larger batches: serial global scan)
__global__ void __launch_bounds__(256)
attn_plan_kernel(const int32_t* __restrict__ cu_g, int B, int break_at, int2* __restrict__ tiles, int32_t* __restrict__ counts) {
  pdl_grid_sync();
  __shared__ int32_t cu_s[kPlanSmemSeqs + 1];
  __shared__ int32_t nxt_s[kPlanSmemSeqs];   // first sequence that no longer fits a tile opened at sequence s (s + 1 for long ones)
  if (B <= kPlanSmemSeqs) {
    // The greedy packing is a sequential chain, but "where does a tile opened at sequence s end" is a binary search on
    // the prefix sums, independent per s: all of them in parallel, then one thread follows ~T/128 pointers.
    for (int i = threadIdx.x; i <= B; i += blockDim.x) cu_s[i] = cu_g[i];
    __syncthreads();
    for (int s = threadIdx.x; s < B; s += blockDim.x) {
      const int limit = s < break_at ? break_at : B;
      int e = s + 1;
      if (cu_s[s + 1] - cu_s[s] <= kTile) {
        int lo = s + 1, hi = limit;                  // largest e in [s + 1, limit] with cu[e] - cu[s] <= 128
        const int base = cu_s[s];
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (cu_s[mid] - base <= kTile) lo = mid; else hi = mid - 1;
        }
        e = lo;
      }
      nxt_s[s] = e;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int n = 0, n_break = break_at == 0 ? 0 : -1, n_long = 0, s = 0;
    while (s < B) {
      const int s0 = cu_s[s];
      if (cu_s[s + 1] - s0 > kTile) {
        ++n_long;
        ++s;
      } else {
        const int e = nxt_s[s];
        const int rows = cu_s[e] - s0;
        if (rows > 0) tiles[n++] = make_int2(s0, rows);
        s = e;
      }
      if (s == break_at) n_break = n;
    }
    counts[0] = n;
    counts[1] = n_break < 0 ? n : n_break;
    counts[2] = n_long;
    return;
  }
  if (threadIdx.x != 0) return;
  const int32_t* cu = cu_g;
  int n = 0, n_break = -1, n_long = 0;
  int start = -1, rows = 0;
  for (int b = 0; b <= B; ++b) {
    if (b == break_at && n_break < 0) {
      if (start >= 0) tiles[n++] = make_int2(start, rows);
      start = -1;
      rows = 0;
      n_break = n;
    }
    if (b == B) break;
    const int s0 = cu[b], L = cu[b + 1] - s0;
    if (L <= 0) continue;
    if (L > kTile) {
      if (start >= 0) tiles[n++] = make_int2(start, rows);
      start = -1;
      rows = 0;
      ++n_long;
      continue;
    }
    if (start >= 0 && rows + L > kTile) {
      tiles[n++] = make_int2(start, rows);
      start = -1;
      rows = 0;
    }
    if (start < 0) start = s0;
    rows += L;
  }
  if (start >= 0) tiles[n++] = make_int2(start, rows);
  counts[0] = n;
  counts[1] = n_break < 0 ? n : n_break;
  counts[2] = n_long;
}

// row_bounds[t] = {first row, one-past-last row} of the sequence token t belongs to
__global__ void attn_row_bounds_kernel(const int32_t* __restrict__ cu, const int32_t* __restrict__ seq_of, int T,
                                       int2* __restrict__ row_bounds) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int s = seq_of[t];
  row_bounds[t] = make_int2(cu[s], cu[s + 1]);
}

// ------------------------------------------------------------------------------------------------------- bit helpers
__device__ __forceinline__ uint32_t ones_below(int x, int w) {   // bits of 32-column word w that are below column x
  const int r = x - 32 * w;
  return r >= 32 ? 0xFFFFFFFFu : (r <= 0 ? 0u : ((1u << r) - 1u));
}
// Keep-bits of keys 16 g16 .. 16 g16 + 15 of one query row (bit jj = key 16 g16 + jj), from the shared counter-based hash:
// one hash per quad {2q, 2q+1, 2q+8, 2q+9} of the group (ptx.cuh attn_quad / attn_lane).
__device__ __forceinline__ uint32_t dropout_group_bits(uint32_t seed, uint32_t row_base, int g16, uint32_t thr) {
  const uint32_t k = 0x10000u - thr;       // lane + k has bit 16 set  <=>  lane >= thr   (lanes are 16-bit)
  uint32_t b16 = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint2 h = dropout_quad(seed, row_base + (uint32_t)(g16 * 4 + q));
    const uint32_t pair = ((((h.x & 0xFFFFu) + k) >> 16) & 1u) | (((((h.x >> 16) + k) >> 16) & 1u) << 1) |
                          (((((h.y & 0xFFFFu) + k) >> 16) & 1u) << 8) | (((((h.y >> 16) + k) >> 16) & 1u) << 9);
    b16 |= pair << (2 * q);
  }
  return b16;
}
// Masks of one query row over the 32 key columns [32 w, 32 w + 32) of its tile: vis = visible keys (same sequence and key
// valid), keep = dropout keep bits moved from key-index space (j = column - lo) into column space. lo / hi are the
// tile-relative column bounds of the row's sequence (lo == hi for rows the tile does not own). Only the <= 3 hash groups
// that overlap the word are evaluated.
template <bool kDrop>
__device__ __forceinline__ void row_masks_word(int lo, int hi, const uint32_t* kvbits, uint32_t seed, uint32_t row_base,
                                               uint32_t thr, int w, uint32_t& vis, uint32_t& keep) {
  vis = ones_below(hi, w) & ~ones_below(lo, w);
  if (kvbits) vis &= kvbits[w];
  keep = 0xFFFFFFFFu;
  if (kDrop) {
    if (vis != 0u) {
      const int L = hi - lo;
      const int jlo = 32 * w - lo;                    // key index of column 32 w (negative: the sequence starts inside the word)
      const int gfirst = jlo <= 0 ? 0 : (jlo >> 4);
      uint64_t win = 0;                               // bit i <-> key 16 gfirst + i
#pragma unroll
      for (int gi = 0; gi < 3; ++gi) {
        const int g16 = gfirst + gi;
        if (g16 * 16 < L) win |= (uint64_t)dropout_group_bits(seed, row_base, g16, thr) << (16 * gi);
      }
      const int sh = jlo - 16 * gfirst;               // in [0, 16) when jlo >= 0, else jlo itself (> -32 because vis != 0)
      keep = sh >= 0 ? (uint32_t)(win >> sh) : (uint32_t)(win << (-sh));
    }
  }
}

// address of 16-byte chunk `chunk16` (8 bf16 = 8 key columns) of row `row` inside a [128 rows][128 columns] bf16 tile
// stored as two K-major SWIZZLE_128B atoms of 64 columns (16 KiB each)
__device__ __forceinline__ uint32_t p_tile_off(int row, int chunk16) {
  return (uint32_t)((chunk16 >> 3) * (int)kMatBytes + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}

// ------------------------------------------------------------------------------------------------------------ forward
// Warp roles (18 warps): two softmax groups of 8 warps ping-pong over consecutive items — thread = (query row, 64-column
// half), the halves exchange row maximum and row sum through shared memory — warp 16 = TMA producer, warp 17 = MMA issuer.
namespace fwd {
constexpr int kStages = 3;                      // input ring: {Q, K, V} tiles of one item per stage
constexpr int kGroups = 2;
constexpr int kGroupWarps = 8;
constexpr int kGroupThreads = 32 * kGroupWarps;
constexpr int kProducerWarp = kGroups * kGroupWarps, kMmaWarp = kProducerWarp + 1;
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr uint32_t kStageBytes = 3 * kMatBytes;
constexpr uint32_t kPBytes = 2 * kMatBytes;     // P tile: 128 x 128 bf16
constexpr uint32_t kPOff = kStages * kStageBytes;
constexpr uint32_t kXchOff = kPOff + kGroups * kPBytes;     // float xch[group][parity][max | sum][half][128 rows]
constexpr uint32_t kXchBytes = kGroups * 2 * 2 * 2 * kTile * 4;
constexpr uint32_t kAuxOff = kXchOff + kXchBytes;           // kvbits[2 groups][2 parities][4 words]
constexpr uint32_t kBarOff = kAuxOff + 64;
constexpr uint32_t kSmemBytes = kBarOff + 256;
constexpr uint32_t kTmemCols = 512;             // per group: S at +0 (128 columns), O at +128 (64 columns); group stride 256
}  // namespace fwd

struct FwdArgs {
  const int2* tiles;
  const int32_t* counts;
  int count_idx;
  const int2* row_bounds;
  const uint8_t* key_valid;
  int heads, T;
  __nv_bfloat16* out;
  float* lse;
  float scale, rscale;
  uint32_t thr, seed;
  const uint32_t* salt;   // per-step dropout salt (ptx.cuh step_salt)
};

template <bool kDrop>
__global__ void __launch_bounds__(fwd::kThreads, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const FwdArgs a) {
  pdl_grid_sync();
  const uint32_t seed_eff = kDrop ? (a.seed ^ step_salt(a.salt)) : 0u;
  using namespace fwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  float* xch_s = reinterpret_cast<float*>(smem + kXchOff);
  uint32_t* kvbits_s = reinterpret_cast<uint32_t*>(smem + kAuxOff);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* in_empty = in_full + kStages;
  uint64_t* s_full = in_empty + kStages;
  uint64_t* p_full = s_full + kGroups;
  uint64_t* o_full = p_full + kGroups;
  uint64_t* o_free = o_full + kGroups;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + kGroups);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = a.counts[a.count_idx];
  const int total = n_tiles * a.heads;
  const int hd = a.heads * kD;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    for (int g = 0; g < kGroups; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], kGroupThreads);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], kGroupThreads);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    int n = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x, ++n) {
      const int st = n % kStages;
      const uint32_t ph = (uint32_t)(n / kStages) & 1u;
      mbar_wait(&in_empty[st], ph ^ 1u);
      if (lane == 0) {
        const int tile = it / a.heads, h = it - tile * a.heads;
        const int row0 = a.tiles[tile].x;
        uint8_t* dst = smem + st * kStageBytes;
        mbar_arrive_expect_tx(&in_full[st], kStageBytes);
        tma_load_2d(&tmQKV, &in_full[st], dst, h * kD, row0);                        // Q
        tma_load_2d(&tmQKV, &in_full[st], dst + kMatBytes, hd + h * kD, row0);       // K
        tma_load_2d(&tmQKV, &in_full[st], dst + 2 * kMatBytes, 2 * hd + h * kD, row0);   // V
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(kTile, kTile, 0, 0);   // S[128 q x 128 keys] = Q (K-major) x K^T (K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(kTile, kD, 0, 1);      // O[128 q x 64]      = P (K-major) x V (MN-major)
    int n_items = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x) ++n_items;
    auto issue_s = [&](int n) {
      const int st = n % kStages, g = n & 1;
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem + st * kStageBytes), sk = sq + kMatBytes;
        const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(d, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[g]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int n) {
      const int st = n % kStages, g = n & 1;
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sp = smem_u32(smem + kPOff + g * kPBytes);
        const uint32_t sv = smem_u32(smem + st * kStageBytes + 2 * kMatBytes);
        const uint32_t d = tmem_base + (uint32_t)g * 256u + 128u;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)
          umma_bf16(d, umma_smem_desc(sp + (k >> 2) * kMatBytes + (k & 3) * 32, 16, 1024),
                    umma_smem_desc(sv + k * 2048, kMatBytes, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(&o_full[g]);
        umma_commit(&in_empty[st]);
      }
      __syncwarp();
    };
    // Fixed order S(n+1), PV(n) with blocking waits. S(n+1) is deliberately held until the epilogue of item n-1 has left
    // the group's TMEM columns (o_free), although only its S columns would have to be free: measured on the bench shape,
    // issuing it as soon as P(n-1) is complete — S MMAs then write the group's TMEM while its epilogue still reads O —
    // costs 4 us per layer (49.4 vs 45.2 us), and an event-driven issuer that polls both conditions is no better.
    for (int n = 0; n < n_items; ++n) {
      if (n == 0) {
        mbar_wait(&in_full[0], 0);
        issue_s(0);
      }
      if (n + 1 < n_items) {
        mbar_wait(&in_full[(n + 1) % kStages], (uint32_t)((n + 1) / kStages) & 1u);
        mbar_wait(&o_free[(n + 1) & 1], ((uint32_t)((n + 1) >> 1) & 1u) ^ 1u);
        issue_s(n + 1);
      }
      mbar_wait(&p_full[n & 1], (uint32_t)(n >> 1) & 1u);
      issue_pv(n);
    }
  } else {
    // ------------------------------------------------------------------ softmax groups: thread = (query row, 64-column half)
    const int g = warp >> 3, q = warp & 3, hf = (warp >> 2) & 1;
    const int r = q * 32 + lane;
    const float sl2 = a.scale * kLog2e;
    uint8_t* p_s = smem + kPOff + g * kPBytes;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
    // this group's items: every second one of the CTA's; the next item's tile / sequence bounds are fetched one item ahead
    auto fetch = [&](int it, int2& tl, int2& bd) {
      tl = make_int2(0, 0);
      bd = make_int2(0, 0);
      if (it < total) {
        tl = a.tiles[it / a.heads];
        if (r < tl.y) bd = a.row_bounds[tl.x + r];
      }
    };
    int u = 0;
    int2 tl_n, bd_n;
    fetch(blockIdx.x + g * gridDim.x, tl_n, bd_n);
    for (int it = blockIdx.x + g * gridDim.x; it < total; it += 2 * gridDim.x) {
      const int tile = it / a.heads, h = it - tile * a.heads;
      const int2 tl = tl_n, bd = bd_n;
      fetch(it + 2 * gridDim.x, tl_n, bd_n);
      const int t = tl.x + r;
      const bool row_ok = r < tl.y;
      const int lo = row_ok ? bd.x - tl.x : 0, hi = row_ok ? bd.y - tl.x : 0;
      float* xmax = xch_s + ((g * 2 + (u & 1)) * 2 + 0) * 2 * kTile;     // [half][row]
      float* xsum = xch_s + ((g * 2 + (u & 1)) * 2 + 1) * 2 * kTile;
      const int pair_bar = 1 + g * 4 + q;          // the two warps that share this row quarter (64 threads)
      const uint32_t* kvb = nullptr;
      if (a.key_valid != nullptr) {
        const bool kv = t < a.T && a.key_valid[t] != 0;
        const uint32_t word = __ballot_sync(0xffffffffu, kv);
        uint32_t* dst = kvbits_s + (g * 2 + (u & 1)) * 4;
        if (lane == 0) dst[q] = word;              // (both column halves of a row quarter write the same word)
        named_bar_sync(9 + g, kGroupThreads);
        kvb = dst;
      }
      uint32_t vis[2], keep[2];
      const uint32_t row_base = attn_quad_row(h, a.T, t);
      row_masks_word<kDrop>(lo, hi, kvb, seed_eff, row_base, a.thr, 2 * hf, vis[0], keep[0]);
      row_masks_word<kDrop>(lo, hi, kvb, seed_eff, row_base, a.thr, 2 * hf + 1, vis[1], keep[1]);
      bool need[2];
      need[0] = __any_sync(0xffffffffu, vis[0] != 0u);
      need[1] = __any_sync(0xffffffffu, vis[1] != 0u);

      mbar_wait(&s_full[g], (uint32_t)u & 1u);
      tc_fence_after();
      // pass 1: an upper bound of the row maximum — the maximum over ALL columns of the chunks this warp touches (keys of
      // neighbouring sequences included: finite scores of the same magnitude). Any bound >= the true maximum gives the
      // same softmax; the exact masked maximum is only recomputed in the (never observed) case that the bound is so far
      // above the row's own scores that the row sum underflows.
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (!need[cc]) continue;
        uint32_t v[32];
        tmem_ld_32x32(lane_base + (2 * hf + cc) * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      xmax[hf * kTile + r] = mx;
      named_bar_sync(pair_bar, 64);
      mx = fmaxf(mx, xmax[(hf ^ 1) * kTile + r]);
      float m2 = (mx == -INFINITY) ? 0.f : mx * sl2;
      // pass 2: P = exp2(S scale log2e - m2); row sum over the visible keys; kept P (the 1 / (1 - p) of dropout is folded
      // into the output scale) to shared memory as bf16
      auto pass2 = [&](float m2_) -> float {
        float ls = 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * hf + cc;
          if (!need[cc]) {
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) *reinterpret_cast<uint4*>(p_s + p_tile_off(r, c * 4 + s4)) = make_uint4(0, 0, 0, 0);
            continue;
          }
          uint32_t v[32];
          tmem_ld_32x32(lane_base + c * 32, v);
          tmem_ld_wait();
          const uint32_t m = vis[cc], kp = keep[cc] & m;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -m2_));
            const float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -m2_));
            if ((m >> j) & 1u) ls += e0;
            if ((m >> (j + 1)) & 1u) ls += e1;
            pk[j >> 1] = pack_bf16x2((kp >> j) & 1u ? e0 : 0.f, (kp >> (j + 1)) & 1u ? e1 : 0.f);
          }
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4)
            *reinterpret_cast<uint4*>(p_s + p_tile_off(r, c * 4 + s4)) = make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]);
        }
        return ls;
      };
      float l = pass2(m2);
      xsum[hf * kTile + r] = l;
      named_bar_sync(pair_bar, 64);
      l += xsum[(hf ^ 1) * kTile + r];
      // (both threads of a row hold the same l — a + b == b + a — and both warps of the pair hold the same rows, so the
      //  vote below is uniform across the pair without further communication)
      if (__any_sync(0xffffffffu, hi > lo && !(l > 1e-30f))) {
        float mexact = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (!need[cc]) continue;
          uint32_t v[32];
          tmem_ld_32x32(lane_base + (2 * hf + cc) * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if ((vis[cc] >> j) & 1u) mexact = fmaxf(mexact, __uint_as_float(v[j]));
        }
        xmax[hf * kTile + r] = mexact;
        named_bar_sync(pair_bar, 64);
        mx = fmaxf(mexact, xmax[(hf ^ 1) * kTile + r]);
        m2 = (mx == -INFINITY) ? 0.f : mx * sl2;
        l = pass2(m2);
        xsum[hf * kTile + r] = l;
        named_bar_sync(pair_bar, 64);
        l += xsum[(hf ^ 1) * kTile + r];
      }
      fence_proxy_async_smem();      // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      // O = P V: 32 of the 64 output columns per thread
      mbar_wait(&o_full[g], (uint32_t)u & 1u);
      tc_fence_after();
      const float inv = l > 0.f ? a.rscale / l : 0.f;
      {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + 128 + hf * 32, v);
        tmem_ld_wait();
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)t * hd + h * kD + hf * 32);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4)
            dst[s4] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * s4]) * inv, __uint_as_float(v[8 * s4 + 1]) * inv),
                                 pack_bf16x2(__uint_as_float(v[8 * s4 + 2]) * inv, __uint_as_float(v[8 * s4 + 3]) * inv),
                                 pack_bf16x2(__uint_as_float(v[8 * s4 + 4]) * inv, __uint_as_float(v[8 * s4 + 5]) * inv),
                                 pack_bf16x2(__uint_as_float(v[8 * s4 + 6]) * inv, __uint_as_float(v[8 * s4 + 7]) * inv));
        }
      }
      if (row_ok && hf == 0) a.lse[(int64_t)h * a.T + t] = l > 0.f ? mx * a.scale + logf(l) : 0.f;
      tc_fence_before();
      mbar_arrive(&o_free[g]);
      ++u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, fwd::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------------------------- backward
// One item = (tile, head), five tensor-core products, every operand tile staged once:
//     S  = Q K^T, dP = dO V^T                 (128 x 128 x 64 each, TMEM buffer set b: S at +0, dP at +128)
//     P  = exp2(S scale log2e - lse), Pd = keep o P, dS' = P o (keep o dP - delta (1 - p))      16 warps, thread = (row, 32 columns)
//     dQ = dS' K, dK = dS'^T Q, dV = Pd^T dO  (128 x 64 x 128 each; they overwrite the S / dP columns of the same set);
//     the dropout factor 1 / (1 - p) is applied once in the epilogue (dS = dS' / (1 - p), dropout(P) = Pd / (1 - p))
// Pd and dS' go through shared memory once as bf16 [query][key] tiles: K-major A operand for dQ and — read transposed by
// the MN-major descriptor — A operand for dK / dV; K, Q, dO are consumed MN-major as B operands from the tiles TMA
// delivered. The compute warps keep item n+1's P / dS in registers while the tensor core works on item n.
namespace bwd {
constexpr int kStages = 2;                      // input ring: {Q, K, V, dO}
constexpr int kComputeWarps = 16;
constexpr int kComputeThreads = 32 * kComputeWarps;
constexpr int kProducerWarp = kComputeWarps, kMmaWarp = kComputeWarps + 1;
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr uint32_t kStageBytes = 4 * kMatBytes;
constexpr uint32_t kPOff = kStages * kStageBytes;           // Pd tile
constexpr uint32_t kDsOff = kPOff + 2 * kMatBytes;          // dS tile
constexpr uint32_t kAuxOff = kDsOff + 2 * kMatBytes;        // kvbits[2 parities][4 words]
constexpr uint32_t kBarOff = kAuxOff + 64;
constexpr uint32_t kSmemBytes = kBarOff + 256;
constexpr uint32_t kTmemCols = 512;
}  // namespace bwd

struct BwdArgs {
  const int2* tiles;
  const int32_t* counts;
  int count_idx;
  const int2* row_bounds;
  const uint8_t* key_valid;
  int heads, T, T_active;
  const float* lse;
  const float* delta;
  int delta_pitch;
  __nv_bfloat16* dqkv;
  float scale, rscale;
  uint32_t thr, seed;
  const uint32_t* salt;   // per-step dropout salt (ptx.cuh step_salt)
};

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <bool kDrop>
__global__ void __launch_bounds__(bwd::kThreads, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const BwdArgs a) {
  pdl_grid_sync();
  const uint32_t seed_eff = kDrop ? (a.seed ^ step_salt(a.salt)) : 0u;
  using namespace bwd;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint32_t* kvbits_s = reinterpret_cast<uint32_t*>(smem + kAuxOff);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* in_empty = in_full + kStages;
  uint64_t* sdp_full = in_empty + kStages;
  uint64_t* out_full = sdp_full + 2;
  uint64_t* acc_free = out_full + 2;
  uint64_t* pds_full = acc_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pds_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = a.counts[a.count_idx];
  const int total = n_tiles * a.heads;
  const int hd = a.heads * kD;
  int n_items = 0;
  for (int it = blockIdx.x; it < total; it += gridDim.x) ++n_items;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sdp_full[b], 1);
      mbar_init(&out_full[b], 1);
      mbar_init(&acc_free[b], kComputeThreads);
    }
    mbar_init(pds_full, kComputeThreads);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    int n = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x, ++n) {
      const int st = n % kStages;
      mbar_wait(&in_empty[st], ((uint32_t)(n / kStages) & 1u) ^ 1u);
      if (lane == 0) {
        const int tile = it / a.heads, h = it - tile * a.heads;
        const int row0 = a.tiles[tile].x;
        uint8_t* dst = smem + st * kStageBytes;
        mbar_arrive_expect_tx(&in_full[st], kStageBytes);
        tma_load_2d(&tmQKV, &in_full[st], dst, h * kD, row0);                            // Q
        tma_load_2d(&tmQKV, &in_full[st], dst + kMatBytes, hd + h * kD, row0);           // K
        tma_load_2d(&tmQKV, &in_full[st], dst + 2 * kMatBytes, 2 * hd + h * kD, row0);   // V
        tma_load_2d(&tmDO, &in_full[st], dst + 3 * kMatBytes, h * kD, row0);             // dO
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(kTile, kTile, 0, 0);    // S, dP : K-major x K-major
    constexpr uint32_t idesc_dq = umma_idesc_bf16(kTile, kD, 0, 1);      // dQ    : dS (K-major) x K (MN-major)
    constexpr uint32_t idesc_dkv = umma_idesc_bf16(kTile, kD, 1, 1);     // dK, dV: dS^T / Pd^T (MN-major) x Q / dO (MN-major)
    auto issue_sdp = [&](int n) {
      const int st = n % kStages, b = n & 1;
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem + st * kStageBytes), sk = sq + kMatBytes, sv = sq + 2 * kMatBytes, so = sq + 3 * kMatBytes;
        const uint32_t d = tmem_base + (uint32_t)b * 256u;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(d, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(d + 128u, umma_smem_desc(so + k * 32, 16, 1024), umma_smem_desc(sv + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&sdp_full[b]);
      }
      __syncwarp();
    };
    auto issue_grads = [&](int n) {
      const int st = n % kStages, b = n & 1;
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem + st * kStageBytes), sk = sq + kMatBytes, so = sq + 3 * kMatBytes;
        const uint32_t sp = smem_u32(smem + kPOff), sds = smem_u32(smem + kDsOff);
        const uint32_t d = tmem_base + (uint32_t)b * 256u;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)      // dQ += dS[:, 16k..] K[16k.., :]
          umma_bf16(d, umma_smem_desc(sds + (k >> 2) * kMatBytes + (k & 3) * 32, 16, 1024),
                    umma_smem_desc(sk + k * 2048, kMatBytes, 1024), idesc_dq, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)      // dK += dS[16k.., :]^T Q[16k.., :]
          umma_bf16(d + 64u, umma_smem_desc(sds + k * 2048, kMatBytes, 1024), umma_smem_desc(sq + k * 2048, kMatBytes, 1024),
                    idesc_dkv, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)      // dV += Pd[16k.., :]^T dO[16k.., :]
          umma_bf16(d + 128u, umma_smem_desc(sp + k * 2048, kMatBytes, 1024), umma_smem_desc(so + k * 2048, kMatBytes, 1024),
                    idesc_dkv, k > 0 ? 1u : 0u);
        umma_commit(&out_full[b]);
        umma_commit(&in_empty[st]);
      }
      __syncwarp();
    };
    // Fixed order grads(n), S/dP(n+1) with blocking waits (measured: putting S/dP(n+1) ahead of the gradient products —
    // compute warps draining TMEM before they publish Pd / dS, event-driven issuer — is 10 % slower, 59.6 vs 53.4 us).
    for (int n = 0; n < n_items; ++n) {
      if (n == 0) {
        mbar_wait(&in_full[0], 0);
        issue_sdp(0);
      }
      mbar_wait(pds_full, (uint32_t)n & 1u);
      issue_grads(n);
      if (n + 1 < n_items) {
        mbar_wait(&in_full[(n + 1) % kStages], (uint32_t)((n + 1) / kStages) & 1u);
        mbar_wait(&acc_free[(n + 1) & 1], ((uint32_t)((n + 1) >> 1) & 1u) ^ 1u);
        issue_sdp(n + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ compute warps: thread = (query row, 32 key columns)
    const int q = warp & 3, c = warp >> 2;
    const int r = q * 32 + lane;
    const float sl2 = a.scale * kLog2e;
    const float keep_p = 1.0f / a.rscale;     // 1 - p
    const int64_t ld = 3 * hd;
    uint32_t ppk[16], dspk[16];      // item n+1's Pd / dS' for this thread's 32 columns, packed bf16 pairs
    int t_cur = 0, h_cur = 0;
    bool ok_cur = false;             // (row, head, owned) of the item whose gradients sit in TMEM

    // per-item row metadata, fetched one item ahead of its use
    struct Meta {
      int2 tl, bd;
      float lse, delta;
    };
    auto fetch = [&](int it) {
      Meta m;
      m.tl = make_int2(0, 0);
      m.bd = make_int2(0, 0);
      m.lse = 0.f;
      m.delta = 0.f;
      if (it < total) {
        const int tile = it / a.heads, h = it - tile * a.heads;
        m.tl = a.tiles[tile];
        if (r < m.tl.y) {
          const int t = m.tl.x + r;
          m.bd = a.row_bounds[t];
          m.lse = a.lse[(int64_t)h * a.T + t];
          m.delta = a.delta[(int64_t)h * a.delta_pitch + t];
        }
      }
      return m;
    };
    auto compute_regs = [&](int n, int it, const Meta& md) {
      const int b = n & 1;
      const int tile = it / a.heads, h = it - tile * a.heads;
      const int2 tl = md.tl;
      const int t = tl.x + r;
      const bool row_ok = r < tl.y;
      const int lo = row_ok ? md.bd.x - tl.x : 0, hi = row_ok ? md.bd.y - tl.x : 0;
      const float lse2 = row_ok ? md.lse * kLog2e : INFINITY;
      const float dl = row_ok ? md.delta * keep_p : 0.f;
      const uint32_t* kvb = nullptr;
      if (a.key_valid != nullptr) {
        const bool kv = t < a.T && a.key_valid[t] != 0;
        const uint32_t word = __ballot_sync(0xffffffffu, kv);
        uint32_t* dst = kvbits_s + (n & 1) * 4;
        if (lane == 0) dst[q] = word;      // (the four column quarters of a row quarter write the same word)
        named_bar_sync(1, kComputeThreads);
        kvb = dst;
      }
      uint32_t m, keep;
      row_masks_word<kDrop>(lo, hi, kvb, seed_eff, attn_quad_row(h, a.T, t), a.thr, c, m, keep);
      const uint32_t kp = keep & m;
      const bool need = __any_sync(0xffffffffu, m != 0u);
      mbar_wait(&sdp_full[b], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      if (!need) {
#pragma unroll
        for (int j = 0; j < 16; ++j) ppk[j] = dspk[j] = 0u;
      } else {
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)b * 256u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {       // two 16-column halves keep the live registers below the 576-thread budget
          uint32_t sv[16], dv[16];
          tmem_ld_32x16(lane_base + c * 32 + hh * 16, sv);
          tmem_ld_32x16(lane_base + 128 + c * 32 + hh * 16, dv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float pd[2], ds[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int bit = hh * 16 + j + e;
              const float ex = ex2_approx(fmaf(__uint_as_float(sv[j + e]), sl2, -lse2));
              const float p = (m >> bit) & 1u ? ex : 0.f;
              const bool kb = (kp >> bit) & 1u;
              pd[e] = kb ? ex : 0.f;
              ds[e] = p * ((kb ? __uint_as_float(dv[j + e]) : 0.f) - dl);
            }
            ppk[hh * 8 + (j >> 1)] = pack_bf16x2(pd[0], pd[1]);
            dspk[hh * 8 + (j >> 1)] = pack_bf16x2(ds[0], ds[1]);
          }
        }
      }
      return make_int4(t, h, row_ok ? 1 : 0, 0);
    };
    auto store_regs = [&]() {
      uint8_t* p_s = smem + kPOff;
      uint8_t* ds_s = smem + kDsOff;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const uint32_t off = p_tile_off(r, c * 4 + s4);
        *reinterpret_cast<uint4*>(p_s + off) = make_uint4(ppk[4 * s4], ppk[4 * s4 + 1], ppk[4 * s4 + 2], ppk[4 * s4 + 3]);
        *reinterpret_cast<uint4*>(ds_s + off) = make_uint4(dspk[4 * s4], dspk[4 * s4 + 1], dspk[4 * s4 + 2], dspk[4 * s4 + 3]);
      }
    };
    auto epilogue = [&](int n) {       // gradients of item n: dQ | dK | dV, 16 of each matrix's 64 columns per thread
      const int b = n & 1;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)b * 256u;
#pragma unroll
      for (int mtx = 0; mtx < 3; ++mtx) {
        uint32_t v[16];
        tmem_ld_32x16(lane_base + mtx * 64 + c * 16, v);
        tmem_ld_wait();
        if (ok_cur) {
          const float sc = mtx < 2 ? a.scale * a.rscale : a.rscale;
          uint4* dst = reinterpret_cast<uint4*>(a.dqkv + (int64_t)t_cur * ld + mtx * hd + h_cur * kD + c * 16);
#pragma unroll
          for (int s4 = 0; s4 < 2; ++s4)
            dst[s4] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * s4]) * sc, __uint_as_float(v[8 * s4 + 1]) * sc),
                                 pack_bf16x2(__uint_as_float(v[8 * s4 + 2]) * sc, __uint_as_float(v[8 * s4 + 3]) * sc),
                                 pack_bf16x2(__uint_as_float(v[8 * s4 + 4]) * sc, __uint_as_float(v[8 * s4 + 5]) * sc),
                                 pack_bf16x2(__uint_as_float(v[8 * s4 + 6]) * sc, __uint_as_float(v[8 * s4 + 7]) * sc));
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free[b]);
    };

    int4 nxt = make_int4(0, 0, 0, 0);
    Meta md = fetch(blockIdx.x);
    if (n_items > 0) nxt = compute_regs(0, blockIdx.x, md);
    md = fetch(blockIdx.x + gridDim.x);
    int it = blockIdx.x;
    for (int n = 0; n < n_items; ++n, it += gridDim.x) {
      if (n > 0) {                    // the tensor core has finished reading item n-1's Pd / dS tiles (and its gradients are in TMEM)
        mbar_wait(&out_full[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1u);
        tc_fence_after();
      }
      store_regs();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      if (n > 0) epilogue(n - 1);
      t_cur = nxt.x;
      h_cur = nxt.y;
      ok_cur = nxt.z != 0;
      if (n + 1 < n_items) {
        const Meta cur = md;
        md = fetch(it + 2 * gridDim.x);
        nxt = compute_regs(n + 1, it + gridDim.x, cur);
      }
    }
    if (n_items > 0) {
      mbar_wait(&out_full[(n_items - 1) & 1], (uint32_t)((n_items - 1) >> 1) & 1u);
      tc_fence_after();
      epilogue(n_items - 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, bwd::kTmemCols);
  }
}

inline uint32_t drop_threshold(float p) {
  const double t = (double)p * 65536.0 + 0.5;   // 16-bit threshold (ptx.cuh dropout_keep)
  return p <= 0.f ? 0u : (t >= 65535.0 ? 65535u : (uint32_t)t);
}

}  // namespace

extern "C" int nbest_attn_plan(nbest_ctx* ctx, const int32_t* cu_seqlens, const int32_t* seq_of, int B, int T, int break_at,
                               int32_t* tiles, int32_t* counts, int32_t* row_bounds, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, cu_seqlens && seq_of && tiles && counts && row_bounds, "null pointer");
  NBEST_CHECK_ARG(ctx, B > 0 && T > 0 && break_at >= 0 && break_at <= B, "need B > 0, T > 0, 0 <= break_at <= B");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  nbest_launch(attn_plan_kernel, dim3(1), dim3(256), 0, s, cu_seqlens, B, break_at, reinterpret_cast<int2*>(tiles), counts);
  NBEST_CHECK_LAUNCH(ctx);
  nbest_launch(attn_row_bounds_kernel, dim3((T + 255) / 256), dim3(256), 0, s, cu_seqlens, seq_of, T, reinterpret_cast<int2*>(row_bounds));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_attn_tiles_fwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* tiles, const int32_t* counts,
                                    int count_idx, int max_tiles, const int32_t* row_bounds, const uint8_t* key_valid,
                                    int heads, int T, void* out_bf16, float* lse, float p_drop, uint32_t seed, void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, qkv_bf16 && tiles && counts && row_bounds && out_bf16 && lse, "null pointer");
  NBEST_CHECK_ARG(ctx, heads > 0 && T > 0 && max_tiles > 0 && (count_idx == 0 || count_idx == 1), "bad sizes");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  NBEST_CHECK_ARG(ctx, p_drop <= 0.f || (int64_t)heads * (int64_t)T < (1LL << 25), "dropout counter (h * T + t) * 128 would wrap 32 bits");
  CUtensorMap tm;
  int rc = nbest_make_tmap_bf16(ctx, &tm, qkv_bf16, (uint64_t)T, (uint64_t)(3 * heads * kD), (uint64_t)(3 * heads * kD), kTile);
  if (rc != NBEST_OK) return rc;
  static bool attr_dev[64] = {};
  if (!attr_dev[ctx->device & 63]) {
    NBEST_CHECK_CUDA(ctx, cudaFuncSetAttribute(attn_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::kSmemBytes));
    NBEST_CHECK_CUDA(ctx, cudaFuncSetAttribute(attn_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::kSmemBytes));
    attr_dev[ctx->device & 63] = true;
  }
  FwdArgs a;
  a.tiles = reinterpret_cast<const int2*>(tiles);
  a.counts = counts;
  a.count_idx = count_idx;
  a.row_bounds = reinterpret_cast<const int2*>(row_bounds);
  a.key_valid = key_valid;
  a.heads = heads;
  a.T = T;
  a.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  a.lse = lse;
  a.scale = 0.125f;
  a.rscale = 1.0f / (1.0f - p_drop);
  a.thr = drop_threshold(p_drop);
  a.seed = seed;
  a.salt = nbest_salt(ctx);
  const int64_t items = (int64_t)max_tiles * heads;
  const int grid = (int)(items < ctx->num_sms ? items : ctx->num_sms);
  if (a.thr != 0u)
    nbest_launch(attn_tc_fwd_kernel<true>, dim3(grid), dim3(fwd::kThreads), fwd::kSmemBytes, reinterpret_cast<cudaStream_t>(stream), tm, a);
  else
    nbest_launch(attn_tc_fwd_kernel<false>, dim3(grid), dim3(fwd::kThreads), fwd::kSmemBytes, reinterpret_cast<cudaStream_t>(stream), tm, a);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

extern "C" int nbest_attn_tiles_bwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* tiles, const int32_t* counts,
                                    int count_idx, int max_tiles, const int32_t* row_bounds, const uint8_t* key_valid,
                                    int heads, int T, int T_active, const void* dout_bf16, const float* lse,
                                    const float* delta, int delta_pitch, void* dqkv_bf16, float p_drop, uint32_t seed,
                                    void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, qkv_bf16 && tiles && counts && row_bounds && dout_bf16 && lse && delta && dqkv_bf16, "null pointer");
  NBEST_CHECK_ARG(ctx, heads > 0 && T > 0 && max_tiles > 0 && (count_idx == 0 || count_idx == 1), "bad sizes");
  NBEST_CHECK_ARG(ctx, T_active > 0 && T_active <= T && delta_pitch >= T_active, "need 0 < T_active <= T, delta_pitch >= T_active");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  NBEST_CHECK_ARG(ctx, p_drop <= 0.f || (int64_t)heads * (int64_t)T < (1LL << 25), "dropout counter (h * T + t) * 128 would wrap 32 bits");
  CUtensorMap tmq, tmo;
  int rc = nbest_make_tmap_bf16(ctx, &tmq, qkv_bf16, (uint64_t)T, (uint64_t)(3 * heads * kD), (uint64_t)(3 * heads * kD), kTile);
  if (rc != NBEST_OK) return rc;
  rc = nbest_make_tmap_bf16(ctx, &tmo, dout_bf16, (uint64_t)T_active, (uint64_t)(heads * kD), (uint64_t)(heads * kD), kTile);
  if (rc != NBEST_OK) return rc;
  static bool attr_dev[64] = {};
  if (!attr_dev[ctx->device & 63]) {
    NBEST_CHECK_CUDA(ctx, cudaFuncSetAttribute(attn_tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmemBytes));
    NBEST_CHECK_CUDA(ctx, cudaFuncSetAttribute(attn_tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmemBytes));
    attr_dev[ctx->device & 63] = true;
  }
  BwdArgs a;
  a.tiles = reinterpret_cast<const int2*>(tiles);
  a.counts = counts;
  a.count_idx = count_idx;
  a.row_bounds = reinterpret_cast<const int2*>(row_bounds);
  a.key_valid = key_valid;
  a.heads = heads;
  a.T = T;
  a.T_active = T_active;
  a.lse = lse;
  a.delta = delta;
  a.delta_pitch = delta_pitch;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_bf16);
  a.scale = 0.125f;
  a.rscale = 1.0f / (1.0f - p_drop);
  a.thr = drop_threshold(p_drop);
  a.seed = seed;
  a.salt = nbest_salt(ctx);
  const int64_t items = (int64_t)max_tiles * heads;
  const int grid = (int)(items < ctx->num_sms ? items : ctx->num_sms);
  if (a.thr != 0u)
    nbest_launch(attn_tc_bwd_kernel<true>, dim3(grid), dim3(bwd::kThreads), bwd::kSmemBytes, reinterpret_cast<cudaStream_t>(stream), tmq, tmo, a);
  else
    nbest_launch(attn_tc_bwd_kernel<false>, dim3(grid), dim3(bwd::kThreads), bwd::kSmemBytes, reinterpret_cast<cudaStream_t>(stream), tmq, tmo, a);
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}
