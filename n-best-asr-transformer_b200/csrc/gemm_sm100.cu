// tcgen05 / TMEM bf16 GEMM for sm_100a, fed by TMA, persistent and warp-specialised.
//
// Replaces the nn.Linear calls inside the HF encoder that models/model.py:43-58 (reference) drives:
// BertSelfAttention q/k/v (transformers modeling_bert.py:179-181), BertSelfOutput.dense (:294-298),
// BertIntermediate.dense + GELU (:339-342), BertOutput.dense (:352-356), and their autograd products
// (dgrad / wgrad of n_best_asr_bert.py:264 `total_loss.backward()`).
//
// CG = 2 (default): CTA PAIRS. The two CTAs of a cluster (one per SM of a TPC) compute a 256 x BN tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), the leader CTA
// (cluster rank 0) issues the MMAs for the pair, and each CTA's TMEM receives its 128 accumulator rows. Per MMA that is
// 1.5x less shared-memory operand traffic and 1.5x less L2->SM traffic than two independent 128 x BN tiles — the two
// limits (128 B/clk/SM smem, ~6.3 KB/clk chip-wide L2) that held the single-CTA kernel at ~67 % tensor-pipe activity.
// CG = 1 keeps the single-CTA variant (NBEST_GEMM_CTA_GROUP=1) for A/B measurements.
//
// One CTA per SM (320 threads): warps 0..7 = epilogue (two warps per TMEM lane quarter, each owning 64-column units
// that leave through TMA stores), warp 8 = TMA producer, warp 9 = TMEM owner + single-thread tcgen05.mma issuer (the issuers have
// the highest warp ids so the scheduler never parks them behind busy epilogue warps). Tiles are 128 x BN (BN = 256 or 128) x 64, the smem ring has
// 4 (BN=256) or 6 (BN=128) stages of 128-byte-swizzled operand tiles, and the fp32 accumulator is double buffered
// in TMEM (2*BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand majorness is a template flag: K-major operands are read as {64 x rows} TMA boxes (one per stage), MN-major
// operands (dgrad weights, both wgrad operands) as 64-column x 64-k-row boxes laid out as the canonical MN-major
// SWIZZLE_128B UMMA layout (LBO = 8 KiB between 64-element chunks, SBO = 1 KiB between 8-row groups).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

using namespace nbest;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarp0 = 0;   // epilogue warps come FIRST: the warp scheduler favours higher warp ids, and the two
                               // single-thread issuers (TMA, MMA) must never starve behind busy epilogue warps
constexpr int kEpiWarps = 8;   // two per TMEM lane quarter; each owns whole 64-column units of the tile
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = 32 * (kEpiWarps + 2);
constexpr uint32_t kChunkBytes = 64 * BK * 2;  // one 64-row (or 64-col) x 64 bf16 box = 8 KiB

struct GemmArgs {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_splits, kb_per_split, num_kb;
  int debug;    // NBEST_GEMM_DEBUG: 1 = epilogue does nothing but release the accumulator, 2 = no global stores
  int stages;   // smem ring depth in use (<= Cfg::kStages; NBEST_GEMM_STAGES experiments)
  void* C;
  int64_t ldc;
  const float* bias;
  const __nv_bfloat16* aux;
  int64_t ldaux;
  __nv_bfloat16* out2;
  uint32_t drop_thresh;
  float drop_scale;
  uint32_t seed;
  const uint32_t* salt;   // per-step dropout salt (ptx.cuh step_salt)
  uint32_t* sched;   // {next work item, finished units} of this launch (context-owned, self-resetting), or NULL = static
};

constexpr int kSchedSlots = 4;   // depth of the work-item ring between the fetching producer and the other warp roles

template <int BN, int CG, bool kTwoOutputs = false>
struct Cfg {
  static constexpr uint32_t kABytes = BM * BK * 2;               // per CTA: its own 128 rows of A
  static constexpr uint32_t kBBytes = (BN / CG) * BK * 2;        // per CTA: BN / CG rows of B
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  // per epilogue warp: one 32-row x 128-byte store unit — two for the GELU epilogue, which produces gelu(u) and gelu'(u)
  // from ONE pass over the accumulator and stores both. (Two alternating units for single-output epilogues were measured:
  // no gain, the epilogues are not waiting for the store engine.)
  static constexpr uint32_t kStagingPerWarp = kTwoOutputs ? 8192 : 4096;
  static constexpr uint32_t kStagingBytes = kEpiWarps * kStagingPerWarp;
  static constexpr uint32_t kBiasBytes = 2 * BN * 4;   // the tile's bias slice, double buffered by tile parity
  static constexpr int kStages = (232448 - 256 - kBiasBytes - kStagingBytes) / kStageBytes;   // fill the 227 KiB window
  static constexpr uint32_t kBarOffset = kStages * kStageBytes + kStagingBytes + kBiasBytes;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256;   // + barriers; the dynamic window itself is 1024-byte aligned
  static constexpr uint32_t kTmemCols = 2 * BN;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int CG, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const GemmArgs g) {
  using C_ = Cfg<BN, CG, EPI == NBEST_EPI_BIAS_GELU>;
  constexpr int kStages = C_::kStages;
  // SWIZZLE_128B operand tiles need 1024-byte alignment; the window starts at the same offset in both CTAs of a pair
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* staging = smem + kStages * C_::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C_::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sched_full = tempty_bar + 2;
  uint64_t* sched_empty = sched_full + kSchedSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_empty + kSchedSlots);
  volatile int32_t* sched_item = reinterpret_cast<volatile int32_t*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  const int unit = blockIdx.x / CG;                           // persistent work unit: a CTA (CG = 1) or a CTA pair
  const int num_units = gridDim.x / CG;
  const int num_work = g.num_m_tiles * g.num_n_tiles * g.num_splits;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI != NBEST_EPI_ACCUM_F32) tma_prefetch_desc(&tmC);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);    // pair mode: only the leader's is used; it collects the bytes of BOTH CTAs' loads
      mbar_init(&empty_bar[s], 1);   // arrived by tcgen05.commit (multicast to both CTAs in pair mode)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps * CG);   // pair mode: the leader's collects both CTAs' epilogue warps
    }
    for (int a = 0; a < kSchedSlots; ++a) {
      mbar_init(&sched_full[a], 1);
      // a slot is free again once every reader has copied the item: leader = MMA warp + epilogue warps, peer = producer +
      // epilogue warps; all of them arrive on the LEADER's barrier (the leader's producer is the only writer)
      mbar_init(&sched_empty[a], (kEpiWarps + 1) * CG);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, C_::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, C_::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncwarp();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();   // everything above (barriers, TMEM, descriptor prefetch) ran under the previous kernel's tail

  // ---- work distribution. Static: unit u takes items u, u + num_units, ... Dynamic (g.sched != NULL): the leader's
  // producer thread draws items from a global counter and publishes them through a small shared-memory ring to the other
  // warp roles of the CTA (and of the peer CTA of a pair). A CTA pair that starts late — because a collective holds its
  // SMs — then simply takes fewer items instead of delaying a statically assigned share of the tiles.
  const bool dyn = g.sched != nullptr;
  auto next_item = [&](int i) -> int {     // consumer side: i = running item count of this warp role
    if (!dyn) {
      const int w = unit + i * num_units;
      return w < num_work ? w : -1;
    }
    const int slot = i % kSchedSlots;
    if constexpr (CG == 2) mbar_wait_acquire_cluster(&sched_full[slot], (uint32_t)(i / kSchedSlots) & 1u);
    else mbar_wait(&sched_full[slot], (uint32_t)(i / kSchedSlots) & 1u);
    const int w = sched_item[slot];
    __syncwarp();
    if (lane == 0) {
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&sched_empty[slot]), 0));
      else mbar_arrive(&sched_empty[slot]);
    }
    return w;
  };

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t stage = 0, phase = 0;
    int fetched = 0;
    if (dyn && rank == 0 && lane == 0) fetched = (int)atomicAdd(g.sched, 1u);
    for (int i = 0;; ++i) {
      int w;
      if (dyn && rank == 0) {
        // leader producer: publish item i (or the end marker) to both CTAs, then draw item i+1 — the atomic's round
        // trip hides under this item's TMA issue
        const int slot = i % kSchedSlots;
        mbar_wait(&sched_empty[slot], ((uint32_t)(i / kSchedSlots) & 1u) ^ 1u);
        w = __shfl_sync(0xffffffffu, fetched, 0);
        if (w >= num_work) w = -1;
        if (lane == 0) {
          sched_item[slot] = w;
          mbar_arrive(&sched_full[slot]);
          if constexpr (CG == 2) {
            st_shared_cluster_u32(mapa_shared(smem_u32(const_cast<int32_t*>(&sched_item[slot])), 1), (uint32_t)w);
            mbar_arrive_release_cluster(mapa_shared(smem_u32(&sched_full[slot]), 1));
          }
          if (w >= 0) fetched = (int)atomicAdd(g.sched, 1u);
        }
      } else {
        w = next_item(i);
      }
      if (w < 0) break;
      const int num_tiles = g.num_m_tiles * g.num_n_tiles;   // tile fastest: concurrent CTAs share one T-range (L2 reuse)
      const int split = w / num_tiles;
      const int tile = w % num_tiles;
      const int m0 = (tile / g.num_n_tiles) * (BM * CG) + (int)rank * BM;            // this CTA's 128 rows of A / D
      const int nb0 = (tile % g.num_n_tiles) * BN + (int)rank * (BN / CG);           // this CTA's BN / CG rows of B
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(kb0 + g.kb_per_split, g.num_kb);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * C_::kStageBytes;
          uint8_t* sb = sa + C_::kABytes;
          // pair mode: both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], C_::kStageBytes * CG);
          const uint32_t bar = (CG == 2) ? mapa_shared(smem_u32(&full_bar[stage]), 0) : 0u;
          auto load = [&](const CUtensorMap* tm, uint8_t* dst, int c0, int c1) {
            if constexpr (CG == 2) tma_load_2d_cg2(tm, bar, dst, c0, c1);
            else tma_load_2d(tm, &full_bar[stage], dst, c0, c1);
          };
          if constexpr (!A_MN) {
            load(&tmA, sa, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(&tmA, sa + c * kChunkBytes, m0 + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            load(&tmB, sb, kb * BK, nb0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / CG / 64; ++c) load(&tmB, sb + c * kChunkBytes, nb0 + c * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (pair mode: the leader CTA only)
    constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    constexpr uint32_t a_lbo = A_MN ? kChunkBytes : 16, b_lbo = B_MN ? kChunkBytes : 16;
    constexpr uint32_t a_kstep = A_MN ? 2048 : 32, b_kstep = B_MN ? 2048 : 32;  // bytes per UMMA_K = 16
    uint32_t stage = 0, phase = 0, it = 0;
    if (rank == 0) {
    for (;; ++it) {
      const int w = next_item((int)it);
      if (w < 0) break;
      const int split = w / (g.num_m_tiles * g.num_n_tiles);
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(kb0 + g.kb_per_split, g.num_kb);
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * C_::kStageBytes);
          const uint32_t sb = sa + C_::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = umma_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = umma_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            if constexpr (CG == 2) umma_bf16_cg2(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) {
            umma_commit_cg2(&empty_bar[stage], 3);                  // frees the slot in BOTH CTAs once these MMAs retire
            if (kb == kb1 - 1) umma_commit_cg2(&tfull_bar[acc], 3);   // accumulator complete -> both CTAs' epilogues
          } else {
            umma_commit(&empty_bar[stage]);                 // frees the smem slot once these MMAs retire
            if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    // 8 warps: warp w reads TMEM lane quarter q = w % 4 (tcgen05.ld restriction) and owns the 64-column units
    // u = w/4, w/4 + 2, ... of the tile. A unit (32 rows x 64 bf16 = 128-byte rows) is assembled in a 4 KiB smem buffer
    // in the TMA SWIZZLE_128B layout and written with ONE cp.async.bulk.tensor store: full 128-byte lines, the M edge
    // is clipped by the tensor map. (Per-lane 16-byte st.global of 64-byte row segments cost 17 % of the GEMM.)
    const int q = warp & 3;
    const int u_first = warp >> 2;
    uint8_t* st = staging + warp * C_::kStagingPerWarp;
    float* bias_tile = reinterpret_cast<float*>(staging + C_::kStagingBytes);
    constexpr bool kHasBias = (EPI == NBEST_EPI_BIAS || EPI == NBEST_EPI_BIAS_GELU || EPI == NBEST_EPI_BIAS_DROP_RES);
    auto unit_off = [](int row, int chunk16) { return row * 128 + ((chunk16 ^ (row & 7)) << 4); };
    uint32_t drop_seed = 0;
    if constexpr (EPI == NBEST_EPI_BIAS_DROP_RES) drop_seed = g.seed ^ step_salt(g.salt);
    uint32_t it = 0;
    const uint32_t tempty_remote[2] = {(CG == 2) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u,
                                       (CG == 2) ? mapa_shared(smem_u32(&tempty_bar[1]), 0) : 0u};
    for (;; ++it) {
      const int w = next_item((int)it);
      if (w < 0) break;
      const int tile = w % (g.num_m_tiles * g.num_n_tiles);
      const int m0 = (tile / g.num_n_tiles) * (BM * CG) + (int)rank * BM;   // this CTA's 128 accumulator rows
      const int n0 = (tile % g.num_n_tiles) * BN;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int row = m0 + q * 32 + lane;  // the accumulator row this thread owns
      constexpr bool kHasAux = (EPI == NBEST_EPI_BIAS_DROP_RES || EPI == NBEST_EPI_DGELU || EPI == NBEST_EPI_ADD ||
                                EPI == NBEST_EPI_DELTA);
      const int crow = lane >> 2, cseg = lane & 3;   // coalesced aux loads: 8 rows x 64 B per warp instruction
      uint4 aux_next[4];
      auto load_aux = [&](int c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gr = m0 + q * 32 + i * 8 + crow;
          aux_next[i] = make_uint4(0, 0, 0, 0);
          if (gr < g.M) aux_next[i] = __ldg(reinterpret_cast<const uint4*>(g.aux + (int64_t)gr * g.ldaux + n0 + c * 32 + cseg * 8));
        }
      };
      if constexpr (kHasAux) load_aux(2 * u_first);
      // this warp's slice of the tile's bias (its own <= two 64-column units; the four warps of a unit write identical
      // values): the global load is issued before the accumulator wait so that its latency hides behind the MMAs, the
      // epilogue then reads it with short-latency broadcast LDS instead of stalling every chunk on global loads
      float* bias_s = bias_tile + acc * BN;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ub = u_first + 2 * (lane >> 4);
      if constexpr (kHasBias) {
        if (ub < BN / 64) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + ub * 64 + (lane & 15) * 4));
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      if constexpr (kHasBias) {
        // safe to overwrite buffer `acc` now: this tile's MMAs only started after every epilogue warp released the
        // accumulator (and with it the bias buffer) of the tile two iterations ago
        if (ub < BN / 64) *reinterpret_cast<float4*>(bias_s + ub * 64 + (lane & 15) * 4) = b4;
        __syncwarp();
      }
      tc_fence_after();
      if constexpr (EPI == NBEST_EPI_ACCUM_F32) {
#pragma unroll 1
        for (int c = u_first; c < BN / 32; c += 2) {
          if (g.debug == 1) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + acc * BN + c * 32, r);
          tmem_ld_wait();
          float* crow_ptr = reinterpret_cast<float*>(g.C) + (int64_t)row * g.ldc + n0 + c * 32;
          if (row < g.M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(crow_ptr + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                         __uint_as_float(r[j + 3]));
          }
        }
      } else {
#pragma unroll 1
        for (int u = u_first; u < BN / 64; u += 2) {
          if (g.debug == 1) break;
          {
            // GELU with out2: gelu(u) -> unit 0 of the staging buffer, gelu'(u) -> unit 1, both from this one pass
            const bool both = (EPI == NBEST_EPI_BIAS_GELU) && g.out2 != nullptr;
            if (lane == 0) bulk_wait_read();   // the previous store from this buffer has finished reading it
            __syncwarp();
            float dsum = 0.f;                  // EPI_DELTA: dot(acc, aux) over this 64-column unit (= one attention head)
            float rsum = 0.f, rsq = 0.f;       // EPI_BIAS_DROP_RES: row sum / sum of squares of this unit (LayerNorm statistics)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = 2 * u + cc;
              const int nc = n0 + c * 32;
              uint32_t r[32];
              tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + acc * BN + c * 32, r);
              uint32_t auxrow[16];
              if constexpr (kHasAux) {
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(st + unit_off(i * 8 + crow, cc * 4 + cseg)) = aux_next[i];
                const int cn = (cc == 0) ? c + 1 : c + 3;   // the next chunk this warp will process
                if (cn < BN / 32) load_aux(cn);
                __syncwarp();
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                  const uint4 t4 = *reinterpret_cast<const uint4*>(st + unit_off(lane, cc * 4 + s4));
                  auxrow[s4 * 4 + 0] = t4.x;
                  auxrow[s4 * 4 + 1] = t4.y;
                  auxrow[s4 * 4 + 2] = t4.z;
                  auxrow[s4 * 4 + 3] = t4.w;
                }
                __syncwarp();
              }
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
              if constexpr (kHasBias) {
                const float* bsrc = bias_s + c * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {     // packed fp32x2 adds (two elements per FADD2)
                  const float4 b = *reinterpret_cast<const float4*>(bsrc + j);
                  f2_unpack(f2_add(f2_pack(v[j], v[j + 1]), f2_pack(b.x, b.y)), v[j], v[j + 1]);
                  f2_unpack(f2_add(f2_pack(v[j + 2], v[j + 3]), f2_pack(b.z, b.w)), v[j + 2], v[j + 3]);
                }
              }
              if constexpr (EPI == NBEST_EPI_BIAS_GELU) {
                // packed fp32x2 evaluation: two elements per FFMA2 / FMUL2 (ptx.cuh gelu_fwd_grad2)
                if (both) {
#pragma unroll
                  for (int s4 = 0; s4 < 4; ++s4) {
                    float d[8];
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                      f32x2 gg, dd;
                      gelu_fwd_grad2(f2_pack(v[8 * s4 + j], v[8 * s4 + j + 1]), gg, dd);
                      f2_unpack(gg, v[8 * s4 + j], v[8 * s4 + j + 1]);
                      f2_unpack(dd, d[j], d[j + 1]);
                    }
                    *reinterpret_cast<uint4*>(st + 4096 + unit_off(lane, cc * 4 + s4)) =
                        make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; j += 2) f2_unpack(gelu_fwd2(f2_pack(v[j], v[j + 1])), v[j], v[j + 1]);
                }
              } else if constexpr (EPI == NBEST_EPI_BIAS_DROP_RES) {
                if (g.drop_thresh != 0) {
                  const uint32_t base = (uint32_t)row * (uint32_t)g.N + (uint32_t)nc;
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    bool k0, k1, k2, k3;
                    dropout_keep4(drop_seed, base + j, g.drop_thresh, k0, k1, k2, k3);   // base % 4 == 0 (N, nc multiples of 32)
                    f2_unpack(f2_mul(f2_pack(v[j], v[j + 1]), f2_pack(k0 ? g.drop_scale : 0.f, k1 ? g.drop_scale : 0.f)), v[j], v[j + 1]);
                    f2_unpack(f2_mul(f2_pack(v[j + 2], v[j + 3]), f2_pack(k2 ? g.drop_scale : 0.f, k3 ? g.drop_scale : 0.f)), v[j + 2],
                              v[j + 3]);
                  }
                }
                f32x2 rs2 = 0ull, rq2 = 0ull;      // (0.f, 0.f)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const f32x2 x2 = f2_add(f2_pack(v[2 * j], v[2 * j + 1]), f2_pack(bf16lo(auxrow[j]), bf16hi(auxrow[j])));
                  f2_unpack(x2, v[2 * j], v[2 * j + 1]);
                  rs2 = f2_add(rs2, x2);           // partial LayerNorm statistics of the row the next kernel normalises
                  rq2 = f2_fma(x2, x2, rq2);
                }
                if (g.out2 != nullptr) {
                  float a0, a1, q0, q1;
                  f2_unpack(rs2, a0, a1);
                  f2_unpack(rq2, q0, q1);
                  rsum += a0 + a1;
                  rsq += q0 + q1;
                }
              } else if constexpr (EPI == NBEST_EPI_DGELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j)     // aux = gelu'(u), saved by the forward's NBEST_EPI_BIAS_GELU epilogue
                  f2_unpack(f2_mul(f2_pack(v[2 * j], v[2 * j + 1]), f2_pack(bf16lo(auxrow[j]), bf16hi(auxrow[j]))), v[2 * j], v[2 * j + 1]);
              } else if constexpr (EPI == NBEST_EPI_ADD) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  f2_unpack(f2_add(f2_pack(v[2 * j], v[2 * j + 1]), f2_pack(bf16lo(auxrow[j]), bf16hi(auxrow[j]))), v[2 * j], v[2 * j + 1]);
              } else if constexpr (EPI == NBEST_EPI_DELTA) {
                f32x2 ds2 = 0ull;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  ds2 = f2_fma(f2_pack(v[2 * j], v[2 * j + 1]), f2_pack(bf16lo(auxrow[j]), bf16hi(auxrow[j])), ds2);
                float e0, e1;
                f2_unpack(ds2, e0, e1);
                dsum += e0 + e1;
              }
#pragma unroll
              for (int s4 = 0; s4 < 4; ++s4)
                *reinterpret_cast<uint4*>(st + unit_off(lane, cc * 4 + s4)) =
                    make_uint4(pack_bf16x2(v[8 * s4], v[8 * s4 + 1]), pack_bf16x2(v[8 * s4 + 2], v[8 * s4 + 3]),
                               pack_bf16x2(v[8 * s4 + 4], v[8 * s4 + 5]), pack_bf16x2(v[8 * s4 + 6], v[8 * s4 + 7]));
            }
            if constexpr (EPI == NBEST_EPI_DELTA) {
              if (row < g.M) reinterpret_cast<float*>(g.out2)[(int64_t)((n0 >> 6) + u) * g.M + row] = dsum;
            }
            if constexpr (EPI == NBEST_EPI_BIAS_DROP_RES) {
              if (g.out2 != nullptr && row < g.M)
                reinterpret_cast<float2*>(g.out2)[(int64_t)row * (g.N >> 6) + (n0 >> 6) + u] = make_float2(rsum, rsq);
            }
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0 && g.debug != 2) {
              tma_store_2d(&tmC, st, n0 + u * 64, m0 + q * 32);
              if (both) tma_store_2d(&tmC2, st + 4096, n0 + u * 64, m0 + q * 32);
              bulk_commit();
            }
            if constexpr (EPI == NBEST_EPI_DGELU) {
              // optional fused bias gradient: out2[n] += sum_m C[m, n]. Lane j sums column pair (2j, 2j+1) of the staged
              // 32 x 64 unit (the bf16 values the store writes; rows beyond M are exact zeros) and adds it to the fp32
              // accumulator: one pass over du less than a separate column-sum kernel.
              if (g.out2 != nullptr) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  const uint32_t w2 = *reinterpret_cast<const uint32_t*>(st + unit_off(r, lane >> 2) + (lane & 3) * 4);
                  s0 += bf16lo(w2);
                  s1 += bf16hi(w2);
                }
                float* dst = reinterpret_cast<float*>(g.out2) + n0 + u * 64 + 2 * lane;
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(s0), "f"(s1) : "memory");
              }
            }
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(tempty_remote[acc]);   // the leader's MMA warp waits for both CTAs
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
    if (lane == 0) bulk_wait_all();   // every bulk store of this warp is globally complete before the CTA exits
  }

  if (dyn && warp == kProducerWarp && rank == 0 && lane == 0) {
    // the last unit to finish re-arms the counters for the launch that will use this scheduler slot next
    __threadfence();
    if (atomicAdd(g.sched + 1, 1u) == (uint32_t)num_units - 1u) {
      g.sched[0] = 0u;
      g.sched[1] = 0u;
      __threadfence();
    }
  }
  tc_fence_before();
  __syncwarp();
  // pair mode: neither CTA may exit (or free TMEM) while the peer can still read its smem / signal its barriers
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, C_::kTmemCols);
    else tmem_dealloc(tmem_base, C_::kTmemCols);
  }
}

template <int BN, int CG, bool A_MN, bool B_MN, int EPI>
int launch(nbest_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
           const GemmArgs& g, cudaStream_t stream) {
  auto kfn = gemm_kernel<BN, CG, A_MN, B_MN, EPI>;
  using C_ = Cfg<BN, CG, EPI == NBEST_EPI_BIAS_GELU>;
  static int max_units_dev[64] = {};  // per instantiation and device: CTAs (CG = 1) or co-resident CTA pairs (CG = 2)
  int& max_units = max_units_dev[ctx->device & 63];
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C_::kSmemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  if (CG == 2) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
    attr[cfg.numAttrs].val.clusterDim.x = 2;
    attr[cfg.numAttrs].val.clusterDim.y = 1;
    attr[cfg.numAttrs].val.clusterDim.z = 1;
    ++cfg.numAttrs;
  }
  const int n_cluster_attrs = cfg.numAttrs;
  if (max_units == 0) {
    NBEST_CHECK_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::kSmemBytes));
    if (CG == 2) {
      cfg.gridDim = dim3(ctx->num_sms / 2 * 2);
      int n = 0;
      NBEST_CHECK_CUDA(ctx, cudaOccupancyMaxActiveClusters(&n, kfn, &cfg));
      if (n < 1) {
        nbest_set_error(ctx, "nbest_gemm_bf16: no CTA pair of %u bytes of shared memory fits on this device", C_::kSmemBytes);
        return NBEST_ECUDA;
      }
      max_units = n < ctx->num_sms / 2 ? n : ctx->num_sms / 2;
    } else {
      max_units = ctx->num_sms;
    }
  }
  const int num_work = g.num_m_tiles * g.num_n_tiles * g.num_splits;
  int avail = max_units - ctx->reserve_sms / CG;
  if (avail < 1) avail = 1;
  const int units = num_work < avail ? num_work : avail;
  cfg.gridDim = dim3(units * CG);
  cfg.numAttrs = n_cluster_attrs;
  if (g_nbest_pdl) {
    attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
    ++cfg.numAttrs;
  }
  NBEST_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, kfn, tmA, tmB, tmC, tmC2, g));
  NBEST_CHECK_LAUNCH(ctx);
  return NBEST_OK;
}

template <int BN, int CG>
int dispatch(nbest_ctx* ctx, int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
             const CUtensorMap& tmC, const CUtensorMap& tmC2, const GemmArgs& g, cudaStream_t s) {
  if (!a_mn && !b_mn) {
    switch (epi) {
      case NBEST_EPI_NONE: return launch<BN, CG, false, false, NBEST_EPI_NONE>(ctx, tmA, tmB, tmC, tmC2, g, s);
      case NBEST_EPI_BIAS: return launch<BN, CG, false, false, NBEST_EPI_BIAS>(ctx, tmA, tmB, tmC, tmC2, g, s);
      case NBEST_EPI_BIAS_GELU: return launch<BN, CG, false, false, NBEST_EPI_BIAS_GELU>(ctx, tmA, tmB, tmC, tmC2, g, s);
      case NBEST_EPI_BIAS_DROP_RES: return launch<BN, CG, false, false, NBEST_EPI_BIAS_DROP_RES>(ctx, tmA, tmB, tmC, tmC2, g, s);
      default: break;
    }
  } else if (!a_mn && b_mn) {
    switch (epi) {
      case NBEST_EPI_NONE: return launch<BN, CG, false, true, NBEST_EPI_NONE>(ctx, tmA, tmB, tmC, tmC2, g, s);
      case NBEST_EPI_DGELU: return launch<BN, CG, false, true, NBEST_EPI_DGELU>(ctx, tmA, tmB, tmC, tmC2, g, s);
      case NBEST_EPI_ADD: return launch<BN, CG, false, true, NBEST_EPI_ADD>(ctx, tmA, tmB, tmC, tmC2, g, s);
      case NBEST_EPI_DELTA: return launch<BN, CG, false, true, NBEST_EPI_DELTA>(ctx, tmA, tmB, tmC, tmC2, g, s);
      default: break;
    }
  } else if (a_mn && b_mn) {
    if (epi == NBEST_EPI_ACCUM_F32) return launch<BN, CG, true, true, NBEST_EPI_ACCUM_F32>(ctx, tmA, tmB, tmC, tmC2, g, s);
  }
  nbest_set_error(ctx, "nbest_gemm_bf16: unsupported (a_mn_major=%d, b_mn_major=%d, epilogue=%d) combination", a_mn, b_mn,
                  epi);
  return NBEST_EINVAL;
}

}  // namespace

extern "C" int nbest_gemm_bf16(nbest_ctx* ctx, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                               int b_mn_major, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* bias,
                               const void* aux_bf16, int64_t ldaux, void* out2_bf16, float p_drop, uint32_t seed,
                               void* stream) {
  if (!ctx) return NBEST_EINVAL;
  NBEST_CHECK_ARG(ctx, A && B && C, "null operand");
  NBEST_CHECK_ARG(ctx, M > 0 && N > 0 && K > 0, "empty problem");
  NBEST_CHECK_ARG(ctx, N % 128 == 0, "N must be a multiple of 128");
  NBEST_CHECK_ARG(ctx, (a_mn_major && b_mn_major) || K % 64 == 0, "K must be a multiple of 64 for K-major operands");
  NBEST_CHECK_ARG(ctx, !a_mn_major || M % 128 == 0, "M must be a multiple of 128 when A is MN-major");
  NBEST_CHECK_ARG(ctx, ldc % 8 == 0, "ldc must be a multiple of 8");
  const bool needs_bias =
      epilogue == NBEST_EPI_BIAS || epilogue == NBEST_EPI_BIAS_GELU || epilogue == NBEST_EPI_BIAS_DROP_RES;
  const bool needs_aux = epilogue == NBEST_EPI_BIAS_DROP_RES || epilogue == NBEST_EPI_DGELU || epilogue == NBEST_EPI_ADD ||
                         epilogue == NBEST_EPI_DELTA;
  NBEST_CHECK_ARG(ctx, epilogue != NBEST_EPI_DELTA || out2_bf16, "NBEST_EPI_DELTA needs out2 (fp32 [N/64][M])");
  NBEST_CHECK_ARG(ctx, !needs_bias || bias, "epilogue needs bias");
  NBEST_CHECK_ARG(ctx, !needs_aux || (aux_bf16 && ldaux % 8 == 0), "epilogue needs aux with ldaux % 8 == 0");
  NBEST_CHECK_ARG(ctx, p_drop >= 0.f && p_drop < 1.f, "p_drop out of range");
  // the dropout mask of element (row, col) is keyed on the 32-bit counter row * N + col (ptx.cuh dropout_keep4)
  NBEST_CHECK_ARG(ctx, !(epilogue == NBEST_EPI_BIAS_DROP_RES && p_drop > 0.f) || (int64_t)M * (int64_t)N < (1LL << 32),
                  "dropout counter row * N + col would wrap 32 bits (M * N >= 2^32)");

  // CTA pairs (256 x BN tiles, tcgen05.mma.cta_group::2) unless NBEST_GEMM_CTA_GROUP=1 asks for the single-CTA kernel.
  const int CG = ctx->knobs.gemm_cta_group;
  // persistent work units: CTAs or CTA pairs. reserve_sms leaves SMs to a collective running next to the GEMM
  // (nbest_ctx_set_sm_reserve): statically scheduled pairs would otherwise queue behind the SMs it occupies.
  const int units = (ctx->num_sms - ctx->reserve_sms) / CG;
  const int tile_m = BM * CG;
  // BN = 256 halves the smem operand traffic per MMA. Pairs: always preferred (out-proj T x 768 x 768: 21 us vs 32 us
  // with BN = 128). Single-CTA kernel: short-K problems with few tiles prefer BN = 128 (more, shorter tiles).
  int BN = (N % 256 == 0) ? 256 : 128;
  if (CG == 1 && BN == 256 && epilogue != NBEST_EPI_ACCUM_F32 && K <= 1024 &&
      (int64_t)((M + tile_m - 1) / tile_m) * (N / 256) < 4LL * units)
    BN = 128;
  if ((ctx->knobs.gemm_force_bn == 128 || ctx->knobs.gemm_force_bn == 256) && N % ctx->knobs.gemm_force_bn == 0)
    BN = ctx->knobs.gemm_force_bn;
  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.num_m_tiles = (M + tile_m - 1) / tile_m;
  g.num_n_tiles = N / BN;
  g.num_kb = (K + BK - 1) / BK;
  g.num_splits = 1;
  if (epilogue == NBEST_EPI_ACCUM_F32) {
    // split the contraction so that (tiles x splits) fills ONE wave of units (measured best on B200 — a second wave or
    // more splits only add fp32 atomic traffic), with >= 8 k-blocks per split
    const int tiles = g.num_m_tiles * g.num_n_tiles;
    int want = units / tiles;
    int max_by_k = g.num_kb / 8 > 0 ? g.num_kb / 8 : 1;
    if (want > max_by_k) want = max_by_k;
    if (want < 1) want = 1;
    if (ctx->knobs.wgrad_splits >= 1) want = ctx->knobs.wgrad_splits < max_by_k ? ctx->knobs.wgrad_splits : max_by_k;
    g.num_splits = want;
  }
  g.debug = ctx->knobs.gemm_debug;
  const bool two = epilogue == NBEST_EPI_BIAS_GELU;
  g.stages = (CG == 2) ? ((BN == 256) ? (two ? Cfg<256, 2, true>::kStages : Cfg<256, 2>::kStages)
                                      : (two ? Cfg<128, 2, true>::kStages : Cfg<128, 2>::kStages))
                       : ((BN == 256) ? (two ? Cfg<256, 1, true>::kStages : Cfg<256, 1>::kStages)
                                      : (two ? Cfg<128, 1, true>::kStages : Cfg<128, 1>::kStages));
  if (ctx->knobs.gemm_stages >= 1 && ctx->knobs.gemm_stages < g.stages) g.stages = ctx->knobs.gemm_stages;
  g.kb_per_split = (g.num_kb + g.num_splits - 1) / g.num_splits;
  g.num_splits = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;  // no empty split
  g.C = C;
  g.ldc = ldc;
  g.bias = bias;
  g.aux = reinterpret_cast<const __nv_bfloat16*>(aux_bf16);
  g.ldaux = ldaux;
  g.out2 = reinterpret_cast<__nv_bfloat16*>(out2_bf16);
  g.seed = seed;
  g.salt = nbest_salt(ctx);
  g.sched = ctx->gemm_dynamic ? ctx->sched_buf + 2 * (ctx->sched_seq++ % kSchedRing) : nullptr;
  if (p_drop > 0.f) {
    const double t = (double)p_drop * 65536.0 + 0.5;   // 16-bit threshold (ptx.cuh dropout_keep)
    g.drop_thresh = t >= 65535.0 ? 65535u : (uint32_t)t;
    g.drop_scale = 1.0f / (1.0f - p_drop);
  } else {
    g.drop_thresh = 0;
    g.drop_scale = 1.0f;
  }

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn_major)
    rc = nbest_make_tmap_bf16(ctx, &tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM);
  else
    rc = nbest_make_tmap_bf16(ctx, &tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64);
  if (rc != NBEST_OK) return rc;
  if (!b_mn_major)
    rc = nbest_make_tmap_bf16(ctx, &tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)(BN / CG));
  else
    rc = nbest_make_tmap_bf16(ctx, &tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64);
  if (rc != NBEST_OK) return rc;

  // output tensor maps for the epilogue's TMA stores: {64 columns x 32 rows} boxes of the bf16 result(s)
  CUtensorMap tmC = tmA, tmC2 = tmA;
  if (epilogue != NBEST_EPI_ACCUM_F32) {
    rc = nbest_make_tmap_bf16(ctx, &tmC, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32);
    if (rc != NBEST_OK) return rc;
    if (epilogue == NBEST_EPI_BIAS_GELU && out2_bf16) {
      rc = nbest_make_tmap_bf16(ctx, &tmC2, out2_bf16, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32);
      if (rc != NBEST_OK) return rc;
    }
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (CG == 2) {
    if (BN == 256) return dispatch<256, 2>(ctx, a_mn_major, b_mn_major, epilogue, tmA, tmB, tmC, tmC2, g, s);
    return dispatch<128, 2>(ctx, a_mn_major, b_mn_major, epilogue, tmA, tmB, tmC, tmC2, g, s);
  }
  if (BN == 256) return dispatch<256, 1>(ctx, a_mn_major, b_mn_major, epilogue, tmA, tmB, tmC, tmC2, g, s);
  return dispatch<128, 1>(ctx, a_mn_major, b_mn_major, epilogue, tmA, tmB, tmC, tmC2, g, s);
}
