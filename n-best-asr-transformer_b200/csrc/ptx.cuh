// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM), ldmatrix / mma.sync.
// Everything here is device-side and header-only; the kernels in this directory include it.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

namespace nbest {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization (common.h nbest_launch): its CTAs may
// become resident while the previous kernel of the stream is still draining, run their prologue, and block here until
// that kernel has completed and its memory is visible. Right behind the wait the kernel releases ITS dependent, so at most
// one successor is ever staged. Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- per-step state (CUDA-graph replay)
// A captured graph of the training step bakes every by-value kernel argument in, the dropout seeds among them. What must
// change from one replay to the next lives in a small device-side record owned by the context (common.h nbest_step_state,
// written by nbest_ctx_set_step_state on the stream ahead of the graph): every dropout kernel XORs `salt` into its seed
// (0 outside graph replay: the by-value seed alone decides), the optimizer reads its schedule multiplier from it.
// Read AFTER pdl_grid_sync (the writer may be the kernel right before this one) and past L1 (.cg).
__device__ __forceinline__ uint32_t step_salt(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (no suspend): for issuer warps that poll several barriers and serve whichever completes first
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load, global -> shared, completion on an mbarrier (complete_tx bytes).
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled store, shared -> global (bulk async-group completion).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate. One thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major  tile (rows = M/N index, 64 bf16 = 128 B per row): LBO ignored (1), SBO = 1024 B (8 rows).
//   MN-major tile (rows = K index, 64 bf16 of M/N per 128-B row): SBO = 1024 B (8 k-rows),
//            LBO = byte distance between consecutive 64-element M/N chunks.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2): clusters of two CTAs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA's window) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Relaxed: the arrive only orders tcgen05 traffic (handled by tcgen05.fence), no generic-proxy data is handed over; the
// default .release.cluster form costs a MEMBAR.ALL + ERRBAR per arrive (12 % of the GELU GEMM's stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Cluster-scope hand-over of a value through a peer CTA's shared memory: store, then release-arrive on the peer's
// barrier; the consumer pairs it with an acquire.cluster wait.
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// TMA load whose completion is signalled on an mbarrier that may live in the PEER CTA of the pair (bar_cluster_addr is
// a shared::cluster address); the data lands in this CTA's shared memory.
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* m, uint32_t bar_cluster_addr, void* smem_dst, int32_t c0,
                                                int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs: 128 rows each] * B[smem of both CTAs: N/2 rows each]; issued by ONE
// thread of the leader CTA (cluster rank 0) for the pair.
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the mbarrier at this offset in EVERY CTA of cta_mask once all previously issued MMAs have completed.
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---------------------------------------------------------------- legacy warp MMA (attention, short sequences)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t saddr, const void* gptr, bool pred) {
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- small math helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// Counter-based dropout RNG: ONE hash yields the keep/drop decisions of a QUAD of elements (4 x 16-bit lanes).
//   x = fold(q * C1) ^ seed;  x ^= x >> 16;  (a, b) = (fold(x * C2), fold(x * C3)),  fold(m) = lo32(m) ^ hi32(m)
// i.e. three 32x32->64 multiplies and four logic ops per four elements (the murmur3 finaliser it replaces cost 9
// instructions per PAIR and was the largest single item of the attention kernels' instruction mix). The host passes
// avalanche-mixed seeds (ops._seed), so per-(step, layer, site) streams are unrelated. Checked on the host restatement
// (tests/test_abi_and_host.py): uniform lanes, no lane / stride / cross-seed correlation, geometric gaps.
// thr16 = round(p_drop * 65536); element kept iff its 16-bit lane >= thr16. Forward and backward regenerate the same
// mask from (seed, element index).
__device__ __forceinline__ uint32_t fold64(uint64_t m) { return static_cast<uint32_t>(m) ^ static_cast<uint32_t>(m >> 32); }
__device__ __forceinline__ uint2 dropout_quad(uint32_t seed, uint32_t quad_idx) {
  uint32_t x = fold64(static_cast<uint64_t>(quad_idx) * 0x9E3779B1u) ^ seed;
  x ^= x >> 16;
  return make_uint2(fold64(static_cast<uint64_t>(x) * 0x85EBCA6Bu), fold64(static_cast<uint64_t>(x) * 0xC2B2AE35u));
}
__device__ __forceinline__ uint32_t quad_lane(uint2 h, uint32_t lane) {   // lane in [0, 4)
  const uint32_t w = (lane & 2u) ? h.y : h.x;
  return (lane & 1u) ? (w >> 16) : (w & 0xFFFFu);
}
// elements idx4 .. idx4 + 3 (idx4 a multiple of 4)
__device__ __forceinline__ void dropout_keep4(uint32_t seed, uint32_t idx4, uint32_t thr16, bool& k0, bool& k1, bool& k2,
                                              bool& k3) {
  const uint2 h = dropout_quad(seed, idx4 >> 2);
  k0 = (h.x & 0xFFFFu) >= thr16;
  k1 = (h.x >> 16) >= thr16;
  k2 = (h.y & 0xFFFFu) >= thr16;
  k3 = (h.y >> 16) >= thr16;
}
__device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t idx, uint32_t thr16) {
  return quad_lane(dropout_quad(seed, idx >> 2), idx & 3u) >= thr16;
}
// Attention probabilities: element (head h, global query token tq, key j inside the sequence, j < 512). The quad of a
// key is chosen so that the four accumulator elements an mma.sync thread holds in one 16-key group (keys c, c+1, c+8,
// c+9 of a query row) share ONE hash, and so do the two keys (j, j+8) a thread of the S^T formulation holds per query.
__device__ __forceinline__ uint32_t attn_quad_row(int h, int T, int tq) { return ((uint32_t)h * (uint32_t)T + (uint32_t)tq) * 128u; }
__device__ __forceinline__ uint32_t attn_quad(int h, int T, int tq, int j) {
  return attn_quad_row(h, T, tq) + (((uint32_t)j >> 4) << 2) + (((uint32_t)j & 7u) >> 1);
}
__device__ __forceinline__ uint32_t attn_lane(int j) { return ((uint32_t)j & 1u) | (((uint32_t)j >> 2) & 2u); }
__device__ __forceinline__ bool attn_dropout_keep(uint32_t seed, int h, int T, int tq, int j, uint32_t thr16) {
  return quad_lane(dropout_quad(seed, attn_quad(h, T, tq, j)), attn_lane(j)) >= thr16;
}

// erf-based GELU pieces. Phi(z) = 0.5*(1+erf(z/sqrt2)) via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7 + approx-rcp/ex2
// error ~1e-6), phi(z) = exp(-z^2/2)/sqrt(2*pi). Far below bf16 resolution of the result; ~16 instructions, 2 MUFU.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// w = h / e with h = 0.5*(1 - erf(|z|/sqrt2)) (the polynomial part of A&S 7.1.26) and e = exp(-z^2/2); xs = |z| *
// sqrt(log2(e)/2) so that e = 2^(-xs^2).
__device__ __forceinline__ void gelu_tail(float z, float& w, float& e) {
  const float xs = fabsf(z) * 0.8493218002880191f;
  const float t = rcp_approx(fmaf(0.2727374808792225f, xs, 1.0f));   // 1 / (1 + 0.3275911 |z|/sqrt2)
  e = ex2_approx(-xs * xs);
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  w = poly * t;
}
// z*Phi(z) with Phi = step(z) - sign(z) h:  relu(z) - |z| w e      (12 instructions, 2 MUFU)
__device__ __forceinline__ float gelu_fwd(float z) {
  float w, e;
  gelu_tail(z, w, e);
  return fmaf(-fabsf(z) * w, e, fmaxf(z, 0.f));
}
// both at once (shared polynomial / exponential): g = z Phi(z), d = Phi(z) + z phi(z)       (17 instructions, 2 MUFU)
__device__ __forceinline__ void gelu_fwd_grad(float z, float& g, float& d) {
  float w, e;
  gelu_tail(z, w, e);
  const float a = fabsf(z);
  g = fmaf(-a * w, e, fmaxf(z, 0.f));
  const float dd = fmaf(0.39894228040143268f, a, -w);
  const float ds = __uint_as_float(__float_as_uint(dd) ^ (__float_as_uint(z) & 0x80000000u));
  d = fmaf(ds, e, __float_as_int(z) >= 0 ? 1.0f : 0.f);
}
// Phi(z) + z phi(z) = step(z) + sign(z) e (|z|/sqrt(2 pi) - w)
__device__ __forceinline__ float gelu_grad(float z) {
  float w, e;
  gelu_tail(z, w, e);
  const float d = fmaf(0.39894228040143268f, fabsf(z), -w);
  const float ds = __uint_as_float(__float_as_uint(d) ^ (__float_as_uint(z) & 0x80000000u));   // sign(z) * d
  return fmaf(ds, e, __float_as_int(z) >= 0 ? 1.0f : 0.f);   // step by SIGN BIT: +0 -> 1 - h(0) = 0.5, -0 -> 0 + h(0)
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue two IEEE fp32 operations per instruction) -----------
// The GELU epilogue of the FFN-in GEMM is bound by the instruction issue of its epilogue warps (ncu round 2: 40.8 M warp
// instructions, 47 % tensor-pipe activity against 72 % for the same FLOPs with a plain epilogue); evaluating gelu and gelu'
// on element PAIRS halves every FMUL / FFMA / FADD of it. Same formulas, same rounding per element as the scalar versions.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_splat(float c) { return f2_pack(c, c); }
// g = z Phi(z), d = Phi(z) + z phi(z) for the pair z = (z0, z1). With a = |z|, xs = a sqrt(log2(e)/2), t = 1/(1 + p xs),
// e = 2^(-xs^2) = exp(-z^2/2), w' = -poly(t) t (so that w' e = -0.5 erfc(a/sqrt2) = -Q):
//     Phi = 0.5 + sign(z) (0.5 - Q),   g = z Phi,   d = Phi + (z e) / sqrt(2 pi)
// (Abramowitz-Stegun 7.1.26 as gelu_tail above: |err| < 1.5e-7 + approx-rcp / ex2 error ~1e-6.)
__device__ __forceinline__ void gelu_fwd_grad2(f32x2 z, f32x2& g, f32x2& d) {
  const f32x2 a = z & 0x7FFFFFFF7FFFFFFFull;
  const f32x2 sgn = z & 0x8000000080000000ull;
  const f32x2 xs = f2_mul(a, f2_splat(0.8493218002880191f));
  const f32x2 den = f2_fma(xs, f2_splat(0.2727374808792225f), f2_splat(1.0f));
  const f32x2 q2 = f2_mul(xs, xs);
  float d0, d1, q0, q1;
  f2_unpack(den, d0, d1);
  f2_unpack(q2, q0, q1);
  const f32x2 t = f2_pack(rcp_approx(d0), rcp_approx(d1));
  const f32x2 e = f2_pack(ex2_approx(-q0), ex2_approx(-q1));
  f32x2 poly = f2_fma(t, f2_splat(-0.5f * 1.061405429f), f2_splat(0.5f * 1.453152027f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * 1.421413741f));
  poly = f2_fma(poly, t, f2_splat(0.5f * 0.284496736f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * 0.254829592f));
  const f32x2 wn = f2_mul(poly, t);                                   // -w
  const f32x2 h = f2_fma(wn, e, f2_splat(0.5f));                      // 0.5 - Q  (>= 0)
  const f32x2 phi_cdf = f2_add(h ^ sgn, f2_splat(0.5f));              // 0.5 + sign(z) (0.5 - Q)
  g = f2_mul(z, phi_cdf);
  d = f2_fma(f2_mul(z, e), f2_splat(0.39894228040143268f), phi_cdf);
}
__device__ __forceinline__ f32x2 gelu_fwd2(f32x2 z) {
  f32x2 g, d;
  gelu_fwd_grad2(z, g, d);
  return g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace nbest
