// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (MMA / TMEM alloc / ld / commit / fences) and UMMA descriptors.
// Hand-written for this project; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace modcr {

#define DEVINL __device__ __forceinline__

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVINL uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------ TMA
DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
DEVINL void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
DEVINL void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   (uint64_t)m),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
DEVINL void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store / UMMA reads)
DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
template <int NCOLS>
DEVINL void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
DEVINL void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, cta_group::1
DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete
DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread = lane/row)
DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset inside CTA `rank` of this cluster
DEVINL uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Same without memory-ordering side effects (the release form costs a MEMBAR.ALL.GPU per arrive).  For hand-offs
// that only order tcgen05 traffic, which tcgen05.fence::before_thread_sync / tcgen05.wait::ld already cover.
DEVINL void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are signalled on a barrier of the pair's
// leader CTA (`bar_cluster_addr` = mapa_rank(barrier, 0))
DEVINL void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"((uint64_t)m), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int NCOLS>
DEVINL void tmem_alloc_pair(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
DEVINL void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 (128 rows per CTA), each CTA holding half of the B rows
DEVINL void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in every CTA of `mask`
DEVINL void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                   "r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version (1)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand tile [rows x 64 bf16] (128 B per row, 8-row 1024 B atoms):
//   SBO = 1024 (next 8-row group), LBO unused; advance K by 16 elems = +32 B.
// MN-major operand tile [64 k-rows x 64 mn] blocks (8 KB each for BLOCK_K=64):
//   SBO = 1024 (next 8 k-rows), LBO = bytes between 64-wide MN blocks;
//   advance K by 16 rows = +2048 B.
DEVINL uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 A/B, fp32 accumulate.
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major (1 = MN)   [16] b_major (1 = MN)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------ misc math
DEVINL uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
DEVINL float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
DEVINL float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

DEVINL float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
DEVINL float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

DEVINL float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Fast erf-GELU: Phi(x) = 0.5 (1 + erf(x / sqrt 2)) = 0.5 (1 + tanh(p(x))) with p = atanh(erf(x / sqrt 2)) fitted by
// the odd quintic x (c0 + c1 x^2 + c2 x^4) (minimax on the GELU error over |x| <= 8: |Phi error| < 5.1e-5,
// |gelu error| < 2.6e-5; x^2 is clamped where tanh has saturated so the negative c2 never bends p back).
// With MUFU.TANH (relative error 2^-11) the total |gelu error| stays below 2.5e-4 |x|, an order of magnitude
// under the bf16 rounding of the stored activation.  7 FP ops + 1 MUFU per element (the Abramowitz-Stegun
// form used before needed 15 + 2 and made the FFN epilogues issue-bound, ncu: profiles/).
DEVINL float gelu_tanh_arg(float x, float& x2) {
  x2 = x * x;
  const float x2c = fminf(x2, 50.0f);
  float q = fmaf(x2c, -3.51517510e-04f, 3.70056508e-02f);
  q = fmaf(x2c, q, 7.97507879e-01f);
  return x * q;
}
DEVINL float gelu_fast(float x) {
  float x2;
  const float t = fast_tanh(gelu_tanh_arg(x, x2));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// GELU'(x) = Phi(x) + x phi(x), phi(x) = exp(-x^2 / 2) / sqrt(2 pi)
DEVINL float gelu_grad_fast(float x) {
  float x2;
  const float t = fast_tanh(gelu_tanh_arg(x, x2));
  const float cdf = fmaf(0.5f, t, 0.5f);
  const float e = fast_ex2(-0.72134752044448170f * x2);
  return fmaf(x * e, 0.3989422804014327f, cdf);
}

// Two elements per instruction: FMUL2 / FFMA2 of sm_100 (`mul/fma.rn.f32x2`) halve the FP32 issue slots of the GELU
// epilogues, which are issue-bound with two epilogue warps per scheduler (ncu, profiles/r02_ncu_gemm_layer_summary.txt:
// dgrad ffn2 x GELU' issue-active 45 %, tensor pipe 47 %).  Each lane rounds exactly like the scalar form above.
DEVINL float2 gelu_tanh_arg2(float2 x, float2& x2) {
  x2 = __fmul2_rn(x, x);
  const float2 x2c = make_float2(fminf(x2.x, 50.0f), fminf(x2.y, 50.0f));
  float2 q = __ffma2_rn(x2c, make_float2(-3.51517510e-04f, -3.51517510e-04f), make_float2(3.70056508e-02f, 3.70056508e-02f));
  q = __ffma2_rn(x2c, q, make_float2(7.97507879e-01f, 7.97507879e-01f));
  return __fmul2_rn(x, q);
}
DEVINL float2 gelu_fast2(float2 x) {
  float2 x2;
  const float2 a = gelu_tanh_arg2(x, x2);
  const float2 t = make_float2(fast_tanh(a.x), fast_tanh(a.y));
  const float2 hx = __fmul2_rn(make_float2(0.5f, 0.5f), x);
  return __ffma2_rn(hx, t, hx);
}
DEVINL float2 gelu_grad_fast2(float2 x) {
  float2 x2;
  const float2 a = gelu_tanh_arg2(x, x2);
  const float2 t = make_float2(fast_tanh(a.x), fast_tanh(a.y));
  const float2 cdf = __ffma2_rn(make_float2(0.5f, 0.5f), t, make_float2(0.5f, 0.5f));
  const float2 ea = __fmul2_rn(make_float2(-0.72134752044448170f, -0.72134752044448170f), x2);
  const float2 e = make_float2(fast_ex2(ea.x), fast_ex2(ea.y));
  return __ffma2_rn(__fmul2_rn(x, e), make_float2(0.3989422804014327f, 0.3989422804014327f), cdf);
}

// explicit shared-state-space 128-bit accesses (a generic pointer derived from the aligned dynamic-smem base makes
// the compiler emit generic LD/ST, which issue slower than LDS/STS)
DEVINL void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
DEVINL void sts128f(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
DEVINL uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
DEVINL float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ------------------------------------------------------------------ dropout masks
// Counter-based: keep(seed, site, idx) is a pure function of the step's seed (read from device memory so that a
// replayed CUDA graph draws fresh masks), the call site and the element's logical index, so the backward pass
// regenerates the mask of its forward instead of storing it.  32-bit integer hash ("lowbias32" finaliser).
DEVINL uint32_t drop_hash(uint32_t seed, uint32_t site, uint32_t idx) {
  uint32_t x = (idx ^ seed) * 0x9E3779B1u;
  x ^= site * 0x85EBCA77u;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
// keep probability 1 - p  <=>  hash >= p * 2^32
DEVINL bool drop_keep(uint32_t seed, uint32_t site, uint32_t idx, uint32_t thr) { return drop_hash(seed, site, idx) >= thr; }
// Two-level form for the attention kernels, where the per-element cost matters (they are issue-bound): the row part
// rowh = drop_hash(seed, site, row) is computed once per thread and item, the element part is a Weyl step on the key id
// followed by one multiply-xorshift round (5 integer ops per element instead of ~16).
DEVINL bool drop_keep_row(uint32_t rowh, uint32_t kid, uint32_t thr) {
  uint32_t x = rowh ^ (kid * 0x9E3779B1u);
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= x >> 13;
  return x >= thr;
}

// byte offset of 16-byte piece `piece` (0..7) of row `row` inside a 128B-swizzled
// [rows x 128 B] tile whose base is 1024 B aligned
DEVINL uint32_t sw128_offset(uint32_t row, uint32_t piece) { return row * 128u + ((piece ^ (row & 7u)) << 4); }
// same for a 64B-swizzled [rows x 64 B] tile (4 pieces per row; address bits [4,6) ^= bits [7,9)), base 512 B aligned
DEVINL uint32_t sw64_offset(uint32_t row, uint32_t piece) { return row * 64u + ((piece ^ ((row >> 1) & 3u)) << 4); }

}  // namespace modcr
