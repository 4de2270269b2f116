// Persistent, warp-specialised bf16 GEMM on tcgen05 / TMEM / TMA for sm_100a.
//
//   C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
//
// Operand storage ("major") is selected per operand:
//   A K-major : A stored [M,K] row-major          A MN-major: A stored [K,M] row-major
//   B K-major : B stored [N,K] row-major          B MN-major: B stored [K,N] row-major
// which gives forward (K,K), dgrad (K,MN: dX = dY * W with W stored [out,in]) and
// wgrad (MN,MN: dW = dY^T * X, contraction over the row dimension of both).
//
// Roles (384 threads, 1 CTA / SM, persistent over tiles):
//   warp 0  TMA producer for the A/B smem ring            (1 elected thread)
//   warp 1  tcgen05.mma issuer, accumulators in TMEM      (1 elected thread)
//   warp 2  TMEM allocator + TMA loader of the optional epilogue input tile
//   warps 2-3  (wgrad with a bias gradient) column sums of the staged A tiles: dW = dY^T X stages dY anyway, so
//           db = dY^T 1 is accumulated from shared memory by the two otherwise idle warps of the n == 0 tiles
//   warps 4-11 epilogue in two groups of four warps (group g takes the 32-column chunks g, g+2, ...):
//           tcgen05.ld -> bias / activation / residual-add / GELU' -> swizzled smem -> TMA store (or TMA
//           reduce-add for split-K fp32); bf16 chunks ping-pong between two buffers per group
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// NCTA = 2 runs the same roles on CTA pairs (cluster of 2, tcgen05 cta_group::2): the pair owns a 256 x BN
// output tile, each CTA loads its own 128 A rows and HALF of the B tile (BN/2 rows), the leader CTA issues
// M=256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM, and each CTA runs the epilogue
// of its own 128 rows.  Per CTA that is 32 KB of operand traffic per k-block instead of 48 KB for the same
// 128 x 256 x 64 MACs, which is what lifts the kernel off the L2->SM bandwidth limit (ncu: profiles/).
#include "gemm.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tma_host.h"

namespace modcr {

struct GemmParams {
  int M, N, K;
  int kb_total;   // ceil(K / 64)
  int kb_per;     // k-blocks per split
  int split_k;
  int m_tiles, n_tiles;
  float alpha;
  const float* bias;
  int act;        // 0 none, 1 gelu(erf), 2 tanh, 3 relu
  int in_op;      // 0 none, 1 add, 2 multiply by gelu'(in)
  int reduce_add; // output through TMA reduce-add (fp32 accumulate in place)
  float* a_colsum; // A MN-major only: a_colsum[m] += sum_k A[k,m] from the staged A tiles of the n == 0 tiles
  // training-mode dropout of f(acc + bias) BEFORE the epilogue input is applied (LayerNorm(dropout(dense(x)) + residual)):
  // keep = hash(*drop_seed, drop_site, m * N + n) >= drop_thr; drop_thr == 0: off
  const uint32_t* drop_seed;
  uint32_t drop_site, drop_thr;
  float drop_inv;
  // in_op 3: residual = LayerNorm(in) rebuilt from the pre-LayerNorm tensor, its row statistics and the affine parameters
  const float* in_mean;
  const float* in_rstd;
  const float* in_gamma;
  const float* in_beta;
};

template <int NCTA, int BN, bool A_MN, bool B_MN, bool HAS_AUX, bool HAS_IN, bool OUT_F32, bool WIDE = false>
struct Cfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int BN_LOAD = BN / NCTA;          // B rows this CTA loads
  static constexpr int B_BYTES = BN_LOAD * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // Epilogue chunks are 128 rows x 64 bytes (SWIZZLE_64B): 32 bf16 or 16 fp32 columns.  Outputs (and the aux copy)
  // and the epilogue-input tiles are double-buffered per epilogue group, so a chunk is written while the TMA store
  // of the previous one still reads its buffer, the input of chunk k+1 is in flight while chunk k is processed, and
  // one block barrier per chunk suffices.
  static constexpr int ES = OUT_F32 ? 4 : 2;
  // WIDE (fp32 output with an fp32 epilogue input, K <= 1024): 32-column fp32 chunks = 128-byte rows (SWIZZLE_128B), half
  // as many chunk hand-offs (block barrier + TMA store + two mbarrier round trips each) per tile, paid for with a
  // 3-stage instead of a 4-stage operand ring -- the short-K dense-output GEMM is bound by its epilogue, not its ring
  static constexpr int CHUNK_COLS = OUT_F32 ? (WIDE ? 32 : 16) : 32;
  static constexpr int ROW_BYTES = CHUNK_COLS * ES;
  static constexpr int CHUNK_BYTES = 128 * ROW_BYTES;
  static constexpr int NCHUNK = BN / CHUNK_COLS;
  static constexpr int NBUF = 2;                      // out (and aux) buffers per group
  static constexpr int IN_DEPTH = 2;                  // epilogue-input buffers per group
  // (A variant of the fp32-residual epilogue without shared-memory staging -- every thread reading / writing the 64
  // contiguous bytes of its own row straight from / to global memory -- was measured SLOWER, 79 vs 50 us for the K = 768
  // dense-output GEMM: a row-per-thread 128-bit access touches 32 different lines per warp instruction and the L1 /
  // LSU wavefronts, not the barriers, become the limit.  profiles/r02_bench_direct_epilogue.log)
  static constexpr int OUT_BUFS = 2 * NBUF * (HAS_AUX ? 2 : 1);
  static constexpr int IN_BUFS = HAS_IN ? 2 * IN_DEPTH : 0;
  static constexpr int BIAS_BYTES = 2 * 128 * 4;      // per group: the 128 bias values of its chunks of one tile
  static constexpr int EPI_BYTES = (OUT_BUFS + IN_BUFS) * CHUNK_BYTES + BIAS_BYTES;
  static constexpr int BAR_BYTES = 320;
  static constexpr int SMEM_MAX = 232448;  // 227 KB
  static constexpr int STAGES_RAW = (SMEM_MAX - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
  static_assert(!(OUT_F32 && HAS_AUX), "fp32 output has no aux path");
  static_assert(NCTA == 1 || NCTA == 2, "CTA pairs at most");
};

struct TileCoord {
  int m, n, kb0, kb1, split;
};

// `m` counts 128 * NCTA row groups (p.m_tiles of them)
DEVINL TileCoord decode_tile(const GemmParams& p, int t) {
  TileCoord c;
  const int mn = p.m_tiles * p.n_tiles;
  c.split = t / mn;
  const int r = t - c.split * mn;
  c.m = r / p.n_tiles;
  c.n = r - c.m * p.n_tiles;
  c.kb0 = c.split * p.kb_per;
  c.kb1 = min(c.kb0 + p.kb_per, p.kb_total);
  return c;
}

template <int NCTA, int BN, bool A_MN, bool B_MN, bool HAS_AUX, bool HAS_IN, bool OUT_F32, bool WIDE = false>
__global__ void __launch_bounds__(384, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
            const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  using C = Cfg<NCTA, BN, A_MN, B_MN, HAS_AUX, HAS_IN, OUT_F32, WIDE>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * C::A_BYTES;
  uint8_t* sOut = sB + STAGES * C::B_BYTES;
  uint8_t* sIn = sOut + C::OUT_BUFS * C::CHUNK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + C::IN_BUFS * C::CHUNK_BYTES + C::BIAS_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* infull = tempty + 2;     // [group][slot]
  uint64_t* inempty = infull + 4;
  uint64_t* consumed = inempty + 4;   // [STAGES] (column-sum warps) the MMAs reading the stage have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(consumed + C::STAGES);
  float* sBiasG = reinterpret_cast<float*>(sIn + C::IN_BUFS * C::CHUNK_BYTES);   // [2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;
  // work is distributed over CTA groups (single CTAs or pairs); `rank` is this CTA's half of the pair
  const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;
  const int group = blockIdx.x / NCTA;
  const int ngroups = gridDim.x / NCTA;
  const bool leader = rank == 0;
  const bool do_colsum = A_MN && !HAS_IN && p.a_colsum != nullptr;   // uniform over the grid
  constexpr int EPI_ARRIVALS = NCTA == 2 ? 16 : 256;   // pairs: one (remote) arrive per epilogue warp

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (HAS_AUX) tma_prefetch_desc(&tmAux);
    if (HAS_IN) tma_prefetch_desc(&tmIn);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], do_colsum ? 2 : 1);   // MMA commit, or the two column-sum warps (which follow the MMAs)
      mbar_init(&consumed[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPI_ARRIVALS);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&infull[i], 1);
      mbar_init(&inempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (NCTA == 2) tmem_alloc_pair<C::TMEM_COLS>(tmem_slot);
    else tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: MMA operands live in uniform registers

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = group; t < total_tiles; t += ngroups) {
        const TileCoord tc = decode_tile(p, t);
        const int m0 = (tc.m * NCTA + rank) * 128;            // this CTA's A rows
        const int n0 = tc.n * BN + rank * C::BN_LOAD;         // this CTA's share of the B rows
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          if constexpr (NCTA == 2) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the pair
            if (leader) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
            const uint32_t bar = mapa_rank(smem_u32(&full[stage]), 0);
            if (!A_MN) {
              tma_load_2d_pair(a_dst, &tmA, bar, kb * 64, m0);
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) tma_load_2d_pair(a_dst + i * 8192, &tmA, bar, m0 + i * 64, kb * 64);
            }
            if (!B_MN) {
              tma_load_2d_pair(b_dst, &tmB, bar, kb * 64, n0);
            } else {
#pragma unroll
              for (int i = 0; i < C::BN_LOAD / 64; ++i)
                tma_load_2d_pair(b_dst + i * 8192, &tmB, bar, n0 + i * 64, kb * 64);
            }
          } else {
            mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(a_dst, &tmA, &full[stage], kb * 64, m0);
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) tma_load_2d(a_dst + i * 8192, &tmA, &full[stage], m0 + i * 64, kb * 64);
            }
            if (!B_MN) {
              tma_load_2d(b_dst, &tmB, &full[stage], kb * 64, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(b_dst + i * 8192, &tmB, &full[stage], n0 + i * 64, kb * 64);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (pairs: leader CTA only)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(128 * NCTA, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = group; t < total_tiles; t += ngroups, ++it) {
        const TileCoord tc = decode_tile(p, t);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_desc = A_MN ? make_smem_desc(a_base + k * 2048, 8192, 1024)
                                         : make_smem_desc(a_base + k * 32, 0, 1024);
            const uint64_t b_desc = B_MN ? make_smem_desc(b_base + k * 2048, 8192, 1024)
                                         : make_smem_desc(b_base + k * 32, 0, 1024);
            if (NCTA == 2) umma_bf16_pair(d_tmem, a_desc, b_desc, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire; with a bias gradient the slot goes to
          // the column-sum warps first, which release it
          uint64_t* rel = do_colsum ? &consumed[stage] : &empty[stage];
          if (NCTA == 2) umma_commit_pair(rel, 3);
          else umma_commit(rel);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (NCTA == 2) umma_commit_pair(&tfull[as], 3);
        else umma_commit(&tfull[as]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ epilogue-input loader
    // chunk c of a tile belongs to epilogue group (c & 1); each group owns IN_DEPTH input buffers
    if (HAS_IN && lane == 0) {
      uint32_t cntg[2] = {0, 0};
      for (int t = group; t < total_tiles; t += ngroups) {
        const TileCoord tc = decode_tile(p, t);
        const int m0 = (tc.m * NCTA + rank) * 128;
        for (int c = 0; c < C::NCHUNK; ++c) {
          const int g = c & 1;
          const uint32_t slot = cntg[g] % C::IN_DEPTH;
          const uint32_t ph = (cntg[g] / C::IN_DEPTH) & 1;
          const int bi = g * C::IN_DEPTH + slot;
          mbar_wait(&inempty[bi], ph ^ 1);
          mbar_expect_tx(&infull[bi], C::CHUNK_BYTES);
          tma_load_2d(sIn + bi * C::CHUNK_BYTES, &tmIn, &infull[bi], tc.n * BN + c * C::CHUNK_COLS, m0);
          ++cntg[g];
        }
      }
    }
  }
  if ((warp == 2 || warp == 3) && do_colsum) {
    // ------------------------------------------------------------ column sums of the staged A tiles
    // A stage is two [64 k-rows x 64 m] 128B-swizzled boxes.  Lane l owns m-columns 64 (l >> 4) + 4 (l & 15) .. + 4
    // (half a 16-byte piece); warp 2 sums k-rows [0,32), warp 3 rows [32,64) of every stage.  Every tile's stages
    // are released by these warps as well (the empty barrier counts them), but only n == 0 tiles are summed.
    const int r0 = (warp - 2) * 32;
    const int box = lane >> 4, piece = (lane & 15) >> 1, halfp = lane & 1;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = group; t < total_tiles; t += ngroups) {
      const TileCoord tc = decode_tile(p, t);
      const bool mine = tc.n == 0;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
        mbar_wait(&consumed[stage], phase);   // the tile landed and the MMAs are done with it
        if (mine) {
          const uint32_t base = smem_u32(sA + stage * C::A_BYTES) + box * 8192 + halfp * 8;
          // k-rows beyond K were zero-filled by TMA
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const uint32_t rr = r0 + r;
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                         : "=r"(lo), "=r"(hi)
                         : "r"(base + rr * 128 + ((piece ^ (rr & 7)) << 4)));
            const float2 a = unpack_bf16(lo), b = unpack_bf16(hi);
            acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
          }
        }
        fence_async_smem();   // generic-proxy reads of the stage before the TMA load that refills it (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (mine) {
        const int m = (tc.m * NCTA + rank) * 128 + box * 64 + piece * 8 + halfp * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (m + j < p.M) atomicAdd(p.a_colsum + m + j, acc[j]);
      }
    }
  }
  if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (2 groups x 4 warps)
    // group g handles chunks c = g, g+2, ... of every tile; both groups read the same TMEM stage.
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;          // TMEM lane quarter == warp % 4
    const int row = q * 32 + lane;   // row inside the 128-row tile
    const int gt = row;              // thread index inside the group
    const bool issuer = (q == 0 && lane == 0);
    const uint32_t bar_id = 1 + g;
    constexpr int CC = C::CHUNK_COLS;
    // byte offset of 16-byte piece `piece` of this thread's row inside a chunk buffer (64- or 128-byte rows)
    auto swz = [](uint32_t r, uint32_t piece) -> uint32_t {
      return C::ROW_BYTES == 128 ? sw128_offset(r, piece) : sw64_offset(r, piece);
    };
    uint8_t* out_base = sOut + g * C::NBUF * C::CHUNK_BYTES;
    uint8_t* aux_base = sOut + (2 * C::NBUF + g * C::NBUF) * C::CHUNK_BYTES;
    float* sBias = sBiasG + g * 128;
    const uint32_t tempty_r0 = NCTA == 2 ? mapa_rank(smem_u32(&tempty[0]), 0) : 0u;
    const uint32_t tempty_r1 = NCTA == 2 ? mapa_rank(smem_u32(&tempty[1]), 0) : 0u;
    int it = 0;
    uint32_t cnt = 0;  // chunks handled by this group
    for (int t = group; t < total_tiles; t += ngroups, ++it) {
      const TileCoord tc = decode_tile(p, t);
      const int m0 = (tc.m * NCTA + rank) * 128;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const bool add_bias = (p.bias != nullptr) && (tc.split == 0);
      const bool in_ln = OUT_F32 && HAS_IN && p.in_mean != nullptr;   // uniform
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (in_ln && m0 + row < p.M) {
        ln_mean = __ldg(p.in_mean + m0 + row);
        ln_rstd = __ldg(p.in_rstd + m0 + row);
      }
      if (add_bias) {
        // stage this tile's bias values of the group's chunks (thread i: chunk g + 2 (i / CC), column i % CC); every
        // thread of the group is past its last bias read of the previous tile (it passed that chunk's barrier)
        const int col = tc.n * BN + (g + 2 * (gt / CC)) * CC + (gt % CC);
        sBias[gt] = (add_bias && col < p.N) ? __ldg(p.bias + col) : 0.f;
        named_bar_sync(bar_id, 128);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = g, k = 0; c < C::NCHUNK; c += 2, ++k, ++cnt) {
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN + c * C::CHUNK_COLS;
        float v[CC];
        if constexpr (CC == 32) tmem_ld_32x32(taddr, reinterpret_cast<uint32_t*>(v));
        else tmem_ld_32x16(taddr, reinterpret_cast<uint32_t*>(v));
        // in_op 3: this chunk's gamma / beta (warp-uniform 128-bit loads, one L1 line each) are requested here, under the
        // TMEM load and the wait for the residual tile, instead of on the critical path behind that wait.  (Staging them
        // in shared memory next to the bias would cost this kernel a pipeline stage.)
        float4 lg[OUT_F32 ? CC / 4 : 1], lb[OUT_F32 ? CC / 4 : 1];
        if constexpr (OUT_F32 && HAS_IN) {
          const int nc = tc.n * BN + c * C::CHUNK_COLS;
          if (in_ln && nc < p.N) {
#pragma unroll
            for (int j = 0; j < CC / 4; ++j) {
              lg[j] = __ldg(reinterpret_cast<const float4*>(p.in_gamma + nc + 4 * j));
              lb[j] = __ldg(reinterpret_cast<const float4*>(p.in_beta + nc + 4 * j));
            }
          }
        }
        tmem_ld_wait();
        if (c + 2 >= C::NCHUNK) {  // this group's last read of the accumulator: hand the TMEM stage back
          tc_fence_before();
          if constexpr (NCTA == 2) {   // to the leader's barrier, one arrive per warp
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(as ? tempty_r1 : tempty_r0);
          } else {
            mbar_arrive(&tempty[as]);
          }
        }
        const int n0 = tc.n * BN + c * C::CHUNK_COLS;
        if (p.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < CC; ++j) v[j] *= p.alpha;
        }
        if (add_bias) {   // broadcast LDS: every lane reads the same address
#pragma unroll
          for (int j = 0; j < CC / 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + k * CC + 4 * j);
            v[4 * j + 0] += b4.x;
            v[4 * j + 1] += b4.y;
            v[4 * j + 2] += b4.z;
            v[4 * j + 3] += b4.w;
          }
        }
        const uint32_t buf = C::NBUF == 2 ? (cnt & 1u) : 0u;
        uint8_t* out_buf = out_base + buf * C::CHUNK_BYTES;
        uint8_t* aux_buf = aux_base + buf * C::CHUNK_BYTES;
        const uint32_t out_s = smem_u32(out_buf), aux_s = smem_u32(aux_buf);
        if (C::NBUF == 1) {
          // single buffer: the TMA store that last used it must have finished reading it
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);
        }
        if constexpr (HAS_AUX) {  // pre-activation copy (bf16) for the backward pass
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(aux_s + sw64_offset(row, j), pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < CC; j += 2) {
            const float2 r = gelu_fast2(make_float2(v[j], v[j + 1]));
            v[j] = r.x;
            v[j + 1] = r.y;
          }
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < CC; ++j) v[j] = tanhf(v[j]);
        } else if (p.act == 3) {
#pragma unroll
          for (int j = 0; j < CC; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.drop_thr != 0) {   // uniform
          const uint32_t dseed = __ldg(p.drop_seed);
          const uint32_t base = (uint32_t)(m0 + row) * (uint32_t)p.N + (uint32_t)n0;
#pragma unroll
          for (int j = 0; j < CC; ++j) v[j] = drop_keep(dseed, p.drop_site, base + j, p.drop_thr) ? v[j] * p.drop_inv : 0.f;
        }
        if (HAS_IN) {
          const uint32_t slot = cnt % C::IN_DEPTH;
          const int bi = g * C::IN_DEPTH + slot;
          const uint32_t in_s = smem_u32(sIn + bi * C::CHUNK_BYTES);
          mbar_wait(&infull[bi], (cnt / C::IN_DEPTH) & 1);
          if constexpr (OUT_F32) {  // fp32 residual stream: C = f(acc) + in (fp32 tile)
            if (in_ln && n0 < p.N) {   // ... + LayerNorm(in): same arithmetic as ln_fwd_kernel, (x - mean) * rstd * gamma + beta
#pragma unroll
              for (int j = 0; j < CC / 4; ++j) {
                const float4 f = lds128f(in_s + swz(row, j));
                const float4 g4 = lg[j], b4 = lb[j];
                v[4 * j + 0] += (f.x - ln_mean) * ln_rstd * g4.x + b4.x;
                v[4 * j + 1] += (f.y - ln_mean) * ln_rstd * g4.y + b4.y;
                v[4 * j + 2] += (f.z - ln_mean) * ln_rstd * g4.z + b4.z;
                v[4 * j + 3] += (f.w - ln_mean) * ln_rstd * g4.w + b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CC / 4; ++j) {
                const float4 f = lds128f(in_s + swz(row, j));
                v[4 * j + 0] += f.x;
                v[4 * j + 1] += f.y;
                v[4 * j + 2] += f.z;
                v[4 * j + 3] += f.w;
              }
            }
          } else if constexpr (!OUT_F32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 pk = lds128(in_s + sw64_offset(row, j));
              const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16(w[e]);
                if (p.in_op == 1) {
                  v[8 * j + 2 * e] += f.x;
                  v[8 * j + 2 * e + 1] += f.y;
                } else {
                  const float2 r = __fmul2_rn(make_float2(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]), gelu_grad_fast2(f));
                  v[8 * j + 2 * e] = r.x;
                  v[8 * j + 2 * e + 1] = r.y;
                }
              }
            }
          }
        }
        if constexpr (OUT_F32) {
#pragma unroll
          for (int j = 0; j < CC / 4; ++j)
            sts128f(out_s + swz(row, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(out_s + sw64_offset(row, j), pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        }
        fence_async_smem();
        if (HAS_IN) {
          // Hand the input slot back only AFTER the proxy fence: this thread's generic-proxy reads of the slot must be
          // ordered before the async-proxy write of the next TMA load into it.  (Arriving right behind the LDS let the
          // next load overwrite the slot under a slow reader about once per 10^7 chunks: the epilogue then added the
          // residual of the chunk 64 columns ahead -- found by scripts/gemm_race.py.)
          const uint32_t slot_r = cnt % C::IN_DEPTH;
          mbar_arrive(&inempty[g * C::IN_DEPTH + slot_r]);
        }
        // double-buffered: the store issued one chunk ago (other buffer) must have been read before anyone passes
        // this barrier and starts the next chunk, which overwrites that buffer; it had a whole chunk to drain
        if (C::NBUF == 2 && issuer) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        if (issuer) {
          if (n0 < p.N) {
            if (p.reduce_add)
              tma_reduce_add_2d(&tmC, out_buf, n0, m0);
            else
              tma_store_2d(&tmC, out_buf, n0, m0);
            if (HAS_AUX) tma_store_2d(&tmAux, aux_buf, n0, m0);
          }
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }
  tc_fence_before();
  if (NCTA == 2) {
    // neither CTA may exit (or free TMEM) while its peer can still arrive on its barriers / read its smem
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
  } else {
    __syncthreads();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------- host
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  const int lim = sm_limit();
  return (lim > 0 && lim < g_num_sms) ? lim : g_num_sms;
}

// MODCR_GEMM_NCTA=1 forces the single-CTA kernel (A/B comparison of the two data paths; default: CTA pairs)
static int g_ncta = 0;
static int gemm_ncta() {
  if (g_ncta == 0) {
    const char* e = getenv("MODCR_GEMM_NCTA");
    g_ncta = (e != nullptr && e[0] == '1') ? 1 : 2;
  }
  return g_ncta;
}

template <int NCTA, int BN, bool A_MN, bool B_MN, bool HAS_AUX, bool HAS_IN, bool OUT_F32, bool WIDE = false>
static int launch_n(const GemmDesc& g, cudaStream_t stream) {
  using C = Cfg<NCTA, BN, A_MN, B_MN, HAS_AUX, HAS_IN, OUT_F32, WIDE>;
  auto kern = gemm_kernel<NCTA, BN, A_MN, B_MN, HAS_AUX, HAS_IN, OUT_F32, WIDE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return MODCR_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmC, tmAux, tmIn;
  int rc;
  // A
  if (!A_MN) rc = make_tmap_2d(&tmA, g.A, 2, g.M, g.K, g.lda, 128, 64);
  else       rc = make_tmap_2d(&tmA, g.A, 2, g.K, g.M, g.lda, 64, 64);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d(&tmB, g.B, 2, g.N, g.K, g.ldb, C::BN_LOAD, 64);
  else       rc = make_tmap_2d(&tmB, g.B, 2, g.K, g.N, g.ldb, 64, 64);
  if (rc) return rc;
  rc = make_tmap_2d(&tmC, g.C, OUT_F32 ? 4 : 2, g.M, g.N, g.ldc, 128, C::CHUNK_COLS);
  if (rc) return rc;
  tmAux = tmC;
  tmIn = tmC;
  if (HAS_AUX) {
    rc = make_tmap_2d(&tmAux, g.aux, 2, g.M, g.N, g.ldaux, 128, C::CHUNK_COLS);
    if (rc) return rc;
  }
  if (HAS_IN) {
    rc = make_tmap_2d(&tmIn, g.epi_in, OUT_F32 ? 4 : 2, g.M, g.N, g.ld_in, 128, C::CHUNK_COLS);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.kb_total = (g.K + 63) / 64;
  int split = g.split_k < 1 ? 1 : g.split_k;
  if (split > p.kb_total) split = p.kb_total;
  p.kb_per = (p.kb_total + split - 1) / split;
  split = (p.kb_total + p.kb_per - 1) / p.kb_per;  // no empty splits
  p.split_k = split;
  p.m_tiles = (g.M + 128 * NCTA - 1) / (128 * NCTA);
  p.n_tiles = (g.N + BN - 1) / BN;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.act = g.act;
  p.in_op = g.in_op;
  p.reduce_add = (g.accumulate || split > 1) ? 1 : 0;
  p.a_colsum = g.a_colsum;
  p.in_mean = nullptr; p.in_rstd = nullptr; p.in_gamma = nullptr; p.in_beta = nullptr;
  if (g.in_op == 3) {
    if (!(OUT_F32 && HAS_IN) || g.N % 16 != 0 || g.in_mean == nullptr || g.in_rstd == nullptr || g.in_gamma == nullptr || g.in_beta == nullptr) {
      set_error("gemm: in_op 3 (LayerNorm of the epilogue input) needs fp32 C / input and the four LayerNorm pointers");
      return MODCR_ERR_UNSUPPORTED;
    }
    p.in_mean = g.in_mean; p.in_rstd = g.in_rstd; p.in_gamma = g.in_gamma; p.in_beta = g.in_beta;
  }
  p.drop_seed = g.drop_seed; p.drop_site = g.drop_site; p.drop_thr = 0; p.drop_inv = 1.f;
  if (g.drop_p > 0.f) {
    if (g.drop_seed == nullptr || !(g.drop_p < 1.f) || (double)g.M * g.N >= 4294967296.0 || g.split_k > 1 || g.accumulate) {
      set_error("gemm: epilogue dropout needs a device seed pointer, p < 1, M*N < 2^32 and no split-K / accumulate");
      return MODCR_ERR_UNSUPPORTED;
    }
    p.drop_thr = (uint32_t)fmin(4294967295.0, (double)g.drop_p * 4294967296.0);
    if (p.drop_thr == 0) p.drop_thr = 1;
    p.drop_inv = 1.0f / (1.0f - g.drop_p);
  }
  if (g.a_colsum != nullptr && (!A_MN || HAS_IN)) {
    set_error("gemm: a_colsum needs A stored [K,M] (a_mn = 1) and no epilogue input");
    return MODCR_ERR_UNSUPPORTED;
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int groups_max = num_sms() / NCTA;
  const int groups = total < groups_max ? total : groups_max;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * NCTA, 1, 1);
  cfg.blockDim = dim3(384, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmAux, tmIn, p);
  if (e != cudaSuccess) {
    set_error("gemm: launch failed: %s", cudaGetErrorString(e));
    return MODCR_ERR_CUDA;
  }
  return check_launch("gemm");
}

template <int BN, bool A_MN, bool B_MN, bool HAS_AUX, bool HAS_IN, bool OUT_F32, bool WIDE = false>
static int launch(const GemmDesc& g, cudaStream_t stream) {
  if (gemm_ncta() == 2) return launch_n<2, BN, A_MN, B_MN, HAS_AUX, HAS_IN, OUT_F32, WIDE>(g, stream);
  return launch_n<1, BN, A_MN, B_MN, HAS_AUX, HAS_IN, OUT_F32, WIDE>(g, stream);
}

// MODCR_GEMM_WIDE=0 disables the 32-column fp32 epilogue chunks of the short-K fp32-residual GEMM (A/B comparison)
static bool gemm_wide_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MODCR_GEMM_WIDE");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Same cost model as engine.py::_split_k: waves of (tile, split) units over the CTA pairs x (k-blocks per unit + a fixed
// epilogue cost of ~8 k-blocks).
int gemm_split_k_heuristic(int M, int N, int K, int bn) {
  const int groups = num_sms() / gemm_ncta();
  const int bm = 128 * gemm_ncta();
  const int tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
  const int kb = (K + 63) / 64;
  int best = 1;
  long best_cost = -1;
  for (int s = 1; s <= 16; ++s) {
    if (s > 1 && kb / s < 4) break;
    const long waves = ((long)tiles * s + groups - 1) / groups;
    const long cost = waves * ((kb + s - 1) / s + 8);
    if (best_cost < 0 || cost < best_cost) {
      best = s;
      best_cost = cost;
    }
  }
  return best;
}

int gemm_launch(const GemmDesc& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) {
    set_error("gemm: bad shape M=%d N=%d K=%d", g.M, g.N, g.K);
    return MODCR_ERR_SHAPE;
  }
  const bool has_aux = g.aux != nullptr;
  const bool has_in = g.epi_in != nullptr && g.in_op != 0;
  if (g.c_fp32) {
    if (has_aux) {
      set_error("gemm: fp32 output does not support an aux output");
      return MODCR_ERR_UNSUPPORTED;
    }
    if (has_in) {  // fp32 residual stream: C(fp32) = acc + bias + in(fp32)   [in_op 3: + LayerNorm(in)]
      if ((g.in_op != 1 && g.in_op != 3) || g.a_mn || g.b_mn || g.accumulate || g.split_k > 1) {
        set_error("gemm: fp32 epilogue input supports only in_op=1 on the NT layout without split-K");
        return MODCR_ERR_UNSUPPORTED;
      }
      if (g.K <= 1024 && gemm_wide_enabled()) return launch<256, false, false, false, true, true, true>(g, stream);
      return launch<256, false, false, false, true, true>(g, stream);
    }
    if (g.split_k > 1 && !g.accumulate) {
      set_error("gemm: split-K requires accumulate=1 (output must be pre-initialised)");
      return MODCR_ERR_UNSUPPORTED;
    }
    if (!g.a_mn && !g.b_mn) return launch<256, false, false, false, false, true>(g, stream);
    if (!g.a_mn && g.b_mn) return launch<256, false, true, false, false, true>(g, stream);
    if (g.a_mn && g.b_mn) return launch<256, true, true, false, false, true>(g, stream);
    set_error("gemm: operand major combination (A MN-major, B K-major) not instantiated");
    return MODCR_ERR_UNSUPPORTED;
  }
  if (g.accumulate || g.split_k > 1) {
    set_error("gemm: accumulate / split-K need an fp32 output");
    return MODCR_ERR_UNSUPPORTED;
  }
  if (g.a_mn) {
    set_error("gemm: A MN-major is only instantiated for fp32 (wgrad) output");
    return MODCR_ERR_UNSUPPORTED;
  }
  if (!g.b_mn) {
    if (has_aux && has_in) {
      set_error("gemm: aux output and epilogue input together are not instantiated");
      return MODCR_ERR_UNSUPPORTED;
    }
    if (has_aux) return launch<256, false, false, true, false, false>(g, stream);
    if (has_in) return launch<256, false, false, false, true, false>(g, stream);
    return launch<256, false, false, false, false, false>(g, stream);
  } else {
    if (has_aux) {
      set_error("gemm: aux output with B MN-major is not instantiated");
      return MODCR_ERR_UNSUPPORTED;
    }
    if (has_in) return launch<256, false, true, false, true, false>(g, stream);
    return launch<256, false, true, false, false, false>(g, stream);
  }
}

}  // namespace modcr
