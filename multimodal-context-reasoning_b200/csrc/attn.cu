// Fused prefix self-attention for sm_100a (tcgen05 / TMEM / TMA), forward and backward.
//
// Replaces CaptionBertSelfAttention.forward (reference modeling/modeling_bert.py:34-75 and
// modeling/modeling_vcr_chunkalign_v10.py:55-107) after the Q/K/V projections:
//
//     probs = softmax(Q K^T / sqrt(d) + mask)     K/V = [prefix rows ; sequence rows]
//     ctx   = probs V
//
// * The learned prefix K/V rows ([N,P,2H], projected from `history_state`) and the sequence
//   K/V rows (inside the fused [N,S,3H] QKV buffer) are TMA-loaded into ONE shared-memory
//   tile ([prefix padded to 16 rows ; sequence]); the concatenation is never materialised
//   in HBM.
// * One CTA per (q-tile of 128 rows, head, sequence).  S = QK^T is one tcgen05.mma group
//   (M=128, N=Kp<=512 in slices of <=256, K=64) into TMEM; 128 softmax threads own one row
//   each (TMEM lane == row): pass 1 max, pass 2 exp/sum and bf16 P into 128B-swizzled shared
//   memory, optional pass 3 stores the normalised text->region probabilities that feed the
//   alignment loss (v10:982-987).  O = P V is a second tcgen05.mma group whose accumulator
//   re-uses TMEM columns [0,64).
// * The mask zoo of the chunk-aware encoder (v10:171-232) is evaluated in
//   registers from compact inputs: an additive per-key bias vector, a per-row chunk bit mask
//   and the mode; the additive -10000 semantics of the reference are reproduced exactly.
// * Backward: one CTA per (head, sequence), K/V resident in shared memory, outer loop over
//   128-key blocks, inner loop over 128-row q tiles (Q / dO streamed through a 2-stage TMA
//   ring).  Per step: S and dP by tcgen05 into TMEM, P = exp(S - lse) and
//   dS = P * (dP + dAlign - delta) * scale by the row threads into shared memory, then
//   dV += P^T dO, dK += dS^T Q (accumulated in TMEM across q tiles) and dQ = dS K (flushed to
//   an fp32 partial when there is more than one key block).
#include "attn.h"

#include <math.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tma_host.h"

namespace modcr {

typedef __nv_bfloat16 bf16;
constexpr float NEG_BIAS = -10000.0f;

// Developer timeline of attn_bwd_kernel (build with -DMODCR_ATTN_TRACE, `make trace`): the first TRACE_CTAS CTAs record
// clock64() at fixed points of the row-thread / MMA-thread programs; read back with modcr_debug_read_attn_trace().
#ifdef MODCR_ATTN_TRACE
constexpr int TRACE_CTAS = 32, TRACE_SLOTS = 128;
__device__ long long g_attn_trace[TRACE_CTAS * TRACE_SLOTS];
// CTAs [0,16) (first wave) and [600,616) (a later wave) are recorded
#define TR(slot)                                                                                   \
  do {                                                                                             \
    int cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                                \
    if (cta_ >= 600) cta_ -= 584;                                                                  \
    if (cta_ < TRACE_CTAS && (slot) < TRACE_SLOTS) g_attn_trace[cta_ * TRACE_SLOTS + (slot)] = clock64(); \
  } while (0)
#else
#define TR(slot) do { } while (0)
#endif

struct AttnParams {
  int N, S, P, Ppad, T, R, heads, H;
  int Kp;      // key rows seen by the MMAs (multiple of 16): Ppad + round16(S); Ppad = round16(P)
  int Krows;   // key rows resident in shared memory: Ppad + 64 * ceil(S / 64)
  int nqt;     // 128-row q tiles
  int nkb;     // 128-key blocks (backward)
  int mode;
  float scale;
  const float* kbias;
  const uint32_t* cbits;
  int cwords;
  float* lse;
  float* probs_img;
  const bf16* ctx;
  const bf16* dctx;
  const float* dalign;
  bf16* dqkv;
  bf16* dpkv;
  float* dq_ws;
  // training-mode dropout of the attention probabilities (modeling_bert.py:62, v10:94): keep(seed, site, element)
  const uint32_t* drop_seed;   // device pointer to the step's seed; nullptr / drop_thr == 0: no dropout
  uint32_t drop_site, drop_thr;
  float drop_inv;              // 1 / (1 - p)
  int kid_split, kid_sub;      // logical key id of shared-memory key row kk: kk < kid_split ? S + kk : kk - kid_sub
  const int64_t* row_sel;      // [N,T] or nullptr: text rows whose text->region probabilities are wanted (align_pos)
};

// logical key id of shared-memory key row kk (same in every kernel, whatever its row order): sequence key j -> j,
// prefix key j -> S + j.  [prefix(pad) ; sequence] rows: split = sub = Ppad; [sequence ; prefix] rows: split = sub = 0.
DEVINL uint32_t drop_key_id(const AttnParams& p, int kk) {
  return kk < p.kid_split ? (uint32_t)(p.S + kk) : (uint32_t)(kk - p.kid_sub);
}

template <int W>
DEVINL void tmem_ld_cols(uint32_t taddr, float* v) {
  if (W == 32)
    tmem_ld_32x32(taddr, reinterpret_cast<uint32_t*>(v));
  else
    tmem_ld_32x16(taddr, reinterpret_cast<uint32_t*>(v));
}

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float NEG2 = NEG_BIAS * LOG2E;  // the additive -10000 in the log2 domain

// Additive-mask semantics of SURVEY.md A.2, evaluated in the log2 domain:
//   s2 = (q.k * scale + bias) * log2(e) = fma(acc, scale*log2e, bias*log2e)
// `b2` is the staged key bias (already times log2e; -inf marks a key row that does not exist:
// prefix padding, rows >= S, MMA padding), so exp2(s2 - max) of a missing key is exactly 0 and an
// additive -10000 underflows to exactly 0 like in the fp32 reference.
template <int MODE>
DEVINL float masked_s2(float acc, float c1, float b2, int kk, int q, bool q_txt, int T, uint32_t w) {
  if (MODE == 1) {
    if (kk < T) b2 = q_txt ? (((w >> (kk & 31)) & 1u) ? 0.f : NEG2) : NEG2;
  } else if (MODE == 2) {
    // branch-free: every alternative is a value select (the if / else-if form compiled to real branches inside the
    // unrolled column loops and made the layers 9-11 kernels branch-resolution bound, ncu: profiles/)
    const float bit = ((w >> (kk & 31)) & 1u) ? 0.f : NEG2;
    const float tval = q_txt ? bit : NEG2;                       // text key: own chunk (text row) / nothing (image row)
    const float diag = (kk == q) ? 0.f : NEG2;                   // image key, image row: identity only
    const float iv2 = (b2 == -INFINITY) ? b2 : diag;
    const float ival = q_txt ? b2 : iv2;                         // image key, text row: the key bias
    b2 = (kk < T) ? tval : ival;
  }
  return fmaf(acc, c1, b2);
}

// stage the per-key bias (times log2e) of sequence n into shared memory, indexed by smem key row
DEVINL void stage_bias(float* sBias, int count, const AttnParams& p, int n, int tid, int nthreads) {
  for (int kk = tid; kk < count; kk += nthreads) {
    float b = -INFINITY;
    if (kk < p.P)
      b = __ldg(p.kbias + (size_t)n * (p.P + p.S) + kk) * LOG2E;
    else if (kk >= p.Ppad && kk - p.Ppad < p.S)
      b = __ldg(p.kbias + (size_t)n * (p.P + p.S) + p.P + (kk - p.Ppad)) * LOG2E;
    sBias[kk] = b;
  }
}

// TMA-load the K and V tiles of (n, head): [prefix rows padded to Ppad ; sequence rows]
DEVINL void load_kv(uint8_t* sK, uint8_t* sV, const CUtensorMap* tmQKV, const CUtensorMap* tmPKV, uint64_t* barK,
                    uint64_t* barV, const AttnParams& p, int n, int h) {
  const int nbox = (p.S + 63) >> 6;
  const int npre = p.Ppad >> 4;
  const uint32_t bytes = (uint32_t)(npre * 2048 + nbox * 8192);
  mbar_expect_tx(barK, bytes);
  for (int i = 0; i < npre; ++i) tma_load_3d(sK + i * 2048, tmPKV, barK, h * 64, i * 16, n);
  for (int i = 0; i < nbox; ++i) tma_load_3d(sK + p.Ppad * 128 + i * 8192, tmQKV, barK, p.H + h * 64, i * 64, n);
  mbar_expect_tx(barV, bytes);
  for (int i = 0; i < npre; ++i) tma_load_3d(sV + i * 2048, tmPKV, barV, p.H + h * 64, i * 16, n);
  for (int i = 0; i < nbox; ++i)
    tma_load_3d(sV + p.Ppad * 128 + i * 8192, tmQKV, barV, 2 * p.H + h * 64, i * 64, n);
}

// ============================================================================ forward
struct FwdLayout {
  int off_k, off_v, off_bias, off_bars;  // byte offsets from the 1024-aligned base (Q tile at 0)
  int off_p[8];                          // P blocks (16 KB each); block 0 overlays Q, block 1 overlays K when it fits
};

// `drow` = drop_hash(seed, site, (n * heads + h) * S + q): the row part of the dropout mask hash
template <int MODE, int W, int PASS>
DEVINL void fwd_chunk(const AttnParams& p, float c1, uint32_t taddr, const float* sBias, uint8_t* blk, int c0, int row,
                      int q, bool q_txt, uint32_t w, float& m, float& sum, float inv, float* pimg_row, uint32_t drow = 0,
                      uint32_t dseed = 0) {
  float v[W];
  tmem_ld_cols<W>(taddr + c0, v);
  float bb[W];
#pragma unroll
  for (int j = 0; j < W / 4; ++j) {
    const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j);
    bb[4 * j] = b4.x; bb[4 * j + 1] = b4.y; bb[4 * j + 2] = b4.z; bb[4 * j + 3] = b4.w;
  }
  tmem_ld_wait();
  if (PASS == 1) {
#pragma unroll
    for (int j = 0; j < W; ++j) m = fmaxf(m, masked_s2<MODE>(v[j], c1, bb[j], c0 + j, q, q_txt, p.T, w));
  } else if (PASS == 2) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const float e = fast_ex2(masked_s2<MODE>(v[j], c1, bb[j], c0 + j, q, q_txt, p.T, w) - m);
      sum += e;
      v[j] = e;
    }
    if (p.drop_thr != 0) {   // uniform: P V uses the dropped probabilities, the row sum the undropped ones
#pragma unroll
      for (int j = 0; j < W; ++j)
        v[j] = drop_keep_row(drow, drop_key_id(p, c0 + j), p.drop_thr) ? v[j] * p.drop_inv : 0.f;
    }
    const int piece0 = (c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < W / 8; ++j) {
      uint4 pk;
      pk.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
      pk.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
      pk.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
      pk.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
      *reinterpret_cast<uint4*>(blk + sw128_offset(row, piece0 + j)) = pk;
    }
  } else {
    if (pimg_row != nullptr) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int kk = c0 + j;
        if (kk >= p.T && kk < p.S) {
          float pr = fast_ex2(masked_s2<MODE>(v[j], c1, bb[j], kk, q, q_txt, p.T, w) - m) * inv;
          if (p.drop_thr != 0) pr = drop_keep_row(drow, drop_key_id(p, kk), p.drop_thr) ? pr * p.drop_inv : 0.f;
          pimg_row[kk - p.T] = pr;
        }
      }
    }
  }
}

template <int TMEM_COLS, int MODE>
__global__ void __launch_bounds__(160, (TMEM_COLS == 256 ? 2 : 1))
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPKV,
                const __grid_constant__ CUtensorMap tmCTX, const AttnParams p, const FwdLayout lay) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nblk = (p.Kp + 63) >> 6;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + lay.off_k;
  uint8_t* sV = smem + lay.off_v;
  float* sBias = reinterpret_cast<float*>(smem + lay.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const float c1 = p.scale * LOG2E;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      if (p.P > 0) tma_prefetch_desc(&tmPKV);
      tma_prefetch_desc(&tmCTX);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_slot);
  } else {
    stage_bias(sBias, nblk * 64, p, n, tid, 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: MMA operands live in uniform registers

  if (warp == 4) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA + MMA issue
      const int nbox = (p.S + 63) >> 6;
      const int npre = p.Ppad >> 4;
      const uint32_t kv_bytes = (uint32_t)(npre * 2048 + nbox * 8192);
      mbar_expect_tx(bar_qk, 16384u + kv_bytes);
      tma_load_3d(sQ, &tmQKV, bar_qk, h * 64, qt * 128, n);
      tma_load_3d(sQ + 8192, &tmQKV, bar_qk, h * 64, qt * 128 + 64, n);
      for (int i = 0; i < npre; ++i) tma_load_3d(sK + i * 2048, &tmPKV, bar_qk, h * 64, i * 16, n);
      for (int i = 0; i < nbox; ++i)
        tma_load_3d(sK + p.Ppad * 128 + i * 8192, &tmQKV, bar_qk, p.H + h * 64, i * 64, n);
      mbar_expect_tx(bar_v, kv_bytes);
      for (int i = 0; i < npre; ++i) tma_load_3d(sV + i * 2048, &tmPKV, bar_v, p.H + h * 64, i * 16, n);
      for (int i = 0; i < nbox; ++i)
        tma_load_3d(sV + p.Ppad * 128 + i * 8192, &tmQKV, bar_v, 2 * p.H + h * 64, i * 64, n);

      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), s_addr = smem_u32(smem);
      for (int n0 = 0; n0 < p.Kp; n0 += 256) {
        const int nn = min(256, p.Kp - n0);
        const uint32_t idesc = make_idesc_bf16(128, nn, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + n0, make_smem_desc(q_addr + k * 32, 0, 1024),
                    make_smem_desc(k_addr + n0 * 128 + k * 32, 0, 1024), idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const int ksteps = p.Kp >> 4;
      for (int ks = 0; ks < ksteps; ++ks)
        umma_bf16(tmem_base, make_smem_desc(s_addr + lay.off_p[ks >> 2] + (ks & 3) * 32, 0, 1024),
                  make_smem_desc(v_addr + ks * 2048, 8192, 1024), idesc_pv, ks > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else {
    // ------------------------------------------------------------------ softmax rows
    const int row = tid;
    const int q = qt * 128 + row;
    const bool alive = q < p.S;
    const bool warp_alive = (qt * 128 + warp * 32) < p.S;
    const bool q_txt = q < p.T;
    const uint32_t dseed = p.drop_thr != 0 ? __ldg(p.drop_seed) : 0u;
    const uint32_t drow = p.drop_thr != 0 ? drop_hash(dseed, p.drop_site, (uint32_t)((n * p.heads + h) * p.S + q)) : 0u;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t* crow = (MODE != 0 && alive && q_txt) ? p.cbits + ((size_t)n * p.T + q) * p.cwords : nullptr;
    float* pimg_row = nullptr;
    if (p.probs_img != nullptr && alive && q_txt && (p.row_sel == nullptr || __ldg(p.row_sel + (size_t)n * p.T + q) != 0))
      pimg_row = p.probs_img + (((size_t)n * p.heads + h) * p.T + q) * p.R;

    mbar_wait(bar_s, 0);   // S complete: the Q and K tiles are dead, P blocks 0/1 may overlay them
    tc_fence_after();
    float m = -INFINITY, sum = 0.f, inv = 0.f;
    if (warp_alive) {
      const int full = p.Kp & ~31;
      for (int c0 = 0; c0 < full; c0 += 32) {
        const uint32_t w = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
        fwd_chunk<MODE, 32, 1>(p, c1, taddr, sBias, nullptr, c0, row, q, q_txt, w, m, sum, inv, pimg_row, drow, dseed);
      }
      if (full < p.Kp) {
        const uint32_t w = (crow != nullptr && (full >> 5) < p.cwords) ? __ldg(crow + (full >> 5)) : 0u;
        fwd_chunk<MODE, 16, 1>(p, c1, taddr, sBias, nullptr, full, row, q, q_txt, w, m, sum, inv, pimg_row, drow, dseed);
      }
      for (int c0 = 0; c0 < full; c0 += 32) {
        const uint32_t w = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
        fwd_chunk<MODE, 32, 2>(p, c1, taddr, sBias, smem + lay.off_p[c0 >> 6], c0, row, q, q_txt, w, m, sum, inv,
                               pimg_row, drow, dseed);
      }
      if (full < p.Kp) {
        const uint32_t w = (crow != nullptr && (full >> 5) < p.cwords) ? __ldg(crow + (full >> 5)) : 0u;
        fwd_chunk<MODE, 16, 2>(p, c1, taddr, sBias, smem + lay.off_p[full >> 6], full, row, q, q_txt, w, m, sum, inv,
                               pimg_row, drow, dseed);
      }
      inv = sum > 0.f ? 1.0f / sum : 0.f;
      if (p.probs_img != nullptr) {  // uniform over the CTA
        const int first = (p.T >> 5) << 5;
        for (int c0 = first; c0 < full; c0 += 32) {
          const uint32_t w = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
          fwd_chunk<MODE, 32, 3>(p, c1, taddr, sBias, nullptr, c0, row, q, q_txt, w, m, sum, inv, pimg_row, drow, dseed);
        }
        if (full < p.Kp && full + 16 > first) {
          const uint32_t w = (crow != nullptr && (full >> 5) < p.cwords) ? __ldg(crow + (full >> 5)) : 0u;
          fwd_chunk<MODE, 16, 3>(p, c1, taddr, sBias, nullptr, full, row, q, q_txt, w, m, sum, inv, pimg_row, drow, dseed);
        }
      }
      if (alive) p.lse[((size_t)n * p.heads + h) * p.S + q] = (m + log2f(sum)) * LN2;
    }
    fence_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);   // O complete: every P block is dead, block 0 (== the Q tile) becomes the staging tile
    tc_fence_after();
    if (warp_alive) {
      float o[64];
      tmem_ld_32x32(taddr, reinterpret_cast<uint32_t*>(o));
      tmem_ld_32x32(taddr + 32, reinterpret_cast<uint32_t*>(o) + 32);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 pk;
        pk.x = pack_bf16(o[8 * j + 0] * inv, o[8 * j + 1] * inv);
        pk.y = pack_bf16(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
        pk.z = pack_bf16(o[8 * j + 4] * inv, o[8 * j + 5] * inv);
        pk.w = pack_bf16(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
        *reinterpret_cast<uint4*>(sQ + sw128_offset(row, j)) = pk;
      }
    }
    fence_async_smem();
    named_bar_sync(1, 128);
    if (tid == 0) {
      tma_store_3d(&tmCTX, sQ, h * 64, qt * 128, n);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ============================================================================ forward, persistent
// For S <= 192 and P+S <= 192 (the PMR / VCR-base shapes).  One CTA per SM loops over (sequence, head) items; both
// 128-row q tiles of an item are in flight at once on two softmax groups that share the item's K/V tiles:
//   warps 0-3  group A: q rows [0,128)          warps 4-5  group B: q rows [128,192) (<= 64 rows, PV still M=128)
//   warp 6     tcgen05.mma issuer + TMEM owner  warp 7     TMA producer (2-stage ring of Q_A | Q_B | K | V)
// TMEM: group g owns columns [256 g, 256 g + 256): S in [0,192), O in [192,256), so S of the next item is issued the
// moment the group has finished reading S of the current one, ahead of P V.  The row threads run
//   pass 1 (max) of item i  ->  epilogue of item i-1 (O / sum -> bf16 -> TMA store)  ->  pass 2 (exp, P) of item i
// which hides the P V latency of item i-1 behind pass 1 of item i.  P of group A overlays the dead Q_A | Q_B | K
// tiles of its stage; group B (64 rows) has its own half-height P blocks whose upper halves are only ever read as
// don't-care rows by the M=128 MMA.
// Key rows in shared memory are [sequence rows ; prefix rows] with NO padding in between (exact-height TMA boxes for
// the sequence tail and the prefix; the swizzle is a function of the shared-memory address, so a box may start on
// any 128-byte row), which keeps P + S = 188 of the PMR shape inside 192 MMA columns; rows [S+P, Kp) are zeroed once.
struct PersistLayout {
  int stage_bytes, off_qb, off_k, off_v;   // per stage: Q_A at 0
  int off_pb, off_oa, off_bias, off_bars;  // P blocks of group B (+ its O staging at +24 KB), O staging of group A
};

template <int MODE>
__global__ void __launch_bounds__(256, 1)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTail,
                        const __grid_constant__ CUtensorMap tmPKV, const __grid_constant__ CUtensorMap tmCTXA,
                        const __grid_constant__ CUtensorMap tmCTXB, const AttnParams p, const PersistLayout lay) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sPB = smem + lay.off_pb;
  uint8_t* sOA = smem + lay.off_oa;
  uint8_t* sOB = sPB + 24576;
  float* sBiasAll = reinterpret_cast<float*>(smem + lay.off_bias);   // [2 groups][2 item parities][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
  uint64_t* qk_full = bars + 0;      // [2] stage: Q_A, Q_B, K landed
  uint64_t* v_full = bars + 2;       // [2] stage: V landed
  uint64_t* stage_free = bars + 4;   // [2] every MMA reading the stage has retired
  uint64_t* s_full = bars + 6;       // [2] group: S = Q K^T complete
  uint64_t* p_ready = bars + 8;      // [2] group: P written (and S consumed)
  uint64_t* o_full = bars + 10;      // [2] group: O = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const float c1 = p.scale * LOG2E;
  const int nblk = (p.Kp + 63) >> 6;
  const bool two = p.nqt > 1;        // group B has rows
  const int total = p.N * p.heads;
  const int n_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 7) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      if (p.P > 0) tma_prefetch_desc(&tmPKV);
      tma_prefetch_desc(&tmCTXA);
      tma_prefetch_desc(&tmCTXB);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&qk_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&stage_free[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&o_full[i], 1);
      }
      mbar_init(&p_ready[0], 128);
      mbar_init(&p_ready[1], 64);
      fence_barrier_init();
    }
  } else if (warp == 6) {
    tmem_alloc<512>(tmem_slot);
  } else {
    // key rows [S+P, Kp) are never written by TMA: clear them once in both stages (K: finite scores under the -inf
    // bias; V: 0 x 0 instead of 0 x garbage)
    const int pad0 = (p.S + p.P) * 128, pad1 = p.Kp * 128;
    for (int st = 0; st < 2; ++st) {
      uint8_t* sK = smem + st * lay.stage_bytes + lay.off_k;
      uint8_t* sV = smem + st * lay.stage_bytes + lay.off_v;
      for (int b = pad0 + tid * 16; b < pad1; b += 192 * 16) {
        *reinterpret_cast<uint4*>(sK + b) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sV + b) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: MMA operands live in uniform registers

  if (warp == 7) {
    // ---------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const int nfull = p.S >> 6;          // full 64-row boxes of the sequence
      const int tail = p.S & 63;           // rows of the exact-height tail box
      const uint32_t kv_bytes = (uint32_t)((p.S + p.P) * 128);
      const uint32_t q_bytes = two ? 24576u : 16384u;   // group B: rows [128,192) only
      for (int i = 0; i < n_items; ++i) {
        const int w = blockIdx.x + i * gridDim.x;
        const int n = w / p.heads, h = w - n * p.heads;
        const int st = i & 1;
        uint8_t* sQ = smem + st * lay.stage_bytes;
        uint8_t* sK = sQ + lay.off_k;
        uint8_t* sV = sQ + lay.off_v;
        if (i >= 2) mbar_wait(&stage_free[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qk_full[st], q_bytes + kv_bytes);
        tma_load_3d(sQ, &tmQKV, &qk_full[st], h * 64, 0, n);
        tma_load_3d(sQ + 8192, &tmQKV, &qk_full[st], h * 64, 64, n);
        if (two) tma_load_3d(sQ + lay.off_qb, &tmQKV, &qk_full[st], h * 64, 128, n);
        for (int j = 0; j < nfull; ++j) tma_load_3d(sK + j * 8192, &tmQKV, &qk_full[st], p.H + h * 64, j * 64, n);
        if (tail) tma_load_3d(sK + nfull * 8192, &tmTail, &qk_full[st], p.H + h * 64, nfull * 64, n);
        if (p.P > 0) tma_load_3d(sK + p.S * 128, &tmPKV, &qk_full[st], h * 64, 0, n);
        mbar_expect_tx(&v_full[st], kv_bytes);
        for (int j = 0; j < nfull; ++j) tma_load_3d(sV + j * 8192, &tmQKV, &v_full[st], 2 * p.H + h * 64, j * 64, n);
        if (tail) tma_load_3d(sV + nfull * 8192, &tmTail, &v_full[st], 2 * p.H + h * 64, nfull * 64, n);
        if (p.P > 0) tma_load_3d(sV + p.S * 128, &tmPKV, &v_full[st], p.H + h * 64, 0, n);
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, p.Kp, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const int ksteps = p.Kp >> 4;
      auto issue_s = [&](int i, int g) {   // S_g(i) = Q_g K^T
        const int st = i & 1;
        const uint32_t q_addr = smem_u32(smem + st * lay.stage_bytes + (g ? lay.off_qb : 0));
        const uint32_t k_addr = smem_u32(smem + st * lay.stage_bytes + lay.off_k);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + g * 256, make_smem_desc(q_addr + k * 32, 0, 1024),
                    make_smem_desc(k_addr + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int i, int g) {  // O_g(i) = P_g V
        const int st = i & 1;
        const uint32_t base = smem_u32(smem + st * lay.stage_bytes);
        const uint32_t v_addr = base + lay.off_v;
        const uint32_t pa[3] = {base, base + (uint32_t)lay.off_qb, base + (uint32_t)lay.off_k};
        const uint32_t pb = smem_u32(sPB);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t blk = g ? pb + (ks >> 2) * 8192 : pa[ks >> 2];
          umma_bf16(tmem_base + g * 256 + 192, make_smem_desc(blk + (ks & 3) * 32, 0, 1024),
                    make_smem_desc(v_addr + ks * 2048, 8192, 1024), idesc_pv, ks > 0 ? 1u : 0u);
        }
        umma_commit(&o_full[g]);
      };
      if (n_items > 0) {
        mbar_wait(&qk_full[0], 0);
        tc_fence_after();
        issue_s(0, 0);
        if (two) issue_s(0, 1);
      }
      for (int i = 0; i < n_items; ++i) {
        const int st = i & 1;
        const bool more = i + 1 < n_items;
        for (int g = 0; g < (two ? 2 : 1); ++g) {
          mbar_wait(&p_ready[g], i & 1);      // P_g(i) written, S_g columns consumed
          tc_fence_after();
          if (more) {
            if (g == 0) {
              mbar_wait(&qk_full[st ^ 1], ((i + 1) >> 1) & 1);
              tc_fence_after();
            }
            issue_s(i + 1, g);
          }
          if (g == 0) {
            mbar_wait(&v_full[st], (i >> 1) & 1);
            tc_fence_after();
          }
          issue_pv(i, g);
        }
        umma_commit(&stage_free[st]);
      }
    }
  } else {
    // ---------------------------------------------------------------------- softmax rows
    const int g = warp >> 2;                 // 0: rows [0,128), 1: rows [128,192)
    const int gthreads = g ? 64 : 128;
    const int gt = tid - g * 128;            // thread index inside the group == row inside the q tile
    const int row = gt;
    const int q = g * 128 + row;
    const bool alive = q < p.S;
    const bool warp_alive = (g * 128 + (warp & 3) * 32) < p.S;
    const bool q_txt = q < p.T;
    const uint32_t dseed = p.drop_thr != 0 ? __ldg(p.drop_seed) : 0u;
    const bool active = g == 0 || two;       // group B without rows only idles
    const uint32_t bar_id = 1 + g;
    const bool issuer = gt == 0;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + g * 256;
    uint8_t* sO = g ? sOB : sOA;
    const CUtensorMap* tmO = g ? &tmCTXB : &tmCTXA;
    const int full = p.Kp & ~31;

    float inv_prev = 0.f;
    int n_prev = 0, h_prev = 0;
    auto epilogue = [&](int j) {   // O of item j -> ctx
      mbar_wait(&o_full[g], j & 1);
      tc_fence_after();
      float o[64];
      if (warp_alive) {
        tmem_ld_32x32(taddr + 192, reinterpret_cast<uint32_t*>(o));
        tmem_ld_32x32(taddr + 224, reinterpret_cast<uint32_t*>(o) + 32);
        tmem_ld_wait();
      }
      if (issuer) tma_store_wait_read<0>();   // the previous store has finished reading the staging tile
      named_bar_sync(bar_id, gthreads);
      if (warp_alive) {
        const uint32_t so = smem_u32(sO);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          sts128(so + sw128_offset(row, jj), pack_bf16(o[8 * jj + 0] * inv_prev, o[8 * jj + 1] * inv_prev),
                 pack_bf16(o[8 * jj + 2] * inv_prev, o[8 * jj + 3] * inv_prev),
                 pack_bf16(o[8 * jj + 4] * inv_prev, o[8 * jj + 5] * inv_prev),
                 pack_bf16(o[8 * jj + 6] * inv_prev, o[8 * jj + 7] * inv_prev));
      }
      fence_async_smem();
      tc_fence_before();
      named_bar_sync(bar_id, gthreads);
      if (issuer) {
        tma_store_3d(tmO, sO, h_prev * 64, g * 128, n_prev);
        tma_store_commit();
      }
    };

    if (active) {
      for (int i = 0; i < n_items; ++i) {
        const int w = blockIdx.x + i * gridDim.x;
        const int n = w / p.heads, h = w - n * p.heads;
        const int st = i & 1;
        uint8_t* stage = smem + st * lay.stage_bytes;
        const uint32_t drow = p.drop_thr != 0 ? drop_hash(dseed, p.drop_site, (uint32_t)((n * p.heads + h) * p.S + q)) : 0u;
        // ---- per-key bias of this sequence, double-buffered by item parity: a slower thread of the group may still
        // be reading the previous item's values, but everyone is past the ones of two items ago (barrier below)
        float* sBias = sBiasAll + (g * 2 + (i & 1)) * 256;
        for (int kk = gt; kk < nblk * 64; kk += gthreads) {   // smem key row kk: sequence rows first, then the prefix
          float b = -INFINITY;
          if (kk < p.S)
            b = __ldg(p.kbias + (size_t)n * (p.P + p.S) + p.P + kk) * LOG2E;
          else if (kk < p.S + p.P)
            b = __ldg(p.kbias + (size_t)n * (p.P + p.S) + (kk - p.S)) * LOG2E;
          sBias[kk] = b;
        }
        named_bar_sync(bar_id, gthreads);
        const uint32_t* crow = (MODE != 0 && alive && q_txt) ? p.cbits + ((size_t)n * p.T + q) * p.cwords : nullptr;
        float* pimg_row = nullptr;
        if (p.probs_img != nullptr && alive && q_txt)
          pimg_row = p.probs_img + (((size_t)n * p.heads + h) * p.T + q) * p.R;
        // rows of this warp whose probabilities are wanted (row_sel = align_pos: the alignment loss reads no other row)
        const bool row_wanted = pimg_row != nullptr && (p.row_sel == nullptr || __ldg(p.row_sel + (size_t)n * p.T + q) != 0);
        const uint32_t want_mask = p.probs_img != nullptr ? __ballot_sync(0xffffffffu, row_wanted) : 0u;

        // ---- pass 1: row maximum
        mbar_wait(&s_full[g], i & 1);
        tc_fence_after();
        float m = -INFINITY, sum = 0.f, inv = 0.f;
        if (warp_alive) {
          for (int c0 = 0; c0 < full; c0 += 32) {
            const uint32_t wd = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
            fwd_chunk<MODE, 32, 1>(p, c1, taddr, sBias, nullptr, c0, row, q, q_txt, wd, m, sum, inv, pimg_row, drow, dseed);
          }
          if (full < p.Kp) {
            const uint32_t wd = (crow != nullptr && (full >> 5) < p.cwords) ? __ldg(crow + (full >> 5)) : 0u;
            fwd_chunk<MODE, 16, 1>(p, c1, taddr, sBias, nullptr, full, row, q, q_txt, wd, m, sum, inv, pimg_row, drow, dseed);
          }
        }
        // ---- epilogue of the previous item (its P V ran under pass 1)
        if (i > 0) epilogue(i - 1);
        // ---- pass 2: exp, row sum, P (bf16, swizzled) for the P V MMA
        if (g == 0 && two) {   // P blocks 1, 2 overlay Q_B and K: group B's S = Q_B K^T must have retired too
          mbar_wait(&s_full[1], i & 1);
        }
        if (warp_alive) {
          for (int c0 = 0; c0 < full; c0 += 32) {
            const uint32_t wd = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
            const int b = c0 >> 6;
            uint8_t* blk = g ? sPB + b * 8192 : stage + (b == 0 ? 0 : (b == 1 ? lay.off_qb : lay.off_k));
            fwd_chunk<MODE, 32, 2>(p, c1, taddr, sBias, blk, c0, row, q, q_txt, wd, m, sum, inv, pimg_row, drow, dseed);
          }
          if (full < p.Kp) {
            const uint32_t wd = (crow != nullptr && (full >> 5) < p.cwords) ? __ldg(crow + (full >> 5)) : 0u;
            const int b = full >> 6;
            uint8_t* blk = g ? sPB + b * 8192 : stage + (b == 0 ? 0 : (b == 1 ? lay.off_qb : lay.off_k));
            fwd_chunk<MODE, 16, 2>(p, c1, taddr, sBias, blk, full, row, q, q_txt, wd, m, sum, inv, pimg_row, drow, dseed);
          }
          inv = sum > 0.f ? 1.0f / sum : 0.f;
          if (alive) p.lse[((size_t)n * p.heads + h) * p.S + q] = (m + log2f(sum)) * LN2;
        }
        if (p.probs_img != nullptr && p.row_sel != nullptr) {  // uniform over the CTA
          // Row-selected form (row_sel = align_pos, ~4 rows of a sequence): the few wanted rows write their own R
          // probabilities; the TMEM loads are warp-collective, the exp / store work is done by the wanted lanes only.
          if (warp_alive && want_mask != 0u) {
            const int first = (p.T >> 5) << 5;
            for (int c0 = first; c0 < p.Kp; c0 += 32) {
              const int wcols = min(32, p.Kp - c0);
              float v[32];
              if (wcols == 32) tmem_ld_32x32(taddr + c0, reinterpret_cast<uint32_t*>(v));
              else tmem_ld_32x16(taddr + c0, reinterpret_cast<uint32_t*>(v));
              tmem_ld_wait();
              if (row_wanted) {
                const uint32_t wd = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int kk = c0 + j;
                  if (j < wcols && kk >= p.T && kk < p.S) {
                    float pr = fast_ex2(masked_s2<MODE>(v[j], c1, sBias[kk], kk, q, q_txt, p.T, wd) - m) * inv;
                    if (p.drop_thr != 0) pr = drop_keep_row(drow, drop_key_id(p, kk), p.drop_thr) ? pr * p.drop_inv : 0.f;
                    pimg_row[kk - p.T] = pr;
                  }
                }
              }
            }
          }
        } else if (p.probs_img != nullptr) {  // uniform over the CTA
          // Normalised text->region probabilities of this head (v10:982 consumes their sum over heads and layers).
          // A thread owns a ROW, global memory wants a row's columns side by side: each warp transposes its
          // 32 x 32 block through its 4 KB slice of the group's O staging tile (XOR-swizzled, conflict-free) and
          // writes 128-byte row segments.
          if (issuer) tma_store_wait_read<0>();   // the previous item's ctx store has finished reading the tile
          named_bar_sync(bar_id, gthreads);
          if (warp_alive && want_mask != 0u) {
            const uint32_t tile = smem_u32(sO) + (uint32_t)(warp & 3) * 4096u;
            const int q0 = g * 128 + (warp & 3) * 32;          // first q row of this warp
            float* pbase = p.probs_img + (((size_t)n * p.heads + h) * p.T) * p.R;
            const int first = (p.T >> 5) << 5;
            for (int c0 = first; c0 < p.Kp; c0 += 32) {
              const int wcols = min(32, p.Kp - c0);             // 32, or 16 for the last chunk
              float v[32];
              if (wcols == 32) tmem_ld_32x32(taddr + c0, reinterpret_cast<uint32_t*>(v));
              else tmem_ld_32x16(taddr + c0, reinterpret_cast<uint32_t*>(v));
              tmem_ld_wait();
              const uint32_t wd = (crow != nullptr && (c0 >> 5) < p.cwords) ? __ldg(crow + (c0 >> 5)) : 0u;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float pr = 0.f;
                if (j < wcols && row_wanted) {
                  pr = fast_ex2(masked_s2<MODE>(v[j], c1, sBias[c0 + j], c0 + j, q, q_txt, p.T, wd) - m) * inv;
                  if (p.drop_thr != 0)   // the alignment head consumes the post-dropout probabilities (v10:982)
                    pr = drop_keep_row(drow, drop_key_id(p, c0 + j), p.drop_thr) ? pr * p.drop_inv : 0.f;
                }
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(tile + lane * 128 + ((j ^ lane) << 2)), "f"(pr) : "memory");
              }
              __syncwarp();
              const int col = c0 + lane;                        // this lane's key column
              const bool col_ok = col >= p.T && col < p.S;
              for (int r = 0; r < 32; ++r) {
                float pr;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pr) : "r"(tile + r * 128 + ((lane ^ r) << 2)));
                const int qr = q0 + r;
                if (col_ok && ((want_mask >> r) & 1u)) pbase[(size_t)qr * p.R + (col - p.T)] = pr;
              }
              __syncwarp();
            }
          }
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(&p_ready[g]);
        inv_prev = inv;
        n_prev = n;
        h_prev = h;
      }
      if (n_items > 0) epilogue(n_items - 1);
      if (issuer) tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 6) tmem_dealloc<512>(tmem_base);
}

// ============================================================================ backward
// Row threads: 8 warps.  Warp w owns TMEM lanes 32*(w&3) (= q rows of the tile) and the key-column half (w>>2)
// of every 128-key block, so two warps share each SM sub-partition's latency.
//
// Schedule (one step = one (key block, q tile) pair, steps run kb-major):
//   MMA thread :  S/dP(0) | for i: for each 64-key sub-block c: wait "S/dP(i,c) consumed" -> issue S/dP(i+1,c)
//                 | wait P/dS(i) -> issue dV,dK,dQ(i) -> commit "free(i)"
//   row threads:  for i: for c: wait S/dP(i,c) -> tcgen05.ld -> arrive "consumed" -> P, dS math into registers
//                 | wait free(i-1) -> read out whatever step i-1 completed (dQ, dK/dV) -> store P/dS(i) -> arrive
// Every thread's two 32-column chunks lie in different 64-key sub-blocks, and S/dP of a sub-block is its own MMA
// group with its own barrier, so the product for step i+1 is re-issued as soon as the registers hold step i's
// values: the exp / dS math of step i runs while the tensor core computes S/dP(i+1) and dV/dK/dQ(i-1).  With at most two q tiles (S <= 256) the dQ accumulators
// of both tiles stay resident in TMEM across key blocks (columns 256.. and 448..), so dQ is written once, as bf16,
// and the fp32 workspace round trip disappears; longer sequences flush dQ per step through the workspace.
// P and dS of one 32-column chunk from the S / dP accumulator values already in registers.  ONE instance per mask mode:
// the partial last key block (stale TMEM columns beyond the MMA width) and the alignment-loss gradient on the region
// columns are handled by the caller with two small warp-uniform loops over sv / dp before this runs, instead of four
// template variants of the whole body (the row-thread program has to stay inside the instruction cache: the variant
// code made every first visit of a branch cost thousands of cycles of instruction fetch, ncu `no_instructions` and the
// clock64() timeline in profiles/r02_attn_bwd_trace_before.log).
template <int MODE>
DEVINL void bwd_chunk_math(const AttnParams& p, float c1, const float* sv, const float* dp, const float* sBias, int kk0,
                           int q, bool q_txt, uint32_t w, float lse2, float delta_s, uint32_t* pp, uint32_t* dd,
                           uint32_t drow = 0) {
#pragma unroll
  for (int g4 = 0; g4 < 8; ++g4) {   // 4 columns at a time keeps the live register set small
    const float4 b4 = *reinterpret_cast<const float4*>(sBias + kk0 + 4 * g4);
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float pr[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * g4 + e;
      const int kk = kk0 + j;
      float g = dp[j];
      pr[e] = fast_ex2(masked_s2<MODE>(sv[j], c1, bb[e], kk, q, q_txt, p.T, w) - lse2);
      float pd = pr[e];
      if (p.drop_thr != 0) {   // uniform: dV uses D o P, dS uses D o dP (D = keep / (1 - p)); delta is unchanged
        const float dm = drop_keep_row(drow, drop_key_id(p, kk), p.drop_thr) ? p.drop_inv : 0.f;
        pd *= dm;
        g *= dm;
      }
      ds[e] = pr[e] * fmaf(g, p.scale, -delta_s);   // dS = P (dP - delta) scale
      pr[e] = pd;
    }
    pp[2 * g4] = pack_bf16(pr[0], pr[1]);
    pp[2 * g4 + 1] = pack_bf16(pr[2], pr[3]);
    dd[2 * g4] = pack_bf16(ds[0], ds[1]);
    dd[2 * g4 + 1] = pack_bf16(ds[2], ds[3]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPKV,
                const __grid_constant__ CUtensorMap tmDO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + p.Krows * 128;
  uint8_t* sQ = sV + p.Krows * 128;   // 2 stages x 16 KB
  uint8_t* sDO = sQ + 2 * 16384;      // 2 stages x 16 KB
  uint8_t* sP = sDO + 2 * 16384;      // 2 blocks of [128 x 64] bf16
  uint8_t* sDS = sP + 2 * 16384;
  float* sBias = reinterpret_cast<float*>(sDS + 2 * 16384);
  float* sDelta = sBias + p.nkb * 128;
  float* sLse = sDelta + p.nqt * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLse + p.nqt * 128);
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* qfull = bars + 2;    // [2]
  uint64_t* qempty = bars + 4;   // [2]
  uint64_t* bar_sdp = bars + 6;       // S/dP of a step complete
  uint64_t* bar_cons = bars + 8;      // ... and loaded into registers by every row thread
  uint64_t* bar_pds = bars + 10;
  uint64_t* bar_free = bars + 11;     // dV/dK/dQ MMAs of a step retired: P/dS smem free, dQ (and dK/dV) readable
  uint64_t* bar_dkv_free = bars + 12; // dK/dV accumulators of a key block have been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int h = blockIdx.x, n = blockIdx.y;
  const float c1 = p.scale * LOG2E;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ0 = 256, COL_DK = 320, COL_DV = 384, COL_DQ1 = 448;
  const bool dq_resident = p.nqt <= 2;   // uniform over the grid
  // at most two q tiles: the Q / dO ring IS the set of q tiles (stage = q tile), loaded once and reused by every key
  // block instead of being streamed again per key block
  const bool q_resident = p.nqt <= 2;
  const int total_steps = p.nkb * p.nqt;
  if (tid == 0) TR(0);
  if (tid == 8 * 32) TR(64);

  if (warp == 9) {
    // ---------------------------------------------------------------------- TMA producer
    // initialises the barriers and starts the K/V and first Q/dO loads BEFORE the block-wide sync, so the loads
    // run under the delta / lse prologue of the row threads
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      if (p.P > 0) tma_prefetch_desc(&tmPKV);
      tma_prefetch_desc(&tmDO);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&qfull[i], 1);
        mbar_init(&qempty[i], 1);
      }
      mbar_init(bar_sdp, 1);
      mbar_init(bar_cons, 256);
      mbar_init(bar_pds, 256);
      mbar_init(bar_free, 1);
      mbar_init(bar_dkv_free, 256);
      fence_barrier_init();
      load_kv(sK, sV, &tmQKV, &tmPKV, bar_k, bar_v, p, n, h);
      for (int step = 0; step < 2 && step < total_steps; ++step) {   // both ring stages are free at start
        const int qt = step % p.nqt;
        mbar_expect_tx(&qfull[step], 32768);
        tma_load_3d(sQ + step * 16384, &tmQKV, &qfull[step], h * 64, qt * 128, n);
        tma_load_3d(sQ + step * 16384 + 8192, &tmQKV, &qfull[step], h * 64, qt * 128 + 64, n);
        tma_load_3d(sDO + step * 16384, &tmDO, &qfull[step], h * 64, qt * 128, n);
        tma_load_3d(sDO + step * 16384 + 8192, &tmDO, &qfull[step], h * 64, qt * 128 + 64, n);
      }
    }
  } else if (warp == 8) {
    tmem_alloc<512>(tmem_slot);
  } else {
    stage_bias(sBias, p.nkb * 128, p, n, tid, 256);
    // delta[q] = sum_d dO[q,d] O[q,d] (+ sum_r P_img[q,r] dAlign[q,r]) ; lse[q] (log2 domain, +inf for dead rows).
    // (Computing the dO . O part from TMA-staged O tiles inside the step pipeline was measured SLOWER, 233 vs 229 us:
    // these strided reads already overlap the K/V loads, profiles/r02_attn_time_6*.log.)
    for (int r = tid; r < p.nqt * 128; r += 256) {
      float d = 0.f, l = INFINITY;
      if (r < p.S) {
        const uint4* po = reinterpret_cast<const uint4*>(p.ctx + ((size_t)n * p.S + r) * p.H + h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dctx + ((size_t)n * p.S + r) * p.H + h * 64);
        uint4 a[8], b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] = __ldg(po + j);
          b[j] = __ldg(pd + j);
        }
        l = __ldg(p.lse + ((size_t)n * p.heads + h) * p.S + r) * LOG2E;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t aw[4] = {a[j].x, a[j].y, a[j].z, a[j].w}, bw[4] = {b[j].x, b[j].y, b[j].z, b[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fa = unpack_bf16(aw[e]), fb = unpack_bf16(bw[e]);
            d = fmaf(fa.x, fb.x, d);
            d = fmaf(fa.y, fb.y, d);
          }
        }
        // alignment term: only rows that feed the alignment loss carry a gradient (row_sel = align_pos; without it every
        // text row is assumed to).  Two independent accumulators + 128-bit loads keep the loads in flight instead of a
        // 50-deep dependent chain per row (that chain was 15k cycles of prologue per CTA in the mode-2 layers).
        if (p.dalign != nullptr && p.probs_img != nullptr && r < p.T &&
            (p.row_sel == nullptr || __ldg(p.row_sel + (size_t)n * p.T + r) != 0)) {
          const float* pi = p.probs_img + (((size_t)n * p.heads + h) * p.T + r) * p.R;
          const float* da = p.dalign + ((size_t)n * p.T + r) * p.R;
          float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
          int j = 0;
          if ((p.R & 1) == 0) {   // even R: every row starts 8-byte aligned; 64-bit loads, 8 in flight per array
#pragma unroll 4
            for (; j + 4 <= p.R; j += 4) {
              const float2 x0 = __ldg(reinterpret_cast<const float2*>(pi + j)), y0 = __ldg(reinterpret_cast<const float2*>(da + j));
              const float2 x1 = __ldg(reinterpret_cast<const float2*>(pi + j + 2)), y1 = __ldg(reinterpret_cast<const float2*>(da + j + 2));
              d0 = fmaf(x0.x, y0.x, d0);
              d1 = fmaf(x0.y, y0.y, d1);
              d2 = fmaf(x1.x, y1.x, d2);
              d3 = fmaf(x1.y, y1.y, d3);
            }
          }
          for (; j < p.R; ++j) d0 = fmaf(__ldg(pi + j), __ldg(da + j), d0);
          d0 += d2;
          d1 += d3;
          d += d0 + d1;
        }
      }
      sDelta[r] = d;
      sLse[r] = l;
    }
    if (tid == 0) TR(1);
  }
  if (lane == 0) TR(100 + warp);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: MMA operands live in uniform registers
  if (tid == 0) TR(2);
  if (tid == 8 * 32) TR(65);

#ifdef MODCR_ATTN_TRACE
  if (warp == 9 && lane == 1) {   // observer: true completion times of the MMA groups
    for (int step = 0; step < total_steps && step < 4; ++step) {
      mbar_wait(bar_sdp, step & 1);
      TR(112 + step * 3 + 4);
      mbar_wait(bar_free, step & 1);
      TR(114 + step * 3 + 4);
    }
  }
#endif
  if (warp == 9) {
    if (lane == 0 && !q_resident) {
      for (int step = 2; step < total_steps; ++step) {
        const int qt = step % p.nqt;
        const int s = step & 1;
        const uint32_t ph = (step >> 1) & 1;
        mbar_wait(&qempty[s], ph ^ 1);
        mbar_expect_tx(&qfull[s], 32768);
        tma_load_3d(sQ + s * 16384, &tmQKV, &qfull[s], h * 64, qt * 128, n);
        tma_load_3d(sQ + s * 16384 + 8192, &tmQKV, &qfull[s], h * 64, qt * 128 + 64, n);
        tma_load_3d(sDO + s * 16384, &tmDO, &qfull[s], h * 64, qt * 128, n);
        tma_load_3d(sDO + s * 16384 + 8192, &tmDO, &qfull[s], h * 64, qt * 128 + 64, n);
      }
    }
  } else if (warp == 8) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      const uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);
      // Descriptors are built ONCE per operand and advanced by adding the byte offset >> 4 to the 64-bit value (the
      // address field is the low 14 bits and shared-memory addresses never carry out of it): the issuing thread is the
      // bottleneck of this kernel -- the MMAs are N = 64 and retire faster than a thread that rebuilds two descriptors
      // per instruction can issue them (clock64() timeline, profiles/r02_attn_bwd_trace_*.log).
      const uint64_t k_km = make_smem_desc(k_addr, 0, 1024), v_km = make_smem_desc(v_addr, 0, 1024);         // K-major rows
      const uint64_t k_mn = make_smem_desc(k_addr, 8192, 1024);                                               // as B of dQ
      const uint64_t p_mn = make_smem_desc(p_addr, 16384, 1024), ds_mn = make_smem_desc(ds_addr, 16384, 1024);
      const uint64_t ds_km = make_smem_desc(ds_addr, 0, 1024);
      uint64_t q_km[2], do_km[2], q_mn[2], do_mn[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uint32_t qa = smem_u32(sQ + s * 16384), da = smem_u32(sDO + s * 16384);
        q_km[s] = make_smem_desc(qa, 0, 1024);
        do_km[s] = make_smem_desc(da, 0, 1024);
        q_mn[s] = make_smem_desc(qa, 8192, 1024);
        do_mn[s] = make_smem_desc(da, 8192, 1024);
      }
      // S = Q K^T and dP = dO V^T of step `step` over the whole key block: N = nb (<= 128) per instruction.  The small
      // MMAs of this kernel read both operands from shared memory and are bound by that read, not by the tensor pipe
      // (~90 cycles per M=128 x N=64 x K=16 instruction against a 32-cycle floor, profiles/r02_attn_bwd_trace_after_hoist.log):
      // one N=128 instruction re-reads the 4 KB A slice once instead of twice.
      auto issue_sdp = [&](int step) {
        const int kb = step / p.nqt;
        const int nb = min(128, p.Kp - kb * 128);
        const uint32_t idesc_s = make_idesc_bf16(128, nb, 0, 0);
        const int s = q_resident ? step % p.nqt : (step & 1);
        mbar_wait(&qfull[s], q_resident ? 0u : (uint32_t)((step >> 1) & 1));
        tc_fence_after();
        const uint32_t kv_off = (uint32_t)(kb * 16384) >> 4;
        const uint64_t qd = s ? q_km[1] : q_km[0], dod = s ? do_km[1] : do_km[0];
        const uint64_t kd = k_km + kv_off, vd = v_km + kv_off;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_S, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_DP, dod + 2 * k, vd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_sdp);
      };
      mbar_wait(bar_k, 0);
      mbar_wait(bar_v, 0);
      TR(66);
      issue_sdp(0);
      TR(67);
      int step = 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int nb = min(128, p.Kp - kb * 128);
        for (int qt = 0; qt < p.nqt; ++qt, ++step) {
          const int s = q_resident ? qt : (step & 1);
          mbar_wait(bar_cons, step & 1);   // S/dP(step) are in registers: the columns can be overwritten
          tc_fence_after();
          if (step + 1 < total_steps) issue_sdp(step + 1);
          TR(68 + step * 6 + 1);
          mbar_wait(bar_pds, step & 1);   // P/dS(step) in smem
          TR(68 + step * 6 + 2);
          if (qt == 0 && kb > 0) mbar_wait(bar_dkv_free, (kb - 1) & 1);
          tc_fence_after();
          const uint32_t acc0 = qt > 0 ? 1u : 0u;
          const uint64_t dod = s ? do_mn[1] : do_mn[0], qd = s ? q_mn[1] : q_mn[0];
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // dV += P^T dO      (k-step = 16 q rows = 2048 B of either operand)
            umma_bf16(tmem_base + COL_DV, p_mn + 128 * ks, dod + 128 * ks, idesc_t, ks > 0 ? 1u : acc0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // dK += dS^T Q
            umma_bf16(tmem_base + COL_DK, ds_mn + 128 * ks, qd + 128 * ks, idesc_t, ks > 0 ? 1u : acc0);
          const uint32_t col_dq = (dq_resident && qt == 1) ? COL_DQ1 : COL_DQ0;
          const uint32_t accq = (dq_resident && kb > 0) ? 1u : 0u;
          const int ksteps = nb >> 4;
          const uint64_t kd = k_mn + (uint32_t)(kb * 16384 >> 4);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // dQ (+)= dS K_kb    (k-step = 16 keys: +32 B inside a 64-key block of dS, +2048 B of K)
            if (ks < ksteps)
              umma_bf16(tmem_base + col_dq, ds_km + (ks >> 2) * 1024 + (ks & 3) * 2, kd + 128 * ks, idesc_q,
                        ks > 0 ? 1u : accq);
          umma_commit(&qempty[s]);
          umma_commit(bar_free);
          TR(68 + step * 6 + 3);
        }
      }
    }
  } else {
    // -------------------------------------------------------------------- row threads
    const int row = tid & 127;
    const int half = tid >> 7;   // key-column half of every 128-key block
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const size_t H3 = (size_t)3 * p.H;
    const uint32_t dseed = p.drop_thr != 0 ? __ldg(p.drop_seed) : 0u;

    // everything a finished step leaves behind in TMEM: dQ of its q tile (when due) and, after the last q tile of a
    // key block, that block's dK (column half 0) / dV (half 1)
    auto read_out = [&](int kb, int qt) {
      const int q = qt * 128 + row;
      const bool alive = q < p.S;
      const bool dq_due = dq_resident ? (kb == p.nkb - 1) : true;
      if (dq_due) {
        float dq[32];
        const uint32_t col_dq = (dq_resident && qt == 1) ? COL_DQ1 : COL_DQ0;
        tmem_ld_32x32(taddr + col_dq + half * 32, reinterpret_cast<uint32_t*>(dq));
        tmem_ld_wait();
        if (alive) {
          bool final = true;
          if (!dq_resident) {
            float4* ws = reinterpret_cast<float4*>(p.dq_ws + (((size_t)n * p.heads + h) * p.S + q) * 64 + half * 32);
            if (kb > 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 f = ws[j];
                dq[4 * j] += f.x;
                dq[4 * j + 1] += f.y;
                dq[4 * j + 2] += f.z;
                dq[4 * j + 3] += f.w;
              }
            }
            if (kb != p.nkb - 1) {
              final = false;
#pragma unroll
              for (int j = 0; j < 8; ++j) ws[j] = make_float4(dq[4 * j], dq[4 * j + 1], dq[4 * j + 2], dq[4 * j + 3]);
            }
          }
          if (final) {
            uint4* out = reinterpret_cast<uint4*>(p.dqkv + ((size_t)n * p.S + q) * H3 + h * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              pk.x = pack_bf16(dq[8 * j + 0], dq[8 * j + 1]);
              pk.y = pack_bf16(dq[8 * j + 2], dq[8 * j + 3]);
              pk.z = pack_bf16(dq[8 * j + 4], dq[8 * j + 5]);
              pk.w = pack_bf16(dq[8 * j + 6], dq[8 * j + 7]);
              out[j] = pk;
            }
          }
        }
      }
      if (qt == p.nqt - 1) {
        const int kk = kb * 128 + row;
        bf16* dst = nullptr;
        if (kk < p.P)
          dst = p.dpkv + ((size_t)n * p.P + kk) * (2 * p.H) + h * 64 + half * p.H;
        else if (kk >= p.Ppad && kk - p.Ppad < p.S)
          dst = p.dqkv + ((size_t)n * p.S + (kk - p.Ppad)) * H3 + p.H + h * 64 + half * p.H;
        float acc[64];
        const uint32_t col = half == 0 ? COL_DK : COL_DV;
        if (tid == 0) TR(110 + kb * 3);
        tmem_ld_32x32(taddr + col, reinterpret_cast<uint32_t*>(acc));
        tmem_ld_32x32(taddr + col + 32, reinterpret_cast<uint32_t*>(acc) + 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_dkv_free);
        if (tid == 0) TR(111 + kb * 3);
        if (dst != nullptr) {
          uint4* out = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 pk;
            pk.x = pack_bf16(acc[8 * j + 0], acc[8 * j + 1]);
            pk.y = pack_bf16(acc[8 * j + 2], acc[8 * j + 3]);
            pk.z = pack_bf16(acc[8 * j + 4], acc[8 * j + 5]);
            pk.w = pack_bf16(acc[8 * j + 6], acc[8 * j + 7]);
            out[j] = pk;
          }
        }
      }
    };

    int step = 0;
    int prev_kb = 0, prev_qt = 0;
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int nb = min(128, p.Kp - kb * 128);
      for (int qt = 0; qt < p.nqt; ++qt, ++step) {
        const int q = qt * 128 + row;
        const bool alive = q < p.S;
        const bool q_txt = q < p.T;
        const uint32_t drow = p.drop_thr != 0 ? drop_hash(dseed, p.drop_site, (uint32_t)((n * p.heads + h) * p.S + q)) : 0u;
        const float lse2 = sLse[q];
        const float delta = sDelta[q];
        const uint32_t* crow = (MODE != 0 && alive && q_txt) ? p.cbits + ((size_t)n * p.T + q) * p.cwords : nullptr;
        const float* darow =
            (p.dalign != nullptr && alive && q_txt && (p.row_sel == nullptr || __ldg(p.row_sel + (size_t)n * p.T + q) != 0))
                ? p.dalign + ((size_t)n * p.T + q) * p.R : nullptr;
        // this thread's two 32-column chunks of the 128-key block are interleaved with the other half's
        // (c0 = 32*half and 32*half + 64) so that a partial last block stays balanced between the two warps
        const bool warp_alive = (qt * 128 + (warp & 3) * 32) < p.S;
        const float delta_s = delta * p.scale;
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int c0 = half * 32 + i * 64;
          const int kk0 = kb * 128 + c0;
          const bool live = warp_alive && c0 < nb;
          float sv[32], dp[32];
          uint32_t pp[16], dd[16];
          if (i == 0) {
            mbar_wait(bar_sdp, step & 1);
            tc_fence_after();
          }
          if (tid == 0) TR(3 + step * 12 + i * 6);
          if (live) {
            tmem_ld_32x32(taddr + COL_S + c0, reinterpret_cast<uint32_t*>(sv));
            tmem_ld_32x32(taddr + COL_DP + c0, reinterpret_cast<uint32_t*>(dp));
            tmem_ld_wait();
          }
          if (i == 1) {   // both chunks of this thread are in registers
            tc_fence_before();
            mbar_arrive(bar_cons);
          }
          if (tid == 0) TR(4 + step * 12 + i * 6);
          if (!live) {
            // dead q rows (rows >= S) or key columns beyond this block: P = dS = 0, no TMEM traffic, no math
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pp[j] = 0u;
              dd[j] = 0u;
            }
          } else {
            const uint32_t w = (crow != nullptr && (kk0 >> 5) < p.cwords) ? __ldg(crow + (kk0 >> 5)) : 0u;
            if (c0 + 32 > nb) {   // partial last key block: TMEM columns beyond the MMA width hold stale data
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (c0 + j >= nb) {
                  sv[j] = 0.f;
                  dp[j] = 0.f;
                }
              }
            }
            if (darow != nullptr && kk0 + 32 > p.T && kk0 < p.S) {   // region columns: + gradient of the alignment loss
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int kk = kk0 + j;
                if (kk >= p.T && kk < p.S) dp[j] += __ldg(darow + (kk - p.T));
              }
            }
            bwd_chunk_math<MODE>(p, c1, sv, dp, sBias, kk0, q, q_txt, w, lse2, delta_s, pp, dd, drow);
          }
          if (tid == 0) TR(5 + step * 12 + i * 6);
          if (i == 0 && step > 0) {
            // the MMAs of the previous step have retired (they ran under the math above): its P/dS tiles are free
            // and its accumulators readable
            mbar_wait(bar_free, (step - 1) & 1);
            if (tid == 0) TR(6 + step * 12);
            tc_fence_after();
            // flush mode re-uses the single dQ accumulator every step: drain it before this step's MMAs may start
            if (!dq_resident) read_out(prev_kb, prev_qt);
          }
          const uint32_t pb = smem_u32(sP + (c0 >> 6) * 16384), db = smem_u32(sDS + (c0 >> 6) * 16384);
          const int piece0 = (c0 & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sts128(pb + sw128_offset(row, piece0 + j), pp[4 * j], pp[4 * j + 1], pp[4 * j + 2], pp[4 * j + 3]);
            sts128(db + sw128_offset(row, piece0 + j), dd[4 * j], dd[4 * j + 1], dd[4 * j + 2], dd[4 * j + 3]);
          }
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(bar_pds);
        if (tid == 0) TR(7 + step * 12);
        // resident mode: nothing this step's MMAs write is still unread (dK/dV of a new key block wait for
        // bar_dkv_free), so the read-out runs after the hand-off, when the P/dS registers are dead
        if (step > 0 && dq_resident) read_out(prev_kb, prev_qt);
        if (tid == 0) TR(8 + step * 12);
        prev_kb = kb;
        prev_qt = qt;
      }
    }
    mbar_wait(bar_free, (step - 1) & 1);
    tc_fence_after();
    if (tid == 0) TR(60);
    read_out(prev_kb, prev_qt);
    if (tid == 0) TR(61);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) TR(62);
  if (warp == 8) tmem_dealloc<512>(tmem_base);
}

#ifdef MODCR_ATTN_TRACE
int attn_trace_read(long long* host, int max_entries) {
  const int n = TRACE_CTAS * TRACE_SLOTS < max_entries ? TRACE_CTAS * TRACE_SLOTS : max_entries;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host, g_attn_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? n : -1;
}
#else
int attn_trace_read(long long*, int) { return 0; }
#endif

// ============================================================================ host
static int fill_params(const AttnDesc& a, AttnParams& p, bool bwd) {
  if (a.N <= 0 || a.S <= 0 || a.heads <= 0 || a.P < 0 || a.P > 64) {
    set_error("prefix_attn: bad shape N=%d S=%d P=%d heads=%d (P <= 64)", a.N, a.S, a.P, a.heads);
    return MODCR_ERR_SHAPE;
  }
  if (a.mode < 0 || a.mode > 2) {
    set_error("prefix_attn: bad mode %d", a.mode);
    return MODCR_ERR_SHAPE;
  }
  if (a.mode != 0 && (a.P != 0 || a.T <= 0 || a.T > a.S || a.cbits == nullptr || a.cwords * 32 < a.T)) {
    set_error("prefix_attn: chunk modes need P=0, 0<T<=S, a chunk bit mask with cwords*32 >= T");
    return MODCR_ERR_SHAPE;
  }
  if ((a.probs_img != nullptr || a.dalign != nullptr) && (a.P != 0 || a.T <= 0 || a.T >= a.S)) {
    set_error("prefix_attn: the text->region probability block needs P=0 and 0<T<S");
    return MODCR_ERR_SHAPE;
  }
  if (a.qkv == nullptr || a.kbias == nullptr || (a.P > 0 && a.pkv == nullptr)) {
    set_error("prefix_attn: null input pointer");
    return MODCR_ERR_SHAPE;
  }
  p.N = a.N; p.S = a.S; p.P = a.P; p.heads = a.heads; p.H = a.heads * 64;
  p.Ppad = (a.P + 15) & ~15;
  p.T = a.T; p.R = a.S - a.T;
  p.Kp = p.Ppad + ((a.S + 15) & ~15);
  p.Krows = p.Ppad + 64 * ((a.S + 63) / 64);
  p.nqt = (a.S + 127) / 128;
  p.nkb = (p.Kp + 127) / 128;
  p.mode = a.mode; p.scale = a.scale;
  p.kbias = a.kbias; p.cbits = a.cbits; p.cwords = a.cwords;
  p.lse = a.lse; p.probs_img = a.probs_img;
  p.ctx = (const bf16*)a.ctx; p.dctx = (const bf16*)a.dctx; p.dalign = a.dalign;
  p.dqkv = (bf16*)a.dqkv; p.dpkv = (bf16*)a.dpkv; p.dq_ws = a.dq_ws;
  p.kid_split = p.Ppad; p.kid_sub = p.Ppad;
  p.row_sel = a.row_sel;
  p.drop_seed = a.drop_seed; p.drop_site = a.drop_site; p.drop_thr = 0; p.drop_inv = 1.f;
  if (a.drop_p > 0.f) {
    if (!(a.drop_p < 1.f) || a.drop_seed == nullptr) {
      set_error("prefix_attn: dropout needs 0 <= p < 1 and a device seed pointer");
      return MODCR_ERR_SHAPE;
    }
    if ((double)a.N * a.heads * a.S * (a.S + a.P) >= 4294967296.0) {
      set_error("prefix_attn: dropout mask index exceeds 32 bits");
      return MODCR_ERR_UNSUPPORTED;
    }
    p.drop_thr = (uint32_t)fmin(4294967295.0, (double)a.drop_p * 4294967296.0);
    if (p.drop_thr == 0) p.drop_thr = 1;
    p.drop_inv = 1.0f / (1.0f - a.drop_p);
  }
  if (a.lse == nullptr || a.ctx == nullptr) {
    set_error("prefix_attn: null ctx / lse pointer");
    return MODCR_ERR_SHAPE;
  }
  if (bwd) {
    if (a.dctx == nullptr || a.dqkv == nullptr || (a.P > 0 && a.dpkv == nullptr)) {
      set_error("prefix_attn_bwd: null gradient pointer");
      return MODCR_ERR_SHAPE;
    }
    if (p.nkb > 1 && p.nqt > 2 && a.dq_ws == nullptr) {
      set_error("prefix_attn_bwd: workspace required when S > 256 and P+S > 128");
      return MODCR_ERR_WORKSPACE;
    }
  }
  return MODCR_OK;
}

size_t attn_bwd_workspace_bytes(int N, int S, int heads) { return (size_t)N * heads * S * 64 * sizeof(float); }

template <int TMEM_COLS, int MODE>
static int launch_fwd(const AttnDesc& a, const AttnParams& p, const CUtensorMap& tmQKV, const CUtensorMap& tmPKV,
                      const CUtensorMap& tmCTX, size_t smem, const FwdLayout& lay, cudaStream_t stream) {
  auto kern = attn_fwd_kernel<TMEM_COLS, MODE>;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("prefix_attn_fwd: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
      return MODCR_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(p.nqt, p.heads, p.N);
  kern<<<grid, 160, smem, stream>>>(tmQKV, tmPKV, tmCTX, p, lay);
  return check_launch("prefix_attn_fwd");
}

template <int MODE>
static int launch_fwd_persist(const AttnParams& p, const CUtensorMap& tmQKV, const CUtensorMap& tmTail,
                              const CUtensorMap& tmPKV, const CUtensorMap& tmA, const CUtensorMap& tmB, size_t smem,
                              const PersistLayout& lay, cudaStream_t stream) {
  auto kern = attn_fwd_persist_kernel<MODE>;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("prefix_attn_fwd: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
      return MODCR_ERR_CUDA;
    }
    attr = smem;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int total = p.N * p.heads;
  const int lim = sm_limit();
  const int avail = (lim > 0 && lim < sms) ? lim : sms;
  const int grid = total < avail ? total : avail;
  kern<<<grid, 256, smem, stream>>>(tmQKV, tmTail, tmPKV, tmA, tmB, p, lay);
  return check_launch("prefix_attn_fwd");
}

// MODCR_ATTN_FWD_TILE=1 forces the one-CTA-per-q-tile kernel (A/B comparison; default: persistent where it applies)
static bool attn_fwd_force_tile() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MODCR_ATTN_FWD_TILE");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int attn_fwd_launch(const AttnDesc& a, cudaStream_t stream) {
  AttnParams p;
  int rc = fill_params(a, p, false);
  if (rc) return rc;
  const int kp2 = (p.S + p.P + 15) & ~15;   // [sequence ; prefix] without inner padding
  if (p.S <= 192 && kp2 <= 192 && !attn_fwd_force_tile()) {
    PersistLayout lay;
    const int kbytes = (kp2 * 128 + 1023) & ~1023;
    lay.off_qb = 16384;
    lay.off_k = 32768;
    lay.off_v = 32768 + (kbytes < 16384 ? 16384 : kbytes);   // P block 2 of group A needs 16 KB behind off_k
    lay.stage_bytes = lay.off_v + kbytes;
    lay.off_pb = 2 * lay.stage_bytes;
    lay.off_oa = lay.off_pb + 32768;
    lay.off_bias = lay.off_oa + 16384;
    lay.off_bars = lay.off_bias + 2 * 2 * 256 * 4;
    const size_t smem = 1024 + (size_t)lay.off_bars + 128;
    if (smem <= 232448) {
      AttnParams pp = p;
      pp.Kp = kp2;
      pp.kid_split = 0; pp.kid_sub = 0;
      CUtensorMap tmQKV, tmTail, tmPKV, tmA, tmB;
      rc = make_tmap_3d(&tmQKV, a.qkv, 2, p.N, p.S, 3 * p.H, 3 * p.H, (uint64_t)p.S * 3 * p.H, 64, 64);
      if (rc) return rc;
      tmTail = tmQKV;
      if (p.S & 63) {
        rc = make_tmap_3d(&tmTail, a.qkv, 2, p.N, p.S, 3 * p.H, 3 * p.H, (uint64_t)p.S * 3 * p.H, p.S & 63, 64);
        if (rc) return rc;
      }
      tmPKV = tmQKV;
      if (p.P > 0) {
        rc = make_tmap_3d(&tmPKV, a.pkv, 2, p.N, p.P, 2 * p.H, 2 * p.H, (uint64_t)p.P * 2 * p.H, p.P, 64);
        if (rc) return rc;
      }
      rc = make_tmap_3d(&tmA, a.ctx, 2, p.N, p.S, p.H, p.H, (uint64_t)p.S * p.H, 128, 64);
      if (rc) return rc;
      rc = make_tmap_3d(&tmB, a.ctx, 2, p.N, p.S, p.H, p.H, (uint64_t)p.S * p.H, 64, 64);
      if (rc) return rc;
      switch (p.mode) {
        case 0: return launch_fwd_persist<0>(pp, tmQKV, tmTail, tmPKV, tmA, tmB, smem, lay, stream);
        case 1: return launch_fwd_persist<1>(pp, tmQKV, tmTail, tmPKV, tmA, tmB, smem, lay, stream);
        default: return launch_fwd_persist<2>(pp, tmQKV, tmTail, tmPKV, tmA, tmB, smem, lay, stream);
      }
    }
  }
  const int nblk = (p.Kp + 63) / 64;
  // shared-memory plan: Q tile | K | V | extra P blocks | bias | barriers.  P block 0 overlays the Q tile and
  // P block 1 the K tile (both dead once S = QK^T has completed), which keeps cfg-2 under 113 KB -> 2 CTAs / SM.
  FwdLayout lay;
  int off = 16384;
  lay.off_k = off; off += p.Krows * 128;
  lay.off_v = off; off += p.Krows * 128;
  for (int j = 0; j < 8; ++j) lay.off_p[j] = 0;
  for (int j = 0; j < nblk && j < 8; ++j) {
    if (j == 0) lay.off_p[j] = 0;
    else if (j == 1 && p.Krows * 128 >= 16384) lay.off_p[j] = lay.off_k;
    else { lay.off_p[j] = off; off += 16384; }
  }
  lay.off_bias = off; off += nblk * 64 * 4;
  lay.off_bars = (off + 15) & ~15;
  const size_t smem = 1024 + (size_t)lay.off_bars + 128;
  if (p.Kp > 512 || smem > 232448) {
    set_error("prefix_attn_fwd: P+S=%d too long for the single-pass kernel (smem %zu B)", p.P + p.S, smem);
    return MODCR_ERR_UNSUPPORTED;
  }
  CUtensorMap tmQKV, tmPKV, tmCTX;
  rc = make_tmap_3d(&tmQKV, a.qkv, 2, p.N, p.S, 3 * p.H, 3 * p.H, (uint64_t)p.S * 3 * p.H, 64, 64);
  if (rc) return rc;
  tmPKV = tmQKV;
  if (p.P > 0) {
    rc = make_tmap_3d(&tmPKV, a.pkv, 2, p.N, p.P, 2 * p.H, 2 * p.H, (uint64_t)p.P * 2 * p.H, 16, 64);
    if (rc) return rc;
  }
  rc = make_tmap_3d(&tmCTX, a.ctx, 2, p.N, p.S, p.H, p.H, (uint64_t)p.S * p.H, 128, 64);
  if (rc) return rc;
  const bool small = p.Kp <= 256;
  switch (p.mode) {
    case 0:
      return small ? launch_fwd<256, 0>(a, p, tmQKV, tmPKV, tmCTX, smem, lay, stream)
                   : launch_fwd<512, 0>(a, p, tmQKV, tmPKV, tmCTX, smem, lay, stream);
    case 1:
      return small ? launch_fwd<256, 1>(a, p, tmQKV, tmPKV, tmCTX, smem, lay, stream)
                   : launch_fwd<512, 1>(a, p, tmQKV, tmPKV, tmCTX, smem, lay, stream);
    default:
      return small ? launch_fwd<256, 2>(a, p, tmQKV, tmPKV, tmCTX, smem, lay, stream)
                   : launch_fwd<512, 2>(a, p, tmQKV, tmPKV, tmCTX, smem, lay, stream);
  }
}

template <int MODE>
static int launch_bwd(const AttnParams& p, const CUtensorMap& tmQKV, const CUtensorMap& tmPKV, const CUtensorMap& tmDO,
                      size_t smem, cudaStream_t stream) {
  auto kern = attn_bwd_kernel<MODE>;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("prefix_attn_bwd: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
      return MODCR_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(p.heads, p.N);
  kern<<<grid, 320, smem, stream>>>(tmQKV, tmPKV, tmDO, p);
  return check_launch("prefix_attn_bwd");
}

int attn_bwd_launch(const AttnDesc& a, cudaStream_t stream) {
  AttnParams p;
  int rc = fill_params(a, p, true);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)2 * p.Krows * 128 + 8 * 16384 + (size_t)(p.nkb * 128 + 2 * p.nqt * 128) * 4 + 128;
  if (smem > 232448) {
    set_error("prefix_attn_bwd: P+S=%d too long (smem %zu B)", p.P + p.S, smem);
    return MODCR_ERR_UNSUPPORTED;
  }
  CUtensorMap tmQKV, tmPKV, tmDO;
  rc = make_tmap_3d(&tmQKV, a.qkv, 2, p.N, p.S, 3 * p.H, 3 * p.H, (uint64_t)p.S * 3 * p.H, 64, 64);
  if (rc) return rc;
  tmPKV = tmQKV;
  if (p.P > 0) {
    rc = make_tmap_3d(&tmPKV, a.pkv, 2, p.N, p.P, 2 * p.H, 2 * p.H, (uint64_t)p.P * 2 * p.H, 16, 64);
    if (rc) return rc;
  }
  rc = make_tmap_3d(&tmDO, a.dctx, 2, p.N, p.S, p.H, p.H, (uint64_t)p.S * p.H, 64, 64);
  if (rc) return rc;
  switch (p.mode) {
    case 0: return launch_bwd<0>(p, tmQKV, tmPKV, tmDO, smem, stream);
    case 1: return launch_bwd<1>(p, tmQKV, tmPKV, tmDO, smem, stream);
    default: return launch_bwd<2>(p, tmQKV, tmPKV, tmDO, smem, stream);
  }
}

}  // namespace modcr
