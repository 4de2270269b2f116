// HBM-bound row kernels of the ModCR hot path (sm_100a): LayerNorm forward/backward, the fused
// text-embedding gather + LayerNorm (a_transformers/modeling_bert.py:184-211), casts / packing,
// column sums (bias gradients), row copies between [N,S,H] buffers and the chunk-mean query
// operator (modeling_vcr_chunkalign_v10.py:66-78).  All activations are bf16, statistics and
// parameter gradients fp32; every global access is a 128-bit vector access, one warp per row.
#include "rowops.h"

#include <cuda_bf16.h>
#include <math.h>

#include "ptx.cuh"
#include "tma_host.h"

namespace modcr {

typedef __nv_bfloat16 bf16;

DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
DEVINL void unpack8(const uint4& u, float* f) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
DEVINL uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}
DEVINL void load8f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// physical row of logical row m: rows are grouped `rows_per_seq` at a time, each group starts
// `seq_stride` rows after the previous one, offset by `row0` (writes text rows / region rows of
// the concatenated [N,S,H] buffer without a torch.cat)
DEVINL int64_t map_row(int m, const RowMap& r) {
  if (r.rows_per_seq <= 0) return m;
  const int s = m / r.rows_per_seq;
  return (int64_t)s * r.seq_stride + r.row0 + (m - s * r.rows_per_seq);
}

constexpr int LN_MAXV = 4;  // up to 4 x 8 x 32 = 1024 columns

// ---------------------------------------------------------------------------- LayerNorm fwd
// y = (x - mean) * rstd * gamma + beta ; two-pass statistics in fp32 registers
// load 8 consecutive elements of a row that is either bf16 or fp32
template <bool F32>
DEVINL void load_row8(const void* row, int vec, float* f) {
  if (F32) {
    load8f(reinterpret_cast<const float*>(row) + vec * 8, f);
  } else {
    unpack8(__ldg(reinterpret_cast<const uint4*>(row) + vec), f);
  }
}
DEVINL void store8f(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

template <bool XF32>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x, RowMap xmap, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float* __restrict__ y32, RowMap ymap, float* __restrict__ mean,
                                                     float* __restrict__ rstd, int M, int H, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nv = H >> 8;  // vectors per lane
  const char* xr = reinterpret_cast<const char*>(x) + map_row(warp, xmap) * H * (XF32 ? 4 : 2);
  float v[LN_MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      load_row8<XF32>(xr, lane + 32 * i, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  const float mu = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mu;
        q += d * d;
      }
    }
  const float rs = rsqrtf(warp_sum(q) / H + eps);
  const int64_t orow = map_row(warp, ymap);
  bf16* yr = y + orow * H;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      const int c = (lane + 32 * i) * 8;
      float g[8], b[8], o[8];
      load8f(gamma + c, g);
      load8f(beta + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
      reinterpret_cast<uint4*>(yr)[lane + 32 * i] = pack8(o);
      if (y32 != nullptr) store8f(y32 + orow * H + c, o);   // fp32 copy feeds the next residual add
    }
  if (lane == 0) {
    if (mean) mean[warp] = mu;
    if (rstd) rstd[warp] = rs;
  }
}

// ---------------------------------------------------------------------------- box-position + type embedding (g1)
// BertImgModel_gtToken (modeling/modeling_bert.py:312-316, 390-394): every region row gets, next to its 2054 -> H feature
// projection,   e[j] = img_token_type_embeddings[tt][j] + LayerNorm_pos(img_pos_linear(pos))[j]
// with pos the 7 box numbers of the region.  K = 7 is no tensor-core shape: one warp per row computes the 7-term
// products in fp32 (W_pos [H,7] staged transposed in shared memory), normalises and adds the type row; the result is the
// fp32 epilogue input of the feature-projection GEMM, which adds it to acc + bias (modcr_gemm in_op 1).
constexpr int POSD = 7;

__global__ void __launch_bounds__(256) pos_type_fwd_kernel(const float* __restrict__ pos, const int64_t* __restrict__ tt,
                                                           const float* __restrict__ Wp, const float* __restrict__ bp,
                                                           const float* __restrict__ gam, const float* __restrict__ bet,
                                                           const float* __restrict__ type_emb, float* __restrict__ out,
                                                           float* __restrict__ mean, float* __restrict__ rstd, int rows, int H,
                                                           float eps) {
  extern __shared__ float sW[];   // [POSD][H] transposed, then bias [H]
  for (int i = threadIdx.x; i < H * POSD; i += blockDim.x) sW[(i % POSD) * H + i / POSD] = __ldg(Wp + i);
  for (int i = threadIdx.x; i < H; i += blockDim.x) sW[POSD * H + i] = __ldg(bp + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nv = H >> 8;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += gridDim.x * (blockDim.x >> 5)) {
    float pv[POSD];
#pragma unroll
    for (int k = 0; k < POSD; ++k) pv[k] = __ldg(pos + (size_t)row * POSD + k);
    float t[LN_MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = sW[POSD * H + c + j];
#pragma unroll
          for (int k = 0; k < POSD; ++k) a = fmaf(pv[k], sW[k * H + c + j], a);
          t[i][j] = a;
          s += a;
        }
      }
    const float mu = warp_sum(s) / (float)H;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = t[i][j] - mu;
          v = fmaf(d, d, v);
        }
      }
    const float rs = rsqrtf(warp_sum(v) / (float)H + eps);
    const float* te = type_emb + (size_t)(tt != nullptr ? tt[row] : 0) * H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 8;
        float g[8], b[8], e[8], o[8];
        load8f(gam + c, g);
        load8f(bet + c, b);
        load8f(te + c, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (t[i][j] - mu) * rs * g[j] + b[j] + e[j];
        store8f(out + (size_t)row * H + c, o);
      }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

// backward of the above from d(sum) [rows,H] (bf16, = the gradient of the feature-projection GEMM's output):
//   dtype_emb[tt] += d ; through LayerNorm_pos: dgamma += d * that, dbeta += d, dt = rstd (g - mean(g) - that mean(g that)),
//   g = d * gamma ; dW_pos[j,k] += dt[j] pos[k] ; db_pos[j] += dt[j]
// Column sums are kept per CTA in shared memory over its rows and leave with one atomicAdd per element and CTA.
__global__ void __launch_bounds__(256) pos_type_bwd_kernel(const bf16* __restrict__ d, const float* __restrict__ pos,
                                                           const int64_t* __restrict__ tt, const float* __restrict__ Wp,
                                                           const float* __restrict__ bp, const float* __restrict__ gam,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float* __restrict__ dWp, float* __restrict__ dbp,
                                                           float* __restrict__ dgam, float* __restrict__ dbet,
                                                           float* __restrict__ dtype, int rows, int H) {
  extern __shared__ float sm[];
  float* sW = sm;                      // [POSD][H] + bias [H]
  float* acc = sW + (POSD + 1) * H;    // dW^T [POSD][H] | db [H] | dgam [H] | dbet [H] | dtype [2][H]
  const int nacc = (POSD + 5) * H;
  for (int i = threadIdx.x; i < H * POSD; i += blockDim.x) sW[(i % POSD) * H + i / POSD] = __ldg(Wp + i);
  for (int i = threadIdx.x; i < H; i += blockDim.x) sW[POSD * H + i] = __ldg(bp + i);
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nv = H >> 8;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += gridDim.x * (blockDim.x >> 5)) {
    float pv[POSD];
#pragma unroll
    for (int k = 0; k < POSD; ++k) pv[k] = __ldg(pos + (size_t)row * POSD + k);
    const float mu = mean[row], rs = rstd[row];
    const int ty = tt != nullptr ? (int)tt[row] : 0;
    float xh[LN_MAXV][8], g[LN_MAXV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 8;
        float dv[8], gm[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(d + (size_t)row * H) + lane + 32 * i), dv);
        load8f(gam + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = sW[POSD * H + c + j];
#pragma unroll
          for (int k = 0; k < POSD; ++k) a = fmaf(pv[k], sW[k * H + c + j], a);
          xh[i][j] = (a - mu) * rs;
          g[i][j] = dv[j] * gm[j];
          s1 += g[i][j];
          s2 = fmaf(g[i][j], xh[i][j], s2);
          atomicAdd(acc + (POSD + 1) * H + c + j, dv[j] * xh[i][j]);   // dgamma
          atomicAdd(acc + (POSD + 2) * H + c + j, dv[j]);              // dbeta
          atomicAdd(acc + (POSD + 3 + ty) * H + c + j, dv[j]);         // dtype_emb[tt]
        }
      }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dt = rs * (g[i][j] - s1 - xh[i][j] * s2);
          atomicAdd(acc + POSD * H + c + j, dt);                       // db_pos
#pragma unroll
          for (int k = 0; k < POSD; ++k) atomicAdd(acc + k * H + c + j, dt * pv[k]);
        }
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * POSD; i += blockDim.x) atomicAdd(dWp + i, acc[(i % POSD) * H + i / POSD]);
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dbp + i, acc[POSD * H + i]);
    atomicAdd(dgam + i, acc[(POSD + 1) * H + i]);
    atomicAdd(dbet + i, acc[(POSD + 2) * H + i]);
    atomicAdd(dtype + i, acc[(POSD + 3) * H + i]);
    atomicAdd(dtype + H + i, acc[(POSD + 4) * H + i]);
  }
}

int pos_type_fwd(const float* pos, const int64_t* tt, const float* Wp, const float* bp, const float* gam, const float* bet,
                 const float* type_emb, float* out, float* mean, float* rstd, int rows, int H, float eps, cudaStream_t st) {
  if (rows <= 0 || H <= 0 || H % 256 != 0 || H > 1024) {
    set_error("pos_type_fwd: bad shape rows=%d H=%d (H a multiple of 256, <= 1024)", rows, H);
    return MODCR_ERR_SHAPE;
  }
  const size_t smem = (size_t)(POSD + 1) * H * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pos_type_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  const int blocks = rows < 8 * 148 ? (rows + 7) / 8 : 148 * 2;
  pos_type_fwd_kernel<<<blocks, 256, smem, st>>>(pos, tt, Wp, bp, gam, bet, type_emb, out, mean, rstd, rows, H, eps);
  return check_launch("pos_type_fwd");
}

int pos_type_bwd(const void* d, const float* pos, const int64_t* tt, const float* Wp, const float* bp, const float* gam,
                 const float* mean, const float* rstd, float* dWp, float* dbp, float* dgam, float* dbet, float* dtype,
                 int rows, int H, cudaStream_t st) {
  if (rows <= 0 || H <= 0 || H % 256 != 0 || H > 1024) {
    set_error("pos_type_bwd: bad shape rows=%d H=%d", rows, H);
    return MODCR_ERR_SHAPE;
  }
  const size_t smem = (size_t)((POSD + 1) + (POSD + 5)) * H * sizeof(float);   // 80 KB at H = 1024
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pos_type_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const int blocks = rows < 8 * 148 ? (rows + 7) / 8 : 148;
  pos_type_bwd_kernel<<<blocks, 256, smem, st>>>((const bf16*)d, pos, tt, Wp, bp, gam, mean, rstd, dWp, dbp, dgam, dbet, dtype,
                                                 rows, H);
  return check_launch("pos_type_bwd");
}

// ---------------------------------------------------------------------------- LayerNorm bwd
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; dbias += sum_rows dx (bias of the Linear
// that produced x, fused here so no extra pass over dx is needed)
constexpr int LNB_WARPS = 12;  // 384 threads, one persistent CTA per SM
template <bool XF32, int NV>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
ln_bwd_kernel(const bf16* __restrict__ dy, RowMap dymap, const void* __restrict__ x, RowMap xmap,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              bf16* __restrict__ dx, RowMap dxmap, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dbias, int M, int H, bf16* __restrict__ dxm, const uint32_t* __restrict__ dseed_ptr,
              uint32_t dsite, uint32_t dthr, float dinv) {
  // dxm (training-mode dropout of the producing dense layer, LayerNorm(dropout(dense(.)) + residual)): second output
  // dx * mask / (1 - p) -- the gradient w.r.t. the dense output -- and `dbias` then sums the masked gradient; `dx`
  // stays the gradient of the residual branch.  Mask index = m * H + column, as in the forward modcr_dropout call.
  __shared__ float red[LNB_WARPS][32 * 8 + 1];
  const uint32_t dseed = dthr != 0 ? __ldg(dseed_ptr) : 0u;
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * LNB_WARPS;
  float ag[NV][8], ab[NV][8], ax[NV][8], gm[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = ax[i][j] = 0.f;
    load8f(gamma + (lane + 32 * i) * 8, gm[i]);
  }
  for (int m = blockIdx.x * LNB_WARPS + wib; m < M; m += wstride) {
    const bf16* dyr = dy + map_row(m, dymap) * H;
    const char* xr = reinterpret_cast<const char*>(x) + map_row(m, xmap) * H * (XF32 ? 4 : 2);
    const float mu = __ldg(mean + m), rs = __ldg(rstd + m);
    float d[NV][8], xh[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {  // issue every load of the row before any arithmetic
      unpack8(__ldg(reinterpret_cast<const uint4*>(dyr) + lane + 32 * i), d[i]);
      load_row8<XF32>(xr, lane + 32 * i, xh[i]);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mu) * rs;
        const float g = d[i][j] * gm[i][j];
        s1 += g;
        s2 = fmaf(g, xh[i][j], s2);
        ag[i][j] = fmaf(d[i][j], xh[i][j], ag[i][j]);
        ab[i][j] += d[i][j];
      }
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
    bf16* dxr = dx + map_row(m, dxmap) * H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rs * (d[i][j] * gm[i][j] - s1 - xh[i][j] * s2);
      reinterpret_cast<uint4*>(dxr)[lane + 32 * i] = pack8(o);
      if (dthr != 0) {   // uniform
        const uint32_t base = (uint32_t)m * (uint32_t)H + (uint32_t)(lane + 32 * i) * 8u;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = drop_keep(dseed, dsite, base + j, dthr) ? o[j] * dinv : 0.f;
        reinterpret_cast<uint4*>(dxm + (size_t)m * H)[lane + 32 * i] = pack8(o);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ax[i][j] += o[j];
    }
  }
  // cross-warp reduction of the three column accumulators, then one atomic per column per CTA
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst == nullptr) continue;  // uniform
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[wib][lane * 8 + j] = which == 0 ? ag[i][j] : (which == 1 ? ab[i][j] : ax[i][j]);
      __syncthreads();
      const int t = threadIdx.x;
      if (t < 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LNB_WARPS; ++w) s += red[w][t];
        // slab i covers columns (lane + 32 i) * 8 + j  ->  t = lane*8 + j
        atomicAdd(dst + ((t >> 3) + 32 * i) * 8 + (t & 7), s);
      }
    }
  }
}

// ---------------------------------------------------------------------------- text embedding + LN
// e = word[ids] + type[tt] + pos[t] (fp32 tables), y = LN(e) written into rows [0,T) of [N,S,H]
__global__ void __launch_bounds__(256) txt_embed_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tts,
                                                            const int64_t* __restrict__ pos_ids,
                                                            const float* __restrict__ word, const float* __restrict__ type,
                                                            const float* __restrict__ pos, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ y,
                                                            float* __restrict__ y32, RowMap ymap, float* __restrict__ mean,
                                                            float* __restrict__ rstd, int M, int T, int H, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nv = H >> 8;
  const int64_t id = ids[warp];
  const int64_t tt = tts ? tts[warp] : 0;
  const int64_t ps = pos_ids ? pos_ids[warp] : (warp % T);
  float v[LN_MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      const int c = (lane + 32 * i) * 8;
      float a[8], b[8], d[8];
      load8f(word + id * H + c, a);
      load8f(type + tt * H + c, b);
      load8f(pos + ps * H + c, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = (a[j] + b[j]) + d[j];  // reference order: (word + type) + position
        s += v[i][j];
      }
    }
  const float mu = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mu;
        q += d * d;
      }
    }
  const float rs = rsqrtf(warp_sum(q) / H + eps);
  const int64_t orow = map_row(warp, ymap);
  bf16* yr = y + orow * H;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      const int c = (lane + 32 * i) * 8;
      float g[8], b[8], o[8];
      load8f(gamma + c, g);
      load8f(beta + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
      reinterpret_cast<uint4*>(yr)[lane + 32 * i] = pack8(o);
      if (y32 != nullptr) store8f(y32 + orow * H + c, o);
    }
  if (lane == 0) {
    mean[warp] = mu;
    rstd[warp] = rs;
  }
}

// backward: recompute e, LN backward, scatter-add into the fp32 table gradients (padding row skipped)
__global__ void __launch_bounds__(256) txt_embed_bwd_kernel(
    const bf16* __restrict__ dy, RowMap dymap, const int64_t* __restrict__ ids, const int64_t* __restrict__ tts,
    const int64_t* __restrict__ pos_ids, const float* __restrict__ word, const float* __restrict__ type,
    const float* __restrict__ pos, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ rstd, float* __restrict__ dword, float* __restrict__ dtype, float* __restrict__ dpos,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int T, int H, int64_t padding_idx) {
  __shared__ float red[8][32 * 8 + 1];
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nv = H >> 8;
  const int wstride = gridDim.x * 8;
  float ag[LN_MAXV][8], ab[LN_MAXV][8], at0[LN_MAXV][8], at1[LN_MAXV][8];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = at0[i][j] = at1[i][j] = 0.f;
  for (int m = blockIdx.x * 8 + wib; m < M; m += wstride) {
    const int64_t id = ids[m];
    const int64_t tt = tts ? tts[m] : 0;
    const int64_t ps = pos_ids ? pos_ids[m] : (m % T);
    const bf16* dyr = dy + map_row(m, dymap) * H;
    const float mu = __ldg(mean + m), rs = __ldg(rstd + m);
    float g[LN_MAXV][8], xh[LN_MAXV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 8;
        float d[8], a[8], b[8], e[8], gmv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dyr) + lane + 32 * i), d);
        load8f(word + id * H + c, a);
        load8f(type + tt * H + c, b);
        load8f(pos + ps * H + c, e);
        load8f(gamma + c, gmv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (((a[j] + b[j]) + e[j]) - mu) * rs;
          g[i][j] = d[j] * gmv[j];
          s1 += g[i][j];
          s2 += g[i][j] * xh[i][j];
          ag[i][j] += d[j] * xh[i][j];
          ab[i][j] += d[j];
        }
      }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const int c = (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float o = rs * (g[i][j] - s1 - xh[i][j] * s2);
          if (id != padding_idx) atomicAdd(dword + id * H + c + j, o);
          if (tt == 0) at0[i][j] += o;                 // the two hot token-type rows are reduced per CTA
          else if (tt == 1) at1[i][j] += o;
          else atomicAdd(dtype + tt * H + c + j, o);
          atomicAdd(dpos + ps * H + c + j, o);
        }
      }
  }
  for (int which = 0; which < 4; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : (which == 2 ? dtype : dtype + H));
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      if (i < nv) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          red[wib][lane * 8 + j] = which == 0 ? ag[i][j] : (which == 1 ? ab[i][j] : (which == 2 ? at0[i][j] : at1[i][j]));
        __syncthreads();
        const int t = threadIdx.x;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][t];
        const int col = ((t >> 3) + 32 * i) * 8 + (t & 7);
        atomicAdd(dst + col, s);
      }
    }
  }
}

// ---------------------------------------------------------------------------- casts / packing
// dst[r, 0:cols] = bf16(src[r, 0:cols]) ; dst[r, cols:ld_dst] = 0   (region features 2054 -> 2112, weights)
__global__ void cast_pad_kernel(const float* __restrict__ src, int64_t ld_src, bf16* __restrict__ dst, int64_t ld_dst,
                                int64_t rows, int cols, int64_t row_step) {
  const int64_t total = rows * (ld_dst >> 3);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (ld_dst >> 3);
    const int c = (int)(i - r * (ld_dst >> 3)) * 8;
    const float* s = src + r * row_step * ld_src + c;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c + j < cols) ? __ldg(s + j) : 0.f;
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = pack8(f);
  }
}

// img_embedding without LayerNorm (use_img_layernorm unset, modeling_transfomres.py:676-680): the fp32 GEMM output rows are
// placed at their [N,S,H] positions as bf16 + fp32 (forward); the bf16 gradient rows are gathered back (backward).
__global__ void rows_place_kernel(const float* __restrict__ src, bf16* __restrict__ y, float* __restrict__ y32, RowMap ymap,
                                  int M, int H8) {
  const int64_t total = (int64_t)M * H8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / H8), c = (int)(i - (int64_t)m * H8) * 8;
    float f[8];
    load8f(src + (int64_t)m * H8 * 8 + c, f);
    const int64_t o = map_row(m, ymap) * H8 * 8 + c;
    *reinterpret_cast<uint4*>(y + o) = pack8(f);
    if (y32 != nullptr) {
      reinterpret_cast<float4*>(y32 + o)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(y32 + o)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}

__global__ void rows_gather_kernel(const bf16* __restrict__ dy, RowMap dymap, bf16* __restrict__ dst, int M, int H8) {
  const int64_t total = (int64_t)M * H8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / H8), c = (int)(i - (int64_t)m * H8);
    reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(dy + map_row(m, dymap) * H8 * 8) + c);
  }
}

// dst(fp32) [+]= src(bf16) elementwise (used for gradients that must come back as fp32)
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n8, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src) + i), f);
    float4* d = reinterpret_cast<float4*>(dst) + 2 * i;
    if (accumulate) {
      const float4 a = d[0], b = d[1];
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// dst[r, c] (+)= src[r, c] for c < cols (fp32; un-pads the region-embedding weight gradient)
__global__ void accum_f32_2d_kernel(float* __restrict__ dst, int64_t ld_dst, const float* __restrict__ src, int64_t ld_src,
                                    int64_t rows, int cols, int accumulate) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = __ldg(src + r * ld_src + c);
    float* d = dst + r * ld_dst + c;
    *d = accumulate ? *d + v : v;
  }
}

// ---------------------------------------------------------------------------- column sums
// out[c] += sum_r x[r, c]    (bias gradients); block = 8 warps x 256 columns, rows strided by grid.y
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, int64_t ld, int M, int N,
                                                     float* __restrict__ out) {
  __shared__ float red[8][32 * 8 + 1];
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < N) {
    // four independent 128-bit loads in flight per thread (the loop is otherwise latency-bound)
    const int stride = gridDim.y * 8;
    for (int r = blockIdx.y * 8 + wib; r < M; r += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = r + i * stride;
        u[i] = rr < M ? __ldg(reinterpret_cast<const uint4*>(x + (int64_t)rr * ld + c)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float f[8];
        unpack8(u[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[wib][lane * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][t];
  const int col = blockIdx.x * 256 + t;
  if (col < N) atomicAdd(out + col, s);
}

// ---------------------------------------------------------------------------- row copies
// dst[seq, drow0 + r, :] (+)= src[seq, srow0 + r, :]   for r < nrows   (bf16, H % 8 == 0)
__global__ void copy_rows_kernel(const bf16* __restrict__ src, int64_t src_seq_stride, int srow0, bf16* __restrict__ dst,
                                 int64_t dst_seq_stride, int drow0, int nseq, int nrows, int H, int accumulate) {
  const int hv = H >> 3;
  const int64_t total = (int64_t)nseq * nrows * hv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % hv);
    const int64_t rr = i / hv;
    const int r = (int)(rr % nrows);
    const int64_t s = rr / nrows;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (s * src_seq_stride + srow0 + r) * H) + v);
    uint4* d = reinterpret_cast<uint4*>(dst + (s * dst_seq_stride + drow0 + r) * H) + v;
    if (accumulate) {
      float a[8], b[8];
      unpack8(u, a);
      unpack8(*d, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      *d = pack8(a);
    } else {
      *d = u;
    }
  }
}

// y = a + b (bf16, elementwise)
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a) + i), fa);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b) + i), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    reinterpret_cast<uint4*>(y)[i] = pack8(fa);
  }
}

// dx = dy * (1 - y^2)   (pooler tanh backward), rows strided
__global__ void tanh_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, bf16* __restrict__ dx, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= (1.f - b[j] * b[j]);
    reinterpret_cast<uint4*>(dx)[i] = pack8(a);
  }
}

// dx = dy * (y > 0)   (ReLU backward of the mlp classifier, modeling/modeling_bert.py:516-528; y = relu output)
__global__ void relu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, bf16* __restrict__ dx, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = b[j] > 0.f ? a[j] : 0.f;
    reinterpret_cast<uint4*>(dx)[i] = pack8(a);
  }
}

// LayerNorm of a few wide rows (the mlp classifier's BertLayerNorm over hidden*cls_hidden_scale columns,
// modeling/modeling_bert.py:519): one CTA per row, any H that is a multiple of 8.  Not a hot kernel (rows = questions).
DEVINL float cta_sum256(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(256) rowln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y,
                                                        float* __restrict__ mean, float* __restrict__ rstd, int H, float eps) {
  __shared__ float sh[8];
  const bf16* xr = x + (size_t)blockIdx.x * H;
  float s = 0.f;
  for (int c = threadIdx.x; c < H; c += 256) s += __bfloat162float(xr[c]);
  const float mu = cta_sum256(s, sh) / (float)H;
  float q = 0.f;
  for (int c = threadIdx.x; c < H; c += 256) {
    const float d = __bfloat162float(xr[c]) - mu;
    q += d * d;
  }
  const float rs = rsqrtf(cta_sum256(q, sh) / (float)H + eps);
  for (int c = threadIdx.x; c < H; c += 256)
    y[(size_t)blockIdx.x * H + c] = __float2bfloat16((__bfloat162float(xr[c]) - mu) * rs * gamma[c] + beta[c]);
  if (threadIdx.x == 0) {
    mean[blockIdx.x] = mu;
    rstd[blockIdx.x] = rs;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy * xhat, dbeta += dy  (atomics over rows)
__global__ void __launch_bounds__(256) rowln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, bf16* __restrict__ dx,
                                                        float* __restrict__ dgamma, float* __restrict__ dbeta, int H) {
  __shared__ float sh[8];
  const size_t base = (size_t)blockIdx.x * H;
  const float mu = mean[blockIdx.x], rs = rstd[blockIdx.x];
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < H; c += 256) {
    const float g = __bfloat162float(dy[base + c]) * gamma[c];
    const float xh = (__bfloat162float(x[base + c]) - mu) * rs;
    s1 += g;
    s2 += g * xh;
  }
  const float m1 = cta_sum256(s1, sh) / (float)H;
  const float m2 = cta_sum256(s2, sh) / (float)H;
  for (int c = threadIdx.x; c < H; c += 256) {
    const float d = __bfloat162float(dy[base + c]);
    const float xh = (__bfloat162float(x[base + c]) - mu) * rs;
    dx[base + c] = __float2bfloat16(rs * (d * gamma[c] - m1 - xh * m2));
    atomicAdd(dgamma + c, d * xh);
    atomicAdd(dbeta + c, d);
  }
}

// ---------------------------------------------------------------------------- masks
// kbias[n, k] = (1 - mask[n, k]) * -10000   (modeling_transfomres.py:628-641, no dtype cast)
__global__ void key_bias_kernel(const float* __restrict__ mask, float* __restrict__ kbias, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    kbias[i] = (1.0f - mask[i]) * -10000.0f;
}

// bits[n, q, w] bit j = (chunk_mask[n, q, 32 w + j] != 0)
// Exactness: the reference adds (1 - C) * -10000 with C in {0, 1} (v10:289-296); a bit reproduces it exactly.
__global__ void chunk_bits_kernel(const float* __restrict__ cmask, uint32_t* __restrict__ bits, int N, int T, int cwords) {
  const int64_t total = (int64_t)N * T * cwords;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % cwords);
    const int64_t row = i / cwords;
    const float* src = cmask + row * T + w * 32;
    uint32_t b = 0;
    for (int j = 0; j < 32; ++j)
      if (w * 32 + j < T && src[j] != 0.f) b |= (1u << j);
    bits[i] = b;
  }
}

// ---------------------------------------------------------------------------- chunk-mean queries
// In place on the Q columns of the fused QKV buffer (row stride ld): for tokens t in [1, 1+len[n]),
// Q[n,t,:] <- mean over tokens of the same chunk id (v10:66-78).  The operator is self-adjoint, so
// the backward pass applies the same kernel to dQ.
// One CTA per (sequence, 64-column slab): the [T x 64] slab is staged in shared memory with 128-bit row loads,
// chunk sums are formed in token order in fp32 (each chunk belongs to one quarter of the CTA, so the order of
// the additions is fixed), then every token row is rewritten with its chunk mean.
__global__ void __launch_bounds__(256) chunk_mean_kernel(bf16* __restrict__ q, int64_t ld, int S,
                                                         const int32_t* __restrict__ chunk_id, int T, int max_chunks) {
  extern __shared__ __align__(16) float csm[];
  float* sums = csm;                                                    // [max_chunks][64]
  int* cnt = reinterpret_cast<int*>(sums + (size_t)max_chunks * 64);    // [max_chunks rounded up to 4]
  int* cid = cnt + ((max_chunks + 3) & ~3);                             // [T rounded up to 4]
  bf16* tile = reinterpret_cast<bf16*>(cid + ((T + 3) & ~3));           // [T][64] (16-byte aligned)
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cg = tid & 7, rl = tid >> 3;
  for (int i = tid; i < max_chunks; i += 256) cnt[i] = 0;
  for (int i = tid; i < max_chunks * 64; i += 256) sums[i] = 0.f;
  for (int t = tid; t < T; t += 256) {
    const int c = chunk_id[(size_t)n * T + t];
    cid[t] = (c >= 0 && c < max_chunks) ? c : -1;
  }
  bf16* base = q + (size_t)n * S * ld + blockIdx.x * 64 + cg * 8;
  for (int t = rl; t < T; t += 32)
    *reinterpret_cast<uint4*>(tile + t * 64 + cg * 8) = *reinterpret_cast<const uint4*>(base + (size_t)t * ld);
  __syncthreads();
  for (int t = tid; t < T; t += 256)
    if (cid[t] >= 0) atomicAdd(&cnt[cid[t]], 1);
  {
    const int col = tid & 63, part = tid >> 6;
    for (int t = 0; t < T; ++t) {
      const int c = cid[t];
      if (c >= 0 && (c & 3) == part) sums[c * 64 + col] += __bfloat162float(tile[t * 64 + col]);
    }
  }
  __syncthreads();
  for (int t = rl; t < T; t += 32) {
    const int c = cid[t];
    if (c < 0) continue;
    const float k = (float)cnt[c];
    const float* sc = sums + c * 64 + cg * 8;
    uint4 pk;
    pk.x = pack_bf16(sc[0] / k, sc[1] / k);
    pk.y = pack_bf16(sc[2] / k, sc[3] / k);
    pk.z = pack_bf16(sc[4] / k, sc[5] / k);
    pk.w = pack_bf16(sc[6] / k, sc[7] / k);
    *reinterpret_cast<uint4*>(base + (size_t)t * ld) = pk;
  }
}

// ---------------------------------------------------------------------------- dropout (training mode)
// Reference sites: BertEmbeddings / img_embedding dropout (a_transformers 209, transfomres 681), BertSelfOutput and
// BertOutput (a_transformers 371, 449: LayerNorm(dropout(dense(x)) + residual)), the heads' hidden dropouts.
//   out16 = in16 * m          (bf16, may alias)      io32 = io32 * m (+ res32)     m = keep / (1 - p), same mask for both
__global__ void __launch_bounds__(256) dropout_kernel(const bf16* __restrict__ in16, bf16* __restrict__ out16,
                                                      float* __restrict__ io32, const float* __restrict__ res32, int64_t nvec,
                                                      const uint32_t* __restrict__ seed_ptr, uint32_t site, uint32_t thr,
                                                      float inv_keep) {
  const uint32_t seed = __ldg(seed_ptr);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = drop_keep(seed, site, (uint32_t)(v * 8 + j), thr) ? inv_keep : 0.f;
    if (in16 != nullptr) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(in16) + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= m[j];
      reinterpret_cast<uint4*>(out16)[v] = pack8(f);
    }
    if (io32 != nullptr) {
      float4 a = reinterpret_cast<float4*>(io32)[2 * v], b = reinterpret_cast<float4*>(io32)[2 * v + 1];
      a.x *= m[0]; a.y *= m[1]; a.z *= m[2]; a.w *= m[3];
      b.x *= m[4]; b.y *= m[5]; b.z *= m[6]; b.w *= m[7];
      if (res32 != nullptr) {
        const float4 ra = __ldg(reinterpret_cast<const float4*>(res32) + 2 * v);
        const float4 rb = __ldg(reinterpret_cast<const float4*>(res32) + 2 * v + 1);
        a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
        b.x += rb.x; b.y += rb.y; b.z += rb.z; b.w += rb.w;
      }
      reinterpret_cast<float4*>(io32)[2 * v] = a;
      reinterpret_cast<float4*>(io32)[2 * v + 1] = b;
    }
  }
}

// ============================================================================ host launchers
static inline int blocks_for(int64_t work, int per_block, int cap = 148 * 16) {
  int64_t b = (work + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

static int check_h(int H, const char* what) {
  if (H <= 0 || H % 256 != 0 || H > 256 * LN_MAXV) {
    set_error("%s: hidden size %d must be a multiple of 256 and <= %d", what, H, 256 * LN_MAXV);
    return MODCR_ERR_SHAPE;
  }
  return MODCR_OK;
}

int layernorm_fwd(const void* x, int x_f32, RowMap xmap, const float* gamma, const float* beta, void* y, float* y32,
                  RowMap ymap, float* mean, float* rstd, int M, int H, float eps, cudaStream_t st) {
  if (int rc = check_h(H, "layernorm_fwd")) return rc;
  if (M <= 0) return MODCR_OK;
  if (x_f32)
    ln_fwd_kernel<true><<<(M + 7) / 8, 256, 0, st>>>(x, xmap, gamma, beta, (bf16*)y, y32, ymap, mean, rstd, M, H, eps);
  else
    ln_fwd_kernel<false><<<(M + 7) / 8, 256, 0, st>>>(x, xmap, gamma, beta, (bf16*)y, y32, ymap, mean, rstd, M, H, eps);
  return check_launch("layernorm_fwd");
}

int layernorm_bwd(const void* dy, RowMap dymap, const void* x, int x_f32, RowMap xmap, const float* mean,
                  const float* rstd, const float* gamma, void* dx, RowMap dxmap, float* dgamma, float* dbeta, float* dbias,
                  int M, int H, void* dxm, const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st) {
  if (int rc = check_h(H, "layernorm_bwd")) return rc;
  if (M <= 0) return MODCR_OK;
  uint32_t dthr = 0;
  float dinv = 1.f;
  if (dxm != nullptr) {
    if (dseed == nullptr || !(dp > 0.f && dp < 1.f) || (double)M * H >= 4294967296.0) {
      set_error("layernorm_bwd: the masked output needs a device seed pointer, 0 < p < 1 and M*H < 2^32");
      return MODCR_ERR_SHAPE;
    }
    dthr = (uint32_t)fmin(4294967295.0, (double)dp * 4294967296.0);
    if (dthr == 0) dthr = 1;
    dinv = 1.0f / (1.0f - dp);
  }
  const int grid = blocks_for(M, LNB_WARPS, 148);
  const int nv = H >> 8;
#define LNB_LAUNCH(F32, NV)                                                                                          \
  ln_bwd_kernel<F32, NV><<<grid, LNB_WARPS * 32, 0, st>>>((const bf16*)dy, dymap, x, xmap, mean, rstd, gamma, (bf16*)dx, \
                                                          dxmap, dgamma, dbeta, dbias, M, H, (bf16*)dxm, dseed, dsite, dthr, dinv)
  if (x_f32) {
    if (nv == 1) LNB_LAUNCH(true, 1); else if (nv == 2) LNB_LAUNCH(true, 2); else if (nv == 3) LNB_LAUNCH(true, 3); else LNB_LAUNCH(true, 4);
  } else {
    if (nv == 1) LNB_LAUNCH(false, 1); else if (nv == 2) LNB_LAUNCH(false, 2); else if (nv == 3) LNB_LAUNCH(false, 3); else LNB_LAUNCH(false, 4);
  }
#undef LNB_LAUNCH
  return check_launch("layernorm_bwd");
}

int txt_embed_ln_fwd(const int64_t* ids, const int64_t* tts, const int64_t* pos_ids, const float* word, const float* type,
                     const float* pos, const float* gamma, const float* beta, void* y, float* y32, RowMap ymap,
                     float* mean, float* rstd, int N, int T, int H, float eps, cudaStream_t st) {
  if (int rc = check_h(H, "txt_embed_ln_fwd")) return rc;
  const int M = N * T;
  if (M <= 0) return MODCR_OK;
  txt_embed_fwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(ids, tts, pos_ids, word, type, pos, gamma, beta, (bf16*)y, y32, ymap,
                                                    mean, rstd, M, T, H, eps);
  return check_launch("txt_embed_ln_fwd");
}

int txt_embed_ln_bwd(const void* dy, RowMap dymap, const int64_t* ids, const int64_t* tts, const int64_t* pos_ids,
                     const float* word, const float* type, const float* pos, const float* gamma, const float* mean,
                     const float* rstd, float* dword, float* dtype, float* dpos, float* dgamma, float* dbeta, int N, int T,
                     int H, int64_t padding_idx, cudaStream_t st) {
  if (int rc = check_h(H, "txt_embed_ln_bwd")) return rc;
  const int M = N * T;
  if (M <= 0) return MODCR_OK;
  const int grid = blocks_for(M, 8 * 4, 148 * 2);
  txt_embed_bwd_kernel<<<grid, 256, 0, st>>>((const bf16*)dy, dymap, ids, tts, pos_ids, word, type, pos, gamma, mean, rstd,
                                             dword, dtype, dpos, dgamma, dbeta, M, T, H, padding_idx);
  return check_launch("txt_embed_ln_bwd");
}

int cast_pad(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols, int64_t row_step,
             cudaStream_t st) {
  if (ld_dst % 8 != 0 || cols > ld_dst || rows < 0) {
    set_error("cast_pad: ld_dst=%lld must be a multiple of 8 and >= cols=%d", (long long)ld_dst, cols);
    return MODCR_ERR_SHAPE;
  }
  if (rows == 0) return MODCR_OK;
  cast_pad_kernel<<<blocks_for(rows * (ld_dst / 8), 256), 256, 0, st>>>(src, ld_src, (bf16*)dst, ld_dst, rows, cols,
                                                                         row_step);
  return check_launch("cast_pad");
}

int rows_place(const float* src, void* y, float* y32, RowMap ymap, int M, int H, cudaStream_t st) {
  if (M <= 0 || H <= 0 || H % 8 != 0) {
    set_error("rows_place: bad shape M=%d H=%d", M, H);
    return MODCR_ERR_SHAPE;
  }
  rows_place_kernel<<<blocks_for((int64_t)M * (H / 8), 256), 256, 0, st>>>(src, (bf16*)y, y32, ymap, M, H / 8);
  return check_launch("rows_place");
}

int rows_gather(const void* dy, RowMap dymap, void* dst, int M, int H, cudaStream_t st) {
  if (M <= 0 || H <= 0 || H % 8 != 0) {
    set_error("rows_gather: bad shape M=%d H=%d", M, H);
    return MODCR_ERR_SHAPE;
  }
  rows_gather_kernel<<<blocks_for((int64_t)M * (H / 8), 256), 256, 0, st>>>((const bf16*)dy, dymap, (bf16*)dst, M, H / 8);
  return check_launch("rows_gather");
}

int bf16_to_f32(const void* src, float* dst, int64_t n, int accumulate, cudaStream_t st) {
  if (n % 8 != 0) {
    set_error("bf16_to_f32: n=%lld must be a multiple of 8", (long long)n);
    return MODCR_ERR_SHAPE;
  }
  if (n == 0) return MODCR_OK;
  bf16_to_f32_kernel<<<blocks_for(n / 8, 256), 256, 0, st>>>((const bf16*)src, dst, n / 8, accumulate);
  return check_launch("bf16_to_f32");
}

int accum_f32_2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int cols, int accumulate,
                 cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return MODCR_OK;
  accum_f32_2d_kernel<<<blocks_for(rows * cols, 256), 256, 0, st>>>(dst, ld_dst, src, ld_src, rows, cols, accumulate);
  return check_launch("accum_f32_2d");
}

int colsum(const void* x, int64_t ld, int M, int N, float* out, cudaStream_t st) {
  if (N % 8 != 0 || ld % 8 != 0) {
    set_error("colsum: N=%d and ld=%lld must be multiples of 8", N, (long long)ld);
    return MODCR_ERR_SHAPE;
  }
  if (M <= 0) return MODCR_OK;
  dim3 grid((N + 255) / 256, blocks_for(M, 8 * 16, 148));
  colsum_kernel<<<grid, 256, 0, st>>>((const bf16*)x, ld, M, N, out);
  return check_launch("colsum");
}

int copy_rows(const void* src, int64_t src_seq_stride, int srow0, void* dst, int64_t dst_seq_stride, int drow0, int nseq,
              int nrows, int H, int accumulate, cudaStream_t st) {
  if (H % 8 != 0) {
    set_error("copy_rows: H=%d must be a multiple of 8", H);
    return MODCR_ERR_SHAPE;
  }
  if (nseq <= 0 || nrows <= 0) return MODCR_OK;
  copy_rows_kernel<<<blocks_for((int64_t)nseq * nrows * (H / 8), 256), 256, 0, st>>>(
      (const bf16*)src, src_seq_stride, srow0, (bf16*)dst, dst_seq_stride, drow0, nseq, nrows, H, accumulate);
  return check_launch("copy_rows");
}

int add_bf16(const void* a, const void* b, void* y, int64_t n, cudaStream_t st) {
  if (n % 8 != 0) {
    set_error("add_bf16: n=%lld must be a multiple of 8", (long long)n);
    return MODCR_ERR_SHAPE;
  }
  if (n == 0) return MODCR_OK;
  add_kernel<<<blocks_for(n / 8, 256), 256, 0, st>>>((const bf16*)a, (const bf16*)b, (bf16*)y, n / 8);
  return check_launch("add_bf16");
}

int relu_bwd(const void* dy, const void* y, void* dx, int64_t n, cudaStream_t st) {
  if (n <= 0 || n % 8 != 0) {
    set_error("relu_bwd: n=%lld must be a positive multiple of 8", (long long)n);
    return MODCR_ERR_SHAPE;
  }
  relu_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, st>>>((const bf16*)dy, (const bf16*)y, (bf16*)dx, n / 8);
  return check_launch("relu_bwd");
}

int rowln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int H,
              float eps, cudaStream_t st) {
  if (M <= 0 || H <= 0 || H % 8 != 0) {
    set_error("rowln_fwd: bad shape M=%d H=%d (H must be a multiple of 8)", M, H);
    return MODCR_ERR_SHAPE;
  }
  rowln_fwd_kernel<<<M, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, H, eps);
  return check_launch("rowln_fwd");
}

int rowln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
              float* dgamma, float* dbeta, int M, int H, cudaStream_t st) {
  if (M <= 0 || H <= 0 || H % 8 != 0) {
    set_error("rowln_bwd: bad shape M=%d H=%d (H must be a multiple of 8)", M, H);
    return MODCR_ERR_SHAPE;
  }
  rowln_bwd_kernel<<<M, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, mean, rstd, gamma, (bf16*)dx, dgamma, dbeta, H);
  return check_launch("rowln_bwd");
}

int tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, cudaStream_t st) {
  if (n % 8 != 0) {
    set_error("tanh_bwd: n=%lld must be a multiple of 8", (long long)n);
    return MODCR_ERR_SHAPE;
  }
  if (n == 0) return MODCR_OK;
  tanh_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, st>>>((const bf16*)dy, (const bf16*)y, (bf16*)dx, n / 8);
  return check_launch("tanh_bwd");
}

int key_bias(const float* mask, float* kbias, int64_t total, cudaStream_t st) {
  if (total <= 0) return MODCR_OK;
  key_bias_kernel<<<blocks_for(total, 256), 256, 0, st>>>(mask, kbias, total);
  return check_launch("key_bias");
}

int chunk_bits(const float* cmask, uint32_t* bits, int N, int T, int cwords, cudaStream_t st) {
  if (cwords * 32 < T) {
    set_error("chunk_bits: cwords=%d too small for T=%d", cwords, T);
    return MODCR_ERR_SHAPE;
  }
  if (N <= 0) return MODCR_OK;
  chunk_bits_kernel<<<blocks_for((int64_t)N * T * cwords, 256), 256, 0, st>>>(cmask, bits, N, T, cwords);
  return check_launch("chunk_bits");
}

int chunk_mean_q(void* q, int64_t ld, int N, int S, int H, const int32_t* chunk_id, int T, int max_chunks,
                 cudaStream_t st) {
  if (H % 64 != 0 || ld % 8 != 0 || (reinterpret_cast<uintptr_t>(q) & 15) != 0 || T > S || max_chunks <= 0) {
    set_error("chunk_mean_q: bad shape H=%d ld=%lld T=%d S=%d max_chunks=%d (H %% 64, ld %% 8, 16-byte base)", H,
              (long long)ld, T, S, max_chunks);
    return MODCR_ERR_SHAPE;
  }
  const size_t smem = (size_t)max_chunks * 64 * 4 + (size_t)((max_chunks + 3) & ~3) * 4 + (size_t)((T + 3) & ~3) * 4 +
                      (size_t)T * 64 * 2;
  if (smem > 200 * 1024) {
    set_error("chunk_mean_q: %d chunks need %zu B of shared memory", max_chunks, smem);
    return MODCR_ERR_UNSUPPORTED;
  }
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(chunk_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("chunk_mean_q: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return MODCR_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(H / 64, N);
  chunk_mean_kernel<<<grid, 256, smem, st>>>((bf16*)q, ld, S, chunk_id, T, max_chunks);
  return check_launch("chunk_mean_q");
}

int dropout(const void* in16, void* out16, float* io32, const float* res32, int64_t n, const uint32_t* seed, uint32_t site,
            float p, cudaStream_t st) {
  if (n % 8 != 0 || seed == nullptr || !(p >= 0.f && p < 1.f) || (in16 == nullptr) != (out16 == nullptr) ||
      (in16 == nullptr && io32 == nullptr)) {
    set_error("dropout: n=%lld must be a multiple of 8, 0 <= p < 1, a seed pointer and at least one tensor are required",
              (long long)n);
    return MODCR_ERR_SHAPE;
  }
  if (n == 0) return MODCR_OK;
  const uint32_t thr = (uint32_t)fmin(4294967295.0, (double)p * 4294967296.0);
  dropout_kernel<<<blocks_for(n / 8, 256), 256, 0, st>>>((const bf16*)in16, (bf16*)out16, io32, res32, n / 8, seed, site, thr,
                                                         1.0f / (1.0f - p));
  return check_launch("dropout");
}

}  // namespace modcr
