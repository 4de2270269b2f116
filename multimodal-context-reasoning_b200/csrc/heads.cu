// Head kernels of the ModCR hot path (sm_100a): chunk-to-region alignment loss, ClsLayer2
// attention pooling and the fused multiple-choice classifier + cross-entropy.  HBM / latency
// bound: warp-shuffle reductions, 128-bit loads, fp32 math.
#include <cuda_bf16.h>
#include <math.h>

#include "ptx.cuh"
#include "rowops.h"
#include "tma_host.h"

namespace modcr {

typedef __nv_bfloat16 bf16;

DEVINL float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEVINL float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
DEVINL void unpack8h(const uint4& u, float* f) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
DEVINL uint4 pack8h(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}
// block-wide sum / max over 256 threads (8 warps); `sh` needs 8 floats
DEVINL float block_sum(float v, float* sh) {
  v = wsum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) s += sh[i];
  return s;
}
DEVINL float block_max(float v, float* sh) {
  v = wmax(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = -INFINITY;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) s = fmaxf(s, sh[i]);
  return s;
}

// ============================================================================ alignment loss
constexpr int AL_MAXR = 8;  // R <= 256

// probs: [nl, N, heads, T, R] fp32.  One warp per (n, t) row; the nl * heads slices of a row are summed in a fixed order
// (per head: layers, then heads -- the reference's stack(...).sum(1).sum(1), v10:982) into lane-striped registers.
// row_sel (optional, = align_pos): rows without a non-zero entry were never written by the attention kernel and get a
// zero row.  VW = elements per load: 4 (128-bit) when R % 4 == 0, 2 when R is even, else 1.
template <int VW>
__global__ void __launch_bounds__(256) align_fwd_kernel(const float* __restrict__ probs, int nl, int heads,
                                                        const int64_t* __restrict__ align_pos,
                                                        const int64_t* __restrict__ total_label,
                                                        const int64_t* __restrict__ row_sel, float* __restrict__ block,
                                                        float* __restrict__ row_loss, int N, int T, int R) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N * T) return;
  const int n = row / T, t = row - n * T;
  // a[i * VW + e] = column (lane + 32 i) * VW + e
  float a[AL_MAXR];
#pragma unroll
  for (int i = 0; i < AL_MAXR; ++i) a[i] = 0.f;
  const bool wanted = row_sel == nullptr || row_sel[row] != 0;
  if (wanted) {
    for (int h = 0; h < heads; ++h) {
      float tmp[AL_MAXR];
#pragma unroll
      for (int i = 0; i < AL_MAXR; ++i) tmp[i] = 0.f;
      for (int l = 0; l < nl; ++l) {
        const float* src = probs + ((((size_t)l * N + n) * heads + h) * T + t) * R;
#pragma unroll
        for (int i = 0; i < AL_MAXR / VW; ++i) {
          const int r = (lane + 32 * i) * VW;
          if (r < R) {
            if (VW == 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + r));
              tmp[4 * i] += v.x; tmp[4 * i + 1] += v.y; tmp[4 * i + 2] += v.z; tmp[4 * i + 3] += v.w;
            } else if (VW == 2) {
              const float2 v = __ldg(reinterpret_cast<const float2*>(src + r));
              tmp[2 * i] += v.x; tmp[2 * i + 1] += v.y;
            } else {
              tmp[i] += __ldg(src + r);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < AL_MAXR; ++i) a[i] += tmp[i];
    }
  }
  // hand the columns to the lane-strided layout of the softmax below (column lane + 32 i) through the block row itself
#pragma unroll
  for (int i = 0; i < AL_MAXR / VW; ++i) {
    const int r = (lane + 32 * i) * VW;
    if (r < R) {
#pragma unroll
      for (int e = 0; e < VW; ++e) block[(size_t)row * R + r + e] = a[i * VW + e];
    }
  }
  if (VW != 1) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      a[i] = r < R ? block[(size_t)row * R + r] : 0.f;
    }
  }
  float loss = 0.f, correct = 0.f;
  if (align_pos != nullptr && align_pos[row] == 1) {
    // z = softmax(masked_fill(A == 0, -1e5)) ; loss = logsumexp(z) - z[label]
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      if (r < R) {
        a[i] = (a[i] == 0.f) ? -1e5f : a[i];
        mx = fmaxf(mx, a[i]);
      }
    }
    mx = wmax(mx);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      if (r < R) {
        a[i] = expf(a[i] - mx);
        se += a[i];
      }
    }
    se = wsum(se);
    const int label = (int)total_label[row];
    float zmax = -INFINITY, zlab = 0.f, s2 = 0.f;
    int zarg = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      if (r < R) {
        const float z = a[i] / se;
        a[i] = z;
        if (z > zmax) {
          zmax = z;
          zarg = r;
        }
        if (r == label) zlab = z;
      }
    }
    // argmax with first-index tie break
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, zmax, o);
      const int oa = __shfl_xor_sync(0xffffffffu, zarg, o);
      if (om > zmax || (om == zmax && oa < zarg)) {
        zmax = om;
        zarg = oa;
      }
    }
    zlab = wsum(zlab);
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      if (r < R) s2 += expf(a[i] - zmax);
    }
    s2 = wsum(s2);
    loss = (zmax + logf(s2)) - zlab;
    correct = (zarg == label) ? 1.f : 0.f;
  }
  if (lane == 0 && row_loss != nullptr) {
    row_loss[2 * (size_t)row] = loss;
    row_loss[2 * (size_t)row + 1] = correct;
  }
}

// out[0] = mean loss over aligned rows, out[1] = #aligned rows, out[2] = #correct, out[3] = loss sum
__global__ void __launch_bounds__(1024) align_reduce_kernel(const float* __restrict__ row_loss,
                                                            const int64_t* __restrict__ align_pos, float* __restrict__ out,
                                                            int rows) {
  __shared__ float sh[32];
  float s = 0.f, c = 0.f, m = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    if (align_pos[i] == 1) {
      s += row_loss[2 * (size_t)i];
      c += row_loss[2 * (size_t)i + 1];
      m += 1.f;
    }
  }
  s = block_sum(s, sh);
  c = block_sum(c, sh);
  m = block_sum(m, sh);
  if (threadIdx.x == 0) {
    out[0] = s / m;
    out[1] = m;
    out[2] = c;
    out[3] = s;
  }
}

__global__ void __launch_bounds__(256) align_bwd_kernel(const float* __restrict__ block, const int64_t* __restrict__ align_pos,
                                                        const int64_t* __restrict__ total_label, const float* __restrict__ out,
                                                        float gout, float* __restrict__ dblock, int N, int T, int R) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N * T) return;
  float d[AL_MAXR];
#pragma unroll
  for (int i = 0; i < AL_MAXR; ++i) d[i] = 0.f;
  if (align_pos[row] == 1) {
    const float scale = gout / out[1];
    float a[AL_MAXR];
    bool zero[AL_MAXR];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      a[i] = 0.f;
      zero[i] = true;
      if (r < R) {
        const float v = block[(size_t)row * R + r];
        zero[i] = (v == 0.f);
        a[i] = zero[i] ? -1e5f : v;
        mx = fmaxf(mx, a[i]);
      }
    }
    mx = wmax(mx);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i)
      if (lane + 32 * i < R) {
        a[i] = expf(a[i] - mx);
        se += a[i];
      }
    se = wsum(se);
    float zmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i)
      if (lane + 32 * i < R) {
        a[i] /= se;
        zmax = fmaxf(zmax, a[i]);
      }
    zmax = wmax(zmax);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i)
      if (lane + 32 * i < R) s2 += expf(a[i] - zmax);
    s2 = wsum(s2);
    const int label = (int)total_label[row];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) {
      const int r = lane + 32 * i;
      if (r < R) {
        const float y = expf(a[i] - zmax) / s2;
        d[i] = (y - (r == label ? 1.f : 0.f)) * scale;  // dL/dz
        dot += a[i] * d[i];
      }
    }
    dot = wsum(dot);
#pragma unroll
    for (int i = 0; i < AL_MAXR; ++i) d[i] = zero[i] ? 0.f : a[i] * (d[i] - dot);
  }
#pragma unroll
  for (int i = 0; i < AL_MAXR; ++i) {
    const int r = lane + 32 * i;
    if (r < R) dblock[(size_t)row * R + r] = d[i];
  }
}

int align_softmax_ce_fwd(const float* probs, int nl, int heads, const int64_t* align_pos, const int64_t* total_label,
                         const int64_t* row_sel, float* block, float* row_loss, float* out, int N, int T, int R,
                         cudaStream_t st) {
  if (R <= 0 || R > 32 * AL_MAXR || N <= 0 || T <= 0 || nl <= 0) {
    set_error("align_softmax_ce: bad shape N=%d T=%d R=%d (R <= %d)", N, T, R, 32 * AL_MAXR);
    return MODCR_ERR_SHAPE;
  }
  const int rows = N * T;
  const bool al16 = (reinterpret_cast<uintptr_t>(probs) & 15) == 0, al8 = (reinterpret_cast<uintptr_t>(probs) & 7) == 0;
  if ((R & 3) == 0 && al16)
    align_fwd_kernel<4><<<(rows + 7) / 8, 256, 0, st>>>(probs, nl, heads, align_pos, total_label, row_sel, block, row_loss, N, T, R);
  else if ((R & 1) == 0 && al8)
    align_fwd_kernel<2><<<(rows + 7) / 8, 256, 0, st>>>(probs, nl, heads, align_pos, total_label, row_sel, block, row_loss, N, T, R);
  else
    align_fwd_kernel<1><<<(rows + 7) / 8, 256, 0, st>>>(probs, nl, heads, align_pos, total_label, row_sel, block, row_loss, N, T, R);
  if (align_pos != nullptr && row_loss != nullptr && out != nullptr) {
    align_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, align_pos, out, rows);
    count_launches(1);
  }
  return check_launch("align_softmax_ce_fwd");
}

int align_softmax_ce_bwd(const float* block, const int64_t* align_pos, const int64_t* total_label, const float* out,
                         float gout, float* dblock, int N, int T, int R, cudaStream_t st) {
  if (R <= 0 || R > 32 * AL_MAXR || N <= 0 || T <= 0) {
    set_error("align_softmax_ce_bwd: bad shape N=%d T=%d R=%d", N, T, R);
    return MODCR_ERR_SHAPE;
  }
  const int rows = N * T;
  align_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(block, align_pos, total_label, out, gout, dblock, N, T, R);
  return check_launch("align_softmax_ce_bwd");
}

// ============================================================================ ClsLayer2 pooling
constexpr int CP_MAXV = 4;  // H <= 1024

// 8 consecutive elements of a bf16 or fp32 row (q and k are kept in fp32: the scores are unscaled 768-dim dot
// products, the most rounding-sensitive spot of the path)
template <bool F32>
DEVINL void ld8(const void* row, int vec, float* f) {
  if (F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(row) + 2 * vec);
    const float4 b = __ldg(reinterpret_cast<const float4*>(row) + 2 * vec + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack8h(__ldg(reinterpret_cast<const uint4*>(row) + vec), f);
  }
}
template <bool F32>
DEVINL float ld1(const void* base, size_t i) {
  return F32 ? reinterpret_cast<const float*>(base)[i] : __bfloat162float(reinterpret_cast<const bf16*>(base)[i]);
}

constexpr int CP_WARPS = 16;   // one CTA per sequence: 512 threads keep enough 128-bit loads in flight

// Single pass over k (key == value, v10:818-819): every warp walks its rows with a running max / sum / pooled
// accumulator (online softmax), the 16 partial states are merged at the end; the raw scores stay in shared memory
// for the normalised weights `w`.   smem: qs[H] | sc[L] | wm[CP_WARPS] | wsum[CP_WARPS] | part[CP_WARPS][H]
template <bool F32>
__global__ void __launch_bounds__(CP_WARPS * 32) cls_pool_fwd_kernel(const void* __restrict__ q, const void* __restrict__ k,
                                                                     const float* __restrict__ mask, float* __restrict__ w,
                                                                     bf16* __restrict__ pooled, int L, int H,
                                                                     const uint32_t* __restrict__ dseed_ptr, uint32_t dsite,
                                                                     uint32_t dthr, float dinv) {
  extern __shared__ float sm[];
  float* qs = sm;
  float* sc = qs + H;
  float* wm = sc + ((L + 3) & ~3);
  const uint32_t dseed = dthr != 0 ? __ldg(dseed_ptr) : 0u;   // dropout of the attention weights (v10:828)
  float* wsm = wm + CP_WARPS;
  float* part = wsm + CP_WARPS;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = H >> 8;
  constexpr int ES = F32 ? 4 : 2;
  for (int c = threadIdx.x; c < H; c += CP_WARPS * 32) qs[c] = ld1<F32>(q, (size_t)n * H + c);
  __syncthreads();
  const char* kn = reinterpret_cast<const char*>(k) + (size_t)n * L * H * ES;
  float m = -INFINITY, se = 0.f;
  float acc[CP_MAXV][8];
#pragma unroll
  for (int i = 0; i < CP_MAXV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int l = warp; l < L; l += CP_WARPS) {
    float f[CP_MAXV][8];
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < CP_MAXV; ++i)
      if (i < nv) {
        ld8<F32>(kn + (size_t)l * H * ES, lane + 32 * i, f[i]);
        const float* qq = qs + (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) d = fmaf(f[i][j], qq[j], d);
      }
    d = wsum(d) + (mask ? __ldg(mask + (size_t)n * L + l) : 0.f);
    if (lane == 0) sc[l] = d;
    const float mn = fmaxf(m, d);
    const float resc = expf(m - mn);     // 0 on the first row (m = -inf)
    const float pe = expf(d - mn);
    se = fmaf(se, resc, pe);
    m = mn;
    // the pooled vector uses the dropped weight, the normaliser the undropped one
    const float ped = (dthr != 0 && !drop_keep(dseed, dsite, (uint32_t)(n * L + l), dthr)) ? 0.f : pe * dinv;
#pragma unroll
    for (int i = 0; i < CP_MAXV; ++i)
      if (i < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(acc[i][j], resc, ped * f[i][j]);
      }
  }
  if (lane == 0) {
    wm[warp] = m;
    wsm[warp] = se;
  }
#pragma unroll
  for (int i = 0; i < CP_MAXV; ++i)
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) part[warp * H + (lane + 32 * i) * 8 + j] = acc[i][j];
    }
  __syncthreads();
  float mx = -INFINITY;
#pragma unroll
  for (int ww = 0; ww < CP_WARPS; ++ww) mx = fmaxf(mx, wm[ww]);
  float tot = 0.f;
  float wscale[CP_WARPS];
#pragma unroll
  for (int ww = 0; ww < CP_WARPS; ++ww) {
    wscale[ww] = wm[ww] == -INFINITY ? 0.f : expf(wm[ww] - mx);   // warps without rows (L < CP_WARPS)
    tot = fmaf(wsm[ww], wscale[ww], tot);
  }
  const float inv = 1.f / tot;
  for (int l = threadIdx.x; l < L; l += CP_WARPS * 32) w[(size_t)n * L + l] = expf(sc[l] - mx) * inv;
  for (int c = threadIdx.x; c < H; c += CP_WARPS * 32) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < CP_WARPS; ++ww) s = fmaf(part[ww * H + c], wscale[ww], s);
    pooled[(size_t)n * H + c] = __float2bfloat16(s * inv);
  }
}

// smem: qs[H] | dps[H] | wv[L] | dsv[L] | part[CP_WARPS][H]
template <bool F32>
__global__ void __launch_bounds__(CP_WARPS * 32) cls_pool_bwd_kernel(const bf16* __restrict__ dpooled, const void* __restrict__ q,
                                                           const void* __restrict__ k, const float* __restrict__ w,
                                                           bf16* __restrict__ dq, bf16* __restrict__ dk, int L, int H,
                                                           const uint32_t* __restrict__ dseed_ptr, uint32_t dsite, uint32_t dthr,
                                                           float dinv) {
  extern __shared__ float sm[];
  __shared__ float sh[CP_WARPS];
  const uint32_t dseed = dthr != 0 ? __ldg(dseed_ptr) : 0u;
  const int Lp = (L + 3) & ~3;
  float* qs = sm;
  float* dps = qs + H;
  float* wv = dps + H;
  float* dsv = wv + Lp;
  float* part = dsv + Lp;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = H >> 8;
  constexpr int ES = F32 ? 4 : 2;
  for (int c = threadIdx.x; c < H; c += CP_WARPS * 32) {
    qs[c] = ld1<F32>(q, (size_t)n * H + c);
    dps[c] = __bfloat162float(dpooled[(size_t)n * H + c]);
  }
  for (int l = threadIdx.x; l < L; l += CP_WARPS * 32) wv[l] = w[(size_t)n * L + l];
  __syncthreads();
  const char* kn = reinterpret_cast<const char*>(k) + (size_t)n * L * H * ES;
  bf16* dkn = dk + (size_t)n * L * H;
  for (int l = warp; l < L; l += CP_WARPS) {  // dw[l] = dpooled . k[l]
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < CP_MAXV; ++i)
      if (i < nv) {
        float f[8];
        ld8<F32>(kn + (size_t)l * H * ES, lane + 32 * i, f);
        const float* dd = dps + (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) d = fmaf(f[j], dd[j], d);
      }
    d = wsum(d);
    // dropout: pooled = sum_l D_l w_l k_l  =>  d w_l = D_l (dpooled . k_l)
    if (dthr != 0) d = drop_keep(dseed, dsite, (uint32_t)(n * L + l), dthr) ? d * dinv : 0.f;
    if (lane == 0) dsv[l] = d;
  }
  __syncthreads();
  float dot = 0.f;
  for (int l = threadIdx.x; l < L; l += CP_WARPS * 32) dot += wv[l] * dsv[l];
  dot = block_sum(dot, sh);
  for (int l = threadIdx.x; l < L; l += CP_WARPS * 32) dsv[l] = wv[l] * (dsv[l] - dot);
  __syncthreads();
  float acc[CP_MAXV][8];
#pragma unroll
  for (int i = 0; i < CP_MAXV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int l = warp; l < L; l += CP_WARPS) {
    float p = wv[l];
    const float ds = dsv[l];
    if (dthr != 0) p = drop_keep(dseed, dsite, (uint32_t)(n * L + l), dthr) ? p * dinv : 0.f;   // value path: D_l w_l
#pragma unroll
    for (int i = 0; i < CP_MAXV; ++i)
      if (i < nv) {
        float f[8], o[8];
        ld8<F32>(kn + (size_t)l * H * ES, lane + 32 * i, f);
        const float* dd = dps + (lane + 32 * i) * 8;
        const float* qq = qs + (lane + 32 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = p * dd[j] + ds * qq[j];           // dk[l] = w[l] dpooled + ds[l] q   (k is key and value)
          acc[i][j] = fmaf(ds, f[j], acc[i][j]);   // dq += ds[l] k[l]
        }
        reinterpret_cast<uint4*>(dkn + (size_t)l * H)[lane + 32 * i] = pack8h(o);
      }
  }
#pragma unroll
  for (int i = 0; i < CP_MAXV; ++i)
    if (i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) part[warp * H + (lane + 32 * i) * 8 + j] = acc[i][j];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += CP_WARPS * 32) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < CP_WARPS; ++ww) s += part[ww * H + c];
    dq[(size_t)n * H + c] = __float2bfloat16(s);
  }
}

static int cls_pool_check(int N, int L, int H, size_t smem, const char* what) {
  if (N <= 0 || L <= 0 || H % 256 != 0 || H > 256 * CP_MAXV) {
    set_error("%s: bad shape N=%d L=%d H=%d", what, N, L, H);
    return MODCR_ERR_SHAPE;
  }
  if (smem > 200 * 1024) {
    set_error("%s: L=%d needs %zu B of shared memory", what, L, smem);
    return MODCR_ERR_UNSUPPORTED;
  }
  return MODCR_OK;
}

static uint32_t drop_threshold(float p) {
  if (!(p > 0.f)) return 0u;
  const uint32_t t = (uint32_t)fmin(4294967295.0, (double)p * 4294967296.0);
  return t == 0 ? 1u : t;
}

int cls_pool_attn_fwd(const void* q, const void* k, int qk_f32, const float* mask, float* w, void* pooled, int N, int L,
                      int H, const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st) {
  const uint32_t dthr = drop_threshold(dp);
  const float dinv = dthr ? 1.0f / (1.0f - dp) : 1.0f;
  if (dthr && (dseed == nullptr || !(dp < 1.f))) {
    set_error("cls_pool_attn_fwd: dropout needs a device seed pointer and p < 1");
    return MODCR_ERR_SHAPE;
  }
  const size_t smem = ((size_t)H + ((L + 3) & ~3) + 2 * CP_WARPS + CP_WARPS * (size_t)H) * 4;
  if (int rc = cls_pool_check(N, L, H, smem, "cls_pool_attn_fwd")) return rc;
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    cudaFuncSetAttribute(cls_pool_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(cls_pool_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  if (qk_f32)
    cls_pool_fwd_kernel<true><<<N, CP_WARPS * 32, smem, st>>>(q, k, mask, w, (bf16*)pooled, L, H, dseed, dsite, dthr, dinv);
  else
    cls_pool_fwd_kernel<false><<<N, CP_WARPS * 32, smem, st>>>(q, k, mask, w, (bf16*)pooled, L, H, dseed, dsite, dthr, dinv);
  return check_launch("cls_pool_attn_fwd");
}

int cls_pool_attn_bwd(const void* dpooled, const void* q, const void* k, int qk_f32, const float* w, void* dq, void* dk,
                      int N, int L, int H, const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st) {
  const uint32_t dthr = drop_threshold(dp);
  const float dinv = dthr ? 1.0f / (1.0f - dp) : 1.0f;
  if (dthr && (dseed == nullptr || !(dp < 1.f))) {
    set_error("cls_pool_attn_bwd: dropout needs a device seed pointer and p < 1");
    return MODCR_ERR_SHAPE;
  }
  const size_t smem = (2 * (size_t)H + 2 * ((L + 3) & ~3) + CP_WARPS * (size_t)H) * 4;
  if (int rc = cls_pool_check(N, L, H, smem, "cls_pool_attn_bwd")) return rc;
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    cudaFuncSetAttribute(cls_pool_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(cls_pool_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  if (qk_f32)
    cls_pool_bwd_kernel<true><<<N, CP_WARPS * 32, smem, st>>>((const bf16*)dpooled, q, k, w, (bf16*)dq, (bf16*)dk, L, H, dseed,
                                                             dsite, dthr, dinv);
  else
    cls_pool_bwd_kernel<false><<<N, CP_WARPS * 32, smem, st>>>((const bf16*)dpooled, q, k, w, (bf16*)dq, (bf16*)dk, L, H, dseed,
                                                              dsite, dthr, dinv);
  return check_launch("cls_pool_attn_bwd");
}

// ============================================================================ cross_attention_lyx (one query)
// ClsLayer_lyx (v10:856-870) attends with ONE query per sequence (the CLS vector) over L key/value rows with 8 heads
// (cross_attention_lyx, v10:692-797: q scaled by head_dim^-0.5, no mask is ever passed, tau = 1, dropout p = 0).
// One CTA per sequence; every lane owns H/32 contiguous channels, i.e. a fixed head (head_dim / (H/32) lanes per head),
// so a head's score is a short shuffle reduction inside its lane group and the value row is weighted lane-locally.
// Single pass with online softmax per head; the XA_WARPS partial states are merged through shared memory.
// smem: qs[H] | wm[XA_WARPS][32] | ws[XA_WARPS][32] | part[XA_WARPS][H]
constexpr int XA_WARPS = 16;

template <int VEC>   // 16-byte vectors per lane: H = 32 * 8 * VEC
__global__ void __launch_bounds__(XA_WARPS * 32) cross_attn_q1_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                                          const bf16* __restrict__ v, bf16* __restrict__ out,
                                                                          float* __restrict__ lse, int L, int H,
                                                                          int lanes_per_head, float scaling,
                                                                          const uint32_t* __restrict__ dseed_ptr, uint32_t dsite,
                                                                          uint32_t dthr, float dinv) {
  extern __shared__ float sm[];
  float* qs = sm;
  const uint32_t dseed = dthr != 0 ? __ldg(dseed_ptr) : 0u;   // attention-probability dropout (v10:785, p = 0.1)
  const int nheads = 32 / lanes_per_head;
  float* wm = qs + H;
  float* wsm = wm + XA_WARPS * 32;
  float* part = wsm + XA_WARPS * 32;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int E = 8 * VEC;
  for (int c = threadIdx.x; c < H; c += XA_WARPS * 32) qs[c] = __bfloat162float(q[(size_t)n * H + c]) * scaling;
  __syncthreads();
  float ql[E];
#pragma unroll
  for (int j = 0; j < E; ++j) ql[j] = qs[lane * E + j];
  const bf16* kn = k + (size_t)n * L * H + lane * E;
  const bf16* vn = v + (size_t)n * L * H + lane * E;
  float m = -INFINITY, se = 0.f;
  float acc[E];
#pragma unroll
  for (int j = 0; j < E; ++j) acc[j] = 0.f;
  for (int l = warp; l < L; l += XA_WARPS) {
    float kf[E], vf[E];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      unpack8h(__ldg(reinterpret_cast<const uint4*>(kn + (size_t)l * H) + i), kf + 8 * i);
      unpack8h(__ldg(reinterpret_cast<const uint4*>(vn + (size_t)l * H) + i), vf + 8 * i);
    }
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) d = fmaf(kf[j], ql[j], d);
    for (int o = 1; o < lanes_per_head; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);   // head score
    const float mn = fmaxf(m, d);
    const float resc = expf(m - mn);     // 0 on the first row (m = -inf)
    const float pe = expf(d - mn);
    se = fmaf(se, resc, pe);
    m = mn;
    const float ped = (dthr != 0 && !drop_keep(dseed, dsite, (uint32_t)((n * nheads + lane / lanes_per_head) * L + l), dthr))
                          ? 0.f : pe * dinv;
#pragma unroll
    for (int j = 0; j < E; ++j) acc[j] = fmaf(acc[j], resc, ped * vf[j]);
  }
  wm[warp * 32 + lane] = m;
  wsm[warp * 32 + lane] = se;
#pragma unroll
  for (int j = 0; j < E; ++j) part[warp * H + lane * E + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += XA_WARPS * 32) {
    const int ln = c / E;   // the lane (hence head) that owns channel c
    float mx = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < XA_WARPS; ++ww) mx = fmaxf(mx, wm[ww * 32 + ln]);
    float tot = 0.f, s = 0.f;
#pragma unroll
    for (int ww = 0; ww < XA_WARPS; ++ww) {
      const float mw = wm[ww * 32 + ln];
      const float sc = mw == -INFINITY ? 0.f : expf(mw - mx);   // warps without rows (L < XA_WARPS)
      tot = fmaf(wsm[ww * 32 + ln], sc, tot);
      s = fmaf(part[ww * H + c], sc, s);
    }
    out[(size_t)n * H + c] = __float2bfloat16(s / tot);
    // log-sum-exp of the head's (scaled) scores, one value per lane group, for the backward pass
    if (lse != nullptr && (c % (E * lanes_per_head)) == 0) lse[(size_t)n * (32 / lanes_per_head) + ln / lanes_per_head] = mx + logf(tot);
  }
}

// Backward of the above in one pass over the rows: with p_l = exp(s_l - lse), delta = dout . out (per head),
//   dv_l = p_l dout,  ds_l = p_l (dout . v_l - delta),  dk_l = ds_l (scaling q),  dq = scaling sum_l ds_l k_l.
// smem: part[XA_WARPS][H]
template <int VEC>
__global__ void __launch_bounds__(XA_WARPS * 32) cross_attn_q1_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                                          const bf16* __restrict__ v, const bf16* __restrict__ out,
                                                                          const float* __restrict__ lse, const bf16* __restrict__ dout,
                                                                          bf16* __restrict__ dq, bf16* __restrict__ dk,
                                                                          bf16* __restrict__ dv, int L, int H, int lanes_per_head,
                                                                          float scaling, const uint32_t* __restrict__ dseed_ptr,
                                                                          uint32_t dsite, uint32_t dthr, float dinv) {
  extern __shared__ float sm[];
  float* part = sm;
  const uint32_t dseed = dthr != 0 ? __ldg(dseed_ptr) : 0u;
  const int nheads = 32 / lanes_per_head;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int E = 8 * VEC;
  float ql[E], dol[E];
  float delta = 0.f;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    ql[j] = __bfloat162float(q[(size_t)n * H + lane * E + j]) * scaling;
    dol[j] = __bfloat162float(dout[(size_t)n * H + lane * E + j]);
    delta = fmaf(dol[j], __bfloat162float(out[(size_t)n * H + lane * E + j]), delta);
  }
  for (int o = 1; o < lanes_per_head; o <<= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
  const float lse_h = lse[(size_t)n * (32 / lanes_per_head) + lane / lanes_per_head];
  const size_t rowbase = (size_t)n * L * H + lane * E;
  float acc[E];
#pragma unroll
  for (int j = 0; j < E; ++j) acc[j] = 0.f;
  for (int l = warp; l < L; l += XA_WARPS) {
    float kf[E], vf[E];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      unpack8h(__ldg(reinterpret_cast<const uint4*>(k + rowbase + (size_t)l * H) + i), kf + 8 * i);
      unpack8h(__ldg(reinterpret_cast<const uint4*>(v + rowbase + (size_t)l * H) + i), vf + 8 * i);
    }
    float sdot = 0.f, dp = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      sdot = fmaf(kf[j], ql[j], sdot);
      dp = fmaf(vf[j], dol[j], dp);
    }
    for (int o = 1; o < lanes_per_head; o <<= 1) {
      sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
      dp += __shfl_xor_sync(0xffffffffu, dp, o);
    }
    const float pr = expf(sdot - lse_h);
    float prd = pr;            // value path uses D o P, the score gradient D o dP (delta = dout . out is unchanged)
    if (dthr != 0) {
      const float dm = drop_keep(dseed, dsite, (uint32_t)((n * nheads + lane / lanes_per_head) * L + l), dthr) ? dinv : 0.f;
      prd *= dm;
      dp *= dm;
    }
    const float ds = pr * (dp - delta);
    float okk[E], ov[E];
#pragma unroll
    for (int j = 0; j < E; ++j) {
      okk[j] = ds * ql[j];
      ov[j] = prd * dol[j];
      acc[j] = fmaf(ds, kf[j], acc[j]);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      reinterpret_cast<uint4*>(dk + rowbase + (size_t)l * H)[i] = pack8h(okk + 8 * i);
      reinterpret_cast<uint4*>(dv + rowbase + (size_t)l * H)[i] = pack8h(ov + 8 * i);
    }
  }
#pragma unroll
  for (int j = 0; j < E; ++j) part[warp * H + lane * E + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += XA_WARPS * 32) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < XA_WARPS; ++ww) s += part[ww * H + c];
    dq[(size_t)n * H + c] = __float2bfloat16(s * scaling);
  }
}

static int xa_check(int N, int L, int H, int heads, const char* what) {
  const int E = H / 32;
  const int hd = heads > 0 ? H / heads : 0;
  if (N <= 0 || L <= 0 || heads <= 0 || H % (32 * 8) != 0 || H % heads != 0 || (E != 24 && E != 32) || hd % E != 0 ||
      ((hd / E) & (hd / E - 1)) != 0) {
    set_error("%s: bad shape N=%d L=%d H=%d heads=%d (H in {768, 1024}, head_dim a power-of-two multiple of H/32)", what,
              N, L, H, heads);
    return MODCR_ERR_SHAPE;
  }
  return MODCR_OK;
}

int cross_attn_q1_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout,
                      void* dq, void* dk, void* dv, int N, int L, int H, int heads, const uint32_t* dseed, uint32_t dsite,
                      float dp, cudaStream_t st) {
  if (int rc = xa_check(N, L, H, heads, "cross_attn_q1_bwd")) return rc;
  const uint32_t dthr = drop_threshold(dp);
  const float dinv = dthr ? 1.0f / (1.0f - dp) : 1.0f;
  if (dthr && (dseed == nullptr || !(dp < 1.f))) {
    set_error("cross_attn_q1_bwd: dropout needs a device seed pointer and p < 1");
    return MODCR_ERR_SHAPE;
  }
  if (lse == nullptr) {
    set_error("cross_attn_q1_bwd: the forward pass must have written lse");
    return MODCR_ERR_SHAPE;
  }
  const int E = H / 32, hd = H / heads;
  const size_t smem = (size_t)XA_WARPS * H * 4;
  const float scaling = 1.0f / sqrtf((float)hd);
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    cudaFuncSetAttribute(cross_attn_q1_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(cross_attn_q1_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  if (E == 24)
    cross_attn_q1_bwd_kernel<3><<<N, XA_WARPS * 32, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)out,
                                                               lse, (const bf16*)dout, (bf16*)dq, (bf16*)dk, (bf16*)dv, L, H,
                                                               hd / E, scaling, dseed, dsite, dthr, dinv);
  else
    cross_attn_q1_bwd_kernel<4><<<N, XA_WARPS * 32, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)out,
                                                               lse, (const bf16*)dout, (bf16*)dq, (bf16*)dk, (bf16*)dv, L, H,
                                                               hd / E, scaling, dseed, dsite, dthr, dinv);
  return check_launch("cross_attn_q1_bwd");
}

int cross_attn_q1_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int N, int L, int H, int heads,
                      const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st) {
  const uint32_t dthr = drop_threshold(dp);
  const float dinv = dthr ? 1.0f / (1.0f - dp) : 1.0f;
  if (dthr && (dseed == nullptr || !(dp < 1.f))) {
    set_error("cross_attn_q1_fwd: dropout needs a device seed pointer and p < 1");
    return MODCR_ERR_SHAPE;
  }
  const int E = H / 32;
  const int hd = heads > 0 ? H / heads : 0;
  if (N <= 0 || L <= 0 || heads <= 0 || H % (32 * 8) != 0 || H % heads != 0 || (E != 24 && E != 32) || hd % E != 0 ||
      ((hd / E) & (hd / E - 1)) != 0) {
    set_error("cross_attn_q1_fwd: bad shape N=%d L=%d H=%d heads=%d (H in {768, 1024}, head_dim a power-of-two multiple of H/32)",
              N, L, H, heads);
    return MODCR_ERR_SHAPE;
  }
  const size_t smem = ((size_t)H + 2 * XA_WARPS * 32 + (size_t)XA_WARPS * H) * 4;
  const float scaling = 1.0f / sqrtf((float)hd);
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    cudaFuncSetAttribute(cross_attn_q1_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(cross_attn_q1_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  if (E == 24)
    cross_attn_q1_fwd_kernel<3><<<N, XA_WARPS * 32, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out, lse,
                                                               L, H, hd / E, scaling, dseed, dsite, dthr, dinv);
  else
    cross_attn_q1_fwd_kernel<4><<<N, XA_WARPS * 32, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out, lse,
                                                               L, H, hd / E, scaling, dseed, dsite, dthr, dinv);
  return check_launch("cross_attn_q1_fwd");
}

// ============================================================================ classifier + CE
constexpr int SC_MAXC = 8;

struct ScorerParams {
  const bf16* x;
  const float* W;
  const float* b;
  int rows, K, C, G, V;
  const int64_t* hard;
  const float* soft;
  float* logits;
  float* loss;
  float* mp;
  int64_t* pred;
  uint8_t* matched;
  int num_choice;
  float gout;
  bf16* dx;
  float* dW;
  float* db;
  int loss_kind;  // 0 cross-entropy over the view [V,G]; 1 instance_bce_with_logits (modeling_bert.py:416-421)
};

DEVINL void scorer_logits(const ScorerParams& p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < p.rows; r += nw) {
    float acc[SC_MAXC];
#pragma unroll
    for (int c = 0; c < SC_MAXC; ++c) acc[c] = 0.f;
    for (int v = lane; v < (p.K >> 3); v += 32) {
      float f[8];
      unpack8h(__ldg(reinterpret_cast<const uint4*>(p.x + (size_t)r * p.K) + v), f);
#pragma unroll
      for (int c = 0; c < SC_MAXC; ++c)
        if (c < p.C) {
          const float* wr = p.W + (size_t)c * p.K + v * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c] = fmaf(f[j], __ldg(wr + j), acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < SC_MAXC; ++c)
      if (c < p.C) {
        const float s = wsum(acc[c]);
        if (lane == 0) p.logits[(size_t)r * p.C + c] = s + (p.b ? p.b[c] : 0.f);
      }
  }
}

__global__ void __launch_bounds__(256) scorer_fwd_kernel(const ScorerParams p) {
  __shared__ float sh[8];
  scorer_logits(p);
  __syncthreads();
  float ls = 0.f;
  for (int v = threadIdx.x; v < p.V; v += blockDim.x) {
    const float* lg = p.logits + (size_t)v * p.G;
    if (p.loss_kind == 1) {
      // mean over all V*G elements of softplus(l) - t*l, times labels.size(1) = G  ==  sum / V
      for (int g = 0; g < p.G; ++g) {
        const float l = lg[g], t = p.soft[(size_t)v * p.G + g];
        ls += fmaxf(l, 0.f) - l * t + log1pf(expf(-fabsf(l)));
      }
      continue;
    }
    float mx = -INFINITY;
    for (int g = 0; g < p.G; ++g) mx = fmaxf(mx, lg[g]);
    float se = 0.f;
    for (int g = 0; g < p.G; ++g) se += expf(lg[g] - mx);
    const float lse = mx + logf(se);
    if (p.hard) {
      ls += lse - lg[(int)p.hard[v]];
    } else if (p.soft) {
      for (int g = 0; g < p.G; ++g) ls += p.soft[(size_t)v * p.G + g] * (lse - lg[g]);
    }
  }
  ls = block_sum(ls, sh);
  if (threadIdx.x == 0 && p.loss) p.loss[0] = ls / (float)p.V;
  if (p.mp != nullptr && p.C == 2) {  // binary_to_mp (v10:363-373) + argmax + matched (v10:1065-1069)
    const int nq = p.rows / p.num_choice;
    for (int b = threadIdx.x; b < nq; b += blockDim.x) {
      float best = -INFINITY, gbest = -INFINITY;
      int arg = 0, garg = 0;
      for (int j = 0; j < p.num_choice; ++j) {
        const int r = b * p.num_choice + j;
        const float l0 = p.logits[2 * (size_t)r], l1 = p.logits[2 * (size_t)r + 1];
        const float mx = fmaxf(l0, l1);
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        const float pr = e1 / (e0 + e1);
        p.mp[r] = pr;
        if (pr > best) {
          best = pr;
          arg = j;
        }
        const float lab = p.hard ? (float)p.hard[r] : (p.soft ? p.soft[2 * (size_t)r + 1] : 0.f);
        if (lab > gbest) {
          gbest = lab;
          garg = j;
        }
      }
      if (p.pred) p.pred[b] = arg;
      if (p.matched) p.matched[b] = (arg == garg) ? 1 : 0;
    }
  }
}

// dynamic smem: dl[rows * C]
__global__ void __launch_bounds__(256) scorer_bwd_kernel(const ScorerParams p) {
  extern __shared__ float dl[];
  const float scale = p.gout / (float)p.V;
  for (int v = threadIdx.x; v < p.V; v += blockDim.x) {
    const float* lg = p.logits + (size_t)v * p.G;
    if (p.loss_kind == 1) {  // d/dl [softplus(l) - t*l] = sigmoid(l) - t
      for (int g = 0; g < p.G; ++g)
        dl[(size_t)v * p.G + g] = (1.f / (1.f + expf(-lg[g])) - p.soft[(size_t)v * p.G + g]) * scale;
      continue;
    }
    float mx = -INFINITY;
    for (int g = 0; g < p.G; ++g) mx = fmaxf(mx, lg[g]);
    float se = 0.f;
    for (int g = 0; g < p.G; ++g) se += expf(lg[g] - mx);
    float tsum = 1.f;
    if (p.soft) {
      tsum = 0.f;
      for (int g = 0; g < p.G; ++g) tsum += p.soft[(size_t)v * p.G + g];
    }
    for (int g = 0; g < p.G; ++g) {
      const float y = expf(lg[g] - mx) / se;
      const float tg = p.hard ? ((int)p.hard[v] == g ? 1.f : 0.f) : p.soft[(size_t)v * p.G + g];
      dl[(size_t)v * p.G + g] = (y * tsum - tg) * scale;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // every CTA recomputes the (tiny) dlogits above; the rows of dx and the elements of dW are split over the grid
  if (p.dx) {
    for (int r = blockIdx.x * nw + warp; r < p.rows; r += gridDim.x * nw) {
      for (int v = lane; v < (p.K >> 3); v += 32) {
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < p.C; ++c) {
          const float d = dl[(size_t)r * p.C + c];
          const float* wr = p.W + (size_t)c * p.K + v * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(d, __ldg(wr + j), o[j]);
        }
        reinterpret_cast<uint4*>(p.dx + (size_t)r * p.K)[v] = pack8h(o);
      }
    }
  }
  if (p.dW) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.C * p.K; i += gridDim.x * blockDim.x) {
      const int c = i / p.K, kk = i - c * p.K;
      float s = 0.f;
      for (int r = 0; r < p.rows; ++r) s = fmaf(dl[(size_t)r * p.C + c], __bfloat162float(p.x[(size_t)r * p.K + kk]), s);
      p.dW[i] += s;
    }
  }
  if (p.db && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < p.rows; ++r) s += dl[(size_t)r * p.C + c];
      p.db[c] += s;
    }
  }
}

static int scorer_params(const ScorerDesc& d, ScorerParams& p, const char* what) {
  if (d.rows <= 0 || d.K <= 0 || d.K % 8 != 0 || d.C <= 0 || d.C > SC_MAXC || d.G <= 0 || (d.rows * d.C) % d.G != 0) {
    set_error("%s: bad shape rows=%d K=%d C=%d G=%d", what, d.rows, d.K, d.C, d.G);
    return MODCR_ERR_SHAPE;
  }
  if (d.x == nullptr || d.W == nullptr || d.logits == nullptr || (d.hard == nullptr && d.soft == nullptr)) {
    set_error("%s: null pointer", what);
    return MODCR_ERR_SHAPE;
  }
  if (d.mp != nullptr && (d.C != 2 || d.G != 2 || d.num_choice <= 0 || d.rows % d.num_choice != 0)) {
    set_error("%s: binary_to_mp needs C=G=2 and rows %% num_choice == 0", what);
    return MODCR_ERR_SHAPE;
  }
  p.x = (const bf16*)d.x; p.W = d.W; p.b = d.b; p.rows = d.rows; p.K = d.K; p.C = d.C; p.G = d.G;
  p.V = d.rows * d.C / d.G; p.hard = d.hard; p.soft = d.soft; p.logits = d.logits; p.loss = d.loss; p.mp = d.mp;
  p.pred = d.pred; p.matched = d.matched; p.num_choice = d.num_choice; p.gout = d.gout; p.dx = (bf16*)d.dx; p.dW = d.dW;
  p.db = d.db;
  p.loss_kind = d.loss_kind;
  if (d.loss_kind != 0 && (d.loss_kind != 1 || d.soft == nullptr)) {
    set_error("%s: loss_kind %d needs fp32 targets (1 = BCE with logits) or must be 0", what, d.loss_kind);
    return MODCR_ERR_SHAPE;
  }
  return MODCR_OK;
}

int mc_scorer_ce_fwd(const ScorerDesc& d, cudaStream_t st) {
  ScorerParams p;
  if (int rc = scorer_params(d, p, "mc_scorer_ce_fwd")) return rc;
  scorer_fwd_kernel<<<1, 256, 0, st>>>(p);
  return check_launch("mc_scorer_ce_fwd");
}

int mc_scorer_ce_bwd(const ScorerDesc& d, cudaStream_t st) {
  ScorerParams p;
  if (int rc = scorer_params(d, p, "mc_scorer_ce_bwd")) return rc;
  const size_t smem = (size_t)d.rows * d.C * 4;
  if (smem > 48 * 1024) {
    set_error("mc_scorer_ce_bwd: rows*C=%d too large", d.rows * d.C);
    return MODCR_ERR_UNSUPPORTED;
  }
  int grid = (d.rows + 7) / 8;
  if (grid > 64) grid = 64;
  if (grid < 1) grid = 1;
  scorer_bwd_kernel<<<grid, 256, smem, st>>>(p);
  return check_launch("mc_scorer_ce_bwd");
}

}  // namespace modcr
