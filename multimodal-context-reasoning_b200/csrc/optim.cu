// Fused multi-tensor gradient-norm clip + AdamW for the training driver.
//
// Replaces, per optimisation step of run_PMR_ModCR.py:211-227 / run_vcr_ModCR.py,
//     torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)      (one reduction + one scale per tensor)
//     transformers.AdamW.step()   (two LR groups: seq_enc x 0.1; eps 1e-5; bias correction; weight_decay 0)
// -- ~485 tensors x ~10 tiny launches in the reference -- by two launches over a (tensor, chunk) block table:
//   1. grad_sqnorm_kernel : sum of squares of every gradient -> one fp32 scalar
//   2. adamw_kernel       : g' = g * min(1, max_norm / (norm + 1e-6)),  m, v, p update in one pass (7 x 4 B / element),
//                           plus the refreshed kernel-operand copy of the parameter (bf16 weight / packed fp32 bias,
//                           `shadow`), so the next forward pass does not re-cast ~100 weight matrices
//   3. grad_clip_kernel   : the in-place clip of the micro-steps that do not end in optimizer.step()
//                           (gradient accumulation, run_PMR_ModCR.py:215: clip_grad_norm_ after EVERY backward)
// The update is the HuggingFace AdamW of the reference (eps added to sqrt(v) before the bias correction):
//     m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2 ; p -= lr_t * m / (sqrt(v) + eps) ; p -= lr * wd * p
//     lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
// Step-dependent scalars (lr, the two bias corrections, max_norm) are read from device memory so that the step can
// sit inside a replayed CUDA graph.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "rowops.h"
#include "tma_host.h"

namespace modcr {

constexpr int OPT_CHUNK = 8192;   // elements per block
constexpr int OPT_THREADS = 256;
constexpr int MULTI_MAX = 64;     // tensors per multi_copy launch (MODCR_MULTI_MAX)

struct OptTensor {   // mirrors modcr_optim_tensor (include/modcr_b200.h)
  float* p;
  float* m;
  float* v;
  void* shadow;          // optional operand copy refreshed with the update: bf16 (kind 1) or fp32 (kind 2), same layout
  int64_t n;
  float lr_mul;
  float wd;
  int32_t shadow_kind;
  int32_t reserved;
};

__global__ void __launch_bounds__(OPT_THREADS) grad_sqnorm_kernel(const OptTensor* __restrict__ meta, const float* const* __restrict__ grads,
                                                                  const int32_t* __restrict__ blk_tensor,
                                                                  const int32_t* __restrict__ blk_chunk, float* __restrict__ out) {
  __shared__ float sh[OPT_THREADS / 32];
  const int t = blk_tensor[blockIdx.x];
  const float* g = grads[t];
  float s = 0.f;
  if (g != nullptr) {
    const int64_t n = meta[t].n;
    const int64_t base = (int64_t)blk_chunk[blockIdx.x] * OPT_CHUNK;
    const int64_t end = min(base + OPT_CHUNK, n);
    if ((n & 3) == 0) {
      for (int64_t i = base + threadIdx.x * 4; i < end; i += OPT_THREADS * 4) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(g + i));
        s = fmaf(x.x, x.x, s); s = fmaf(x.y, x.y, s); s = fmaf(x.z, x.z, s); s = fmaf(x.w, x.w, s);
      }
    } else {
      for (int64_t i = base + threadIdx.x; i < end; i += OPT_THREADS) s = fmaf(g[i], g[i], s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < OPT_THREADS / 32; ++i) tot += sh[i];
    atomicAdd(out, tot);
  }
}

// hyper = {lr, 1 - b1^t, 1 - b2^t, max_grad_norm (<= 0: no clipping)}
__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const OptTensor* __restrict__ meta, const float* const* __restrict__ grads,
                                                            const int32_t* __restrict__ blk_tensor,
                                                            const int32_t* __restrict__ blk_chunk, const float* __restrict__ hyper,
                                                            const float* __restrict__ grad_sq, float b1, float b2, float eps) {
  const int t = blk_tensor[blockIdx.x];
  const float* g = grads[t];
  if (g == nullptr) return;   // parameter without a gradient this step: untouched, like torch optimizers
  const OptTensor mt = meta[t];
  const float lr = __ldg(hyper) * mt.lr_mul;
  const float bc1 = __ldg(hyper + 1), bc2 = __ldg(hyper + 2), max_norm = __ldg(hyper + 3);
  float clip = 1.f;
  if (grad_sq != nullptr && max_norm > 0.f) clip = fminf(1.f, max_norm / (sqrtf(__ldg(grad_sq)) + 1e-6f));
  const float step = lr * sqrtf(bc2) / bc1;
  const float decay = 1.f - lr * mt.wd;
  const int64_t base = (int64_t)blk_chunk[blockIdx.x] * OPT_CHUNK;
  const int64_t end = min(base + OPT_CHUNK, mt.n);
  auto upd = [&](float gi, float& m, float& v, float& p) {
    gi *= clip;
    m = fmaf(b1, m, (1.f - b1) * gi);
    v = fmaf(b2, v, (1.f - b2) * gi * gi);
    p = fmaf(-step, m / (sqrtf(v) + eps), p);
    p *= decay;
  };
  if ((mt.n & 3) == 0) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += OPT_THREADS * 4) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
      float4 m = *reinterpret_cast<float4*>(mt.m + i), v = *reinterpret_cast<float4*>(mt.v + i);
      float4 p = *reinterpret_cast<float4*>(mt.p + i);
      upd(gg.x, m.x, v.x, p.x); upd(gg.y, m.y, v.y, p.y); upd(gg.z, m.z, v.z, p.z); upd(gg.w, m.w, v.w, p.w);
      *reinterpret_cast<float4*>(mt.m + i) = m;
      *reinterpret_cast<float4*>(mt.v + i) = v;
      *reinterpret_cast<float4*>(mt.p + i) = p;
      if (mt.shadow_kind == 1) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(mt.shadow) + i) = pk;
      } else if (mt.shadow_kind == 2) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(mt.shadow) + i) = p;
      }
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += OPT_THREADS) {
      upd(g[i], mt.m[i], mt.v[i], mt.p[i]);
      if (mt.shadow_kind == 1) reinterpret_cast<__nv_bfloat16*>(mt.shadow)[i] = __float2bfloat16(mt.p[i]);
      else if (mt.shadow_kind == 2) reinterpret_cast<float*>(mt.shadow)[i] = mt.p[i];
    }
  }
}

// in-place g *= min(1, max_norm / (sqrt(*grad_sq) + 1e-6)): torch.nn.utils.clip_grad_norm_ on an accumulating gradient
__global__ void __launch_bounds__(OPT_THREADS) grad_clip_kernel(const OptTensor* __restrict__ meta, float* const* __restrict__ grads,
                                                                const int32_t* __restrict__ blk_tensor,
                                                                const int32_t* __restrict__ blk_chunk,
                                                                const float* __restrict__ grad_sq, float max_norm) {
  const float clip = max_norm / (sqrtf(__ldg(grad_sq)) + 1e-6f);
  if (!(clip < 1.f)) return;   // torch clamps the coefficient to 1: nothing to do
  const int t = blk_tensor[blockIdx.x];
  float* g = grads[t];
  if (g == nullptr) return;
  const int64_t n = meta[t].n;
  const int64_t base = (int64_t)blk_chunk[blockIdx.x] * OPT_CHUNK;
  const int64_t end = min(base + OPT_CHUNK, n);
  if ((n & 3) == 0) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += OPT_THREADS * 4) {
      float4 x = *reinterpret_cast<float4*>(g + i);
      x.x *= clip; x.y *= clip; x.z *= clip; x.w *= clip;
      *reinterpret_cast<float4*>(g + i) = x;
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += OPT_THREADS) g[i] *= clip;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-tensor copy / cast of the data-parallel gradient buckets (parallel.GradReducer; the reference's only bucketing
// logic is the dead Horovod helper utils/distributed.py:46-97): up to MULTI_MAX (src, dst, n) triples travel BY VALUE in the
// kernel parameters -- no device-side table, so the launch can sit in a captured CUDA graph -- and one launch gathers
// the fp32 gradients of a layer into the (bf16 or fp32) bucket, or scatters the reduced bucket back, optionally scaled.
struct MultiCopyArgs {
  const void* src[MULTI_MAX];
  void* dst[MULTI_MAX];
  int64_t n[MULTI_MAX];
  int32_t blk0[MULTI_MAX + 1];   // first block of tensor i (prefix sum of ceil(n / OPT_CHUNK))
  int32_t count;
  int32_t src_bf16, dst_bf16;
  float scale;
};

template <bool SRC16, bool DST16>
__global__ void __launch_bounds__(OPT_THREADS) multi_copy_kernel(const __grid_constant__ MultiCopyArgs a) {
  int lo = 0, hi = a.count - 1;            // tensor of this block: last i with blk0[i] <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (a.blk0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  using SrcT = typename std::conditional<SRC16, __nv_bfloat16, float>::type;
  using DstT = typename std::conditional<DST16, __nv_bfloat16, float>::type;
  const SrcT* __restrict__ src = reinterpret_cast<const SrcT*>(a.src[lo]);
  DstT* __restrict__ dst = reinterpret_cast<DstT*>(a.dst[lo]);
  const int64_t n = a.n[lo];
  const int64_t base = (int64_t)((int)blockIdx.x - a.blk0[lo]) * OPT_CHUNK;
  const int64_t end = min(base + (int64_t)OPT_CHUNK, n);
  const float sc = a.scale;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) & (SRC16 ? 7 : 15)) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst) & (DST16 ? 7 : 15)) == 0);
  int64_t i = base + threadIdx.x * 4;
  if (vec) {
    for (; i + 4 <= end; i += OPT_THREADS * 4) {
      float4 x;
      if (SRC16) {
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(src + i));
        const float2 lo2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
        const float2 hi2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
        x = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
      } else {
        x = __ldg(reinterpret_cast<const float4*>(src + i));
      }
      x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
      if (DST16) {
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(x.x, x.y), h2 = __floats2bfloat162_rn(x.z, x.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&l2);
        pk.y = *reinterpret_cast<const uint32_t*>(&h2);
        *reinterpret_cast<uint2*>(dst + i) = pk;
      } else {
        *reinterpret_cast<float4*>(dst + i) = x;
      }
    }
    // ragged tail of the chunk (n % 4 != 0): the thread whose vector straddles `end`
    if (i < end) {
      for (int64_t j = i; j < end; ++j) {
        const float x = SRC16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[j]) : reinterpret_cast<const float*>(src)[j];
        if (DST16) reinterpret_cast<__nv_bfloat16*>(dst)[j] = __float2bfloat16(x * sc);
        else reinterpret_cast<float*>(dst)[j] = x * sc;
      }
    }
  } else {
    for (int64_t j = base + threadIdx.x; j < end; j += OPT_THREADS) {
      const float x = SRC16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[j]) : reinterpret_cast<const float*>(src)[j];
      if (DST16) reinterpret_cast<__nv_bfloat16*>(dst)[j] = __float2bfloat16(x * sc);
      else reinterpret_cast<float*>(dst)[j] = x * sc;
    }
  }
}

int multi_copy(const void* const* src, void* const* dst, const int64_t* n, int count, int src_bf16, int dst_bf16, float scale,
               cudaStream_t st) {
  if (count <= 0) return MODCR_OK;
  if (count > MULTI_MAX || src == nullptr || dst == nullptr || n == nullptr) {
    set_error("multi_copy: 1..64 (src, dst, n) triples per call");
    return MODCR_ERR_SHAPE;
  }
  MultiCopyArgs a;
  int64_t blocks = 0;
  for (int i = 0; i < count; ++i) {
    if (src[i] == nullptr || dst[i] == nullptr || n[i] < 0) {
      set_error("multi_copy: null tensor or negative length");
      return MODCR_ERR_SHAPE;
    }
    a.src[i] = src[i]; a.dst[i] = dst[i]; a.n[i] = n[i];
    a.blk0[i] = (int32_t)blocks;
    blocks += (n[i] + OPT_CHUNK - 1) / OPT_CHUNK;
    if (blocks > 0x7fffffff) {
      set_error("multi_copy: too many elements for one launch");
      return MODCR_ERR_SHAPE;
    }
  }
  for (int i = count; i <= MULTI_MAX; ++i) a.blk0[i] = (int32_t)blocks;
  for (int i = count; i < MULTI_MAX; ++i) { a.src[i] = nullptr; a.dst[i] = nullptr; a.n[i] = 0; }
  a.count = count; a.src_bf16 = src_bf16; a.dst_bf16 = dst_bf16; a.scale = scale;
  if (blocks == 0) return MODCR_OK;
  if (src_bf16 && dst_bf16) multi_copy_kernel<true, true><<<(unsigned)blocks, OPT_THREADS, 0, st>>>(a);
  else if (src_bf16) multi_copy_kernel<true, false><<<(unsigned)blocks, OPT_THREADS, 0, st>>>(a);
  else if (dst_bf16) multi_copy_kernel<false, true><<<(unsigned)blocks, OPT_THREADS, 0, st>>>(a);
  else multi_copy_kernel<false, false><<<(unsigned)blocks, OPT_THREADS, 0, st>>>(a);
  return check_launch("multi_copy");
}

int grad_sqnorm(const void* meta, const void* grads, const int32_t* blk_tensor, const int32_t* blk_chunk, int nblocks,
                float* out, cudaStream_t st) {
  if (nblocks <= 0) return MODCR_OK;
  if (meta == nullptr || grads == nullptr || blk_tensor == nullptr || blk_chunk == nullptr || out == nullptr) {
    set_error("grad_sqnorm: null table pointer");
    return MODCR_ERR_SHAPE;
  }
  grad_sqnorm_kernel<<<nblocks, OPT_THREADS, 0, st>>>((const OptTensor*)meta, (const float* const*)grads, blk_tensor, blk_chunk,
                                                     out);
  return check_launch("grad_sqnorm");
}

int grad_clip(const void* meta, const void* grads, const int32_t* blk_tensor, const int32_t* blk_chunk, int nblocks,
              const float* grad_sq, float max_norm, cudaStream_t st) {
  if (nblocks <= 0) return MODCR_OK;
  if (meta == nullptr || grads == nullptr || blk_tensor == nullptr || blk_chunk == nullptr || grad_sq == nullptr) {
    set_error("grad_clip: null table pointer");
    return MODCR_ERR_SHAPE;
  }
  if (!(max_norm > 0.f)) {
    set_error("grad_clip: max_norm must be positive");
    return MODCR_ERR_SHAPE;
  }
  grad_clip_kernel<<<nblocks, OPT_THREADS, 0, st>>>((const OptTensor*)meta, (float* const*)grads, blk_tensor, blk_chunk, grad_sq,
                                                   max_norm);
  return check_launch("grad_clip");
}

int adamw_step(const void* meta, const void* grads, const int32_t* blk_tensor, const int32_t* blk_chunk, int nblocks,
               const float* hyper, const float* grad_sq, float b1, float b2, float eps, cudaStream_t st) {
  if (nblocks <= 0) return MODCR_OK;
  if (meta == nullptr || grads == nullptr || blk_tensor == nullptr || blk_chunk == nullptr || hyper == nullptr) {
    set_error("adamw_step: null table pointer");
    return MODCR_ERR_SHAPE;
  }
  adamw_kernel<<<nblocks, OPT_THREADS, 0, st>>>((const OptTensor*)meta, (const float* const*)grads, blk_tensor, blk_chunk, hyper,
                                               grad_sq, b1, b2, eps);
  return check_launch("adamw_step");
}

}  // namespace modcr
