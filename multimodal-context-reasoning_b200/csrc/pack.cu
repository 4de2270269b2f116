// Device-side batch packer: replaces the reference collate function SNLIGPT_gen_collate
// (Data/VCRChunkAlign.py:690-741) and the per-choice tensor building of __getitem__ (:655-686).
//
// The reference creates ~15 small CUDA tensors per (question, choice) on the host thread, pads them with
// pad_sequence, pads every dense [L,L] chunk mask with two torch.cat calls and stacks the image features of the four
// choices of a question (four identical copies cross PCIe).  Here the host writes ONE pinned buffer
//   header  int32 [N,4]    : text length L, number of boxes, label, index of the question's feature block
//   tokens  int32 [N,T,2]  : word 0 = input id; word 1 = (chunk id + 1) | segment << 10 | align_pos << 11 | total_label << 12
//   feats   fp32  [U,R,D]  : one block per QUESTION (the four choices share it)
// which crosses PCIe once, and two kernels expand it into the padded batch tensors of the reference's batch contract
// plus the forms the kernels of this library consume directly (chunk-bit mask, chunk ids) -- the dense fp32 [N,T,T]
// chunk mask is never built unless the caller asks for it.
//
// Chunk mask semantics (utils/GetChunk_v4_vcr.py:104-140): for a sequence of L tokens
//   mask = eye(L);  mask[0, :L] = 1;  mask[L-1, :L] = 1;  mask[a, b] = 1 for a, b in the same chunk (tokens 1..L-2)
// zero-padded to [T,T] by the collate (:716-723).
#include <cuda_runtime.h>

#include <cstdint>

#include "rowops.h"
#include "tma_host.h"

namespace modcr {

struct PackParams {
  const int32_t* header;
  const int32_t* tokens;
  const float* feats;
  int N, T, R, D, cwords;
  int64_t* input_ids;
  int64_t* token_type_ids;
  int64_t* total_label;
  int64_t* align_pos;
  int64_t* label;
  float* label_f;
  float* input_mask;
  int32_t* chunk_id;
  uint32_t* chunk_bits;
  float* chunk_mask;
  float* img_feat;
};

constexpr int PACK_MAXT = 512;

// one CTA per sequence
__global__ void __launch_bounds__(256) pack_tokens_kernel(const PackParams p) {
  __shared__ int s_cid[PACK_MAXT];
  const int n = blockIdx.x;
  const int L = p.header[4 * n + 0], nbox = p.header[4 * n + 1], lab = p.header[4 * n + 2];
  const int T = p.T, R = p.R;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int2 tok = reinterpret_cast<const int2*>(p.tokens)[(size_t)n * T + t];
    const bool live = t < L;
    const uint32_t m = (uint32_t)tok.y;
    const int cid = live ? (int)(m & 1023u) - 1 : -1;
    s_cid[t] = cid;
    const size_t o = (size_t)n * T + t;
    p.input_ids[o] = live ? (int64_t)tok.x : 0;                       // pad_sequence(padding_value=0) :700
    if (p.token_type_ids) p.token_type_ids[o] = live ? (int64_t)((m >> 10) & 1u) : 0;
    if (p.align_pos) p.align_pos[o] = live ? (int64_t)((m >> 11) & 1u) : 0;
    if (p.total_label) p.total_label[o] = live ? (int64_t)(m >> 12) : 0;
    if (p.chunk_id) p.chunk_id[o] = cid;
    p.input_mask[(size_t)n * (T + R) + t] = live ? 1.f : 0.f;
  }
  for (int r = threadIdx.x; r < R; r += blockDim.x) p.input_mask[(size_t)n * (T + R) + T + r] = r < nbox ? 1.f : 0.f;
  if (threadIdx.x == 0) {
    if (p.label) p.label[n] = lab;
    if (p.label_f) p.label_f[n] = (float)lab;
  }
  __syncthreads();
  if (p.chunk_bits) {
    const int cw = p.cwords;
    for (int i = threadIdx.x; i < T * cw; i += blockDim.x) {
      const int q = i / cw, w = i - q * cw;
      uint32_t b = 0;
      if (q < L) {
        const bool full = (q == 0) || (q == L - 1);
        const int cq = s_cid[q];
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const int k = 32 * w + j;
          if (k < L && (full || k == q || (cq >= 0 && s_cid[k] == cq))) b |= 1u << j;
        }
      }
      p.chunk_bits[((size_t)n * T + q) * cw + w] = b;
    }
  }
  if (p.chunk_mask) {  // dense form of the reference batch contract (optional)
    for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
      const int q = i / T, k = i - q * T;
      float v = 0.f;
      if (q < L && k < L) {
        const int cq = s_cid[q];
        if (q == 0 || q == L - 1 || k == q || (cq >= 0 && s_cid[k] == cq)) v = 1.f;
      }
      p.chunk_mask[(size_t)n * T * T + i] = v;
    }
  }
}

// img_feat[n, r, :] = r < nbox[n] ? feats[block[n], r, :] : 0      (D even: 64-bit accesses)
__global__ void __launch_bounds__(256) pack_feats_kernel(const PackParams p) {
  const int D2 = p.D >> 1;
  const int64_t rows = (int64_t)p.N * p.R;
  const int warps_per_cta = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < rows;
       row += (int64_t)gridDim.x * warps_per_cta) {
    const int n = (int)(row / p.R), r = (int)(row - (int64_t)n * p.R);
    const int nbox = p.header[4 * n + 1], blk = p.header[4 * n + 3];
    float2* dst = reinterpret_cast<float2*>(p.img_feat + row * p.D);
    if (r < nbox) {
      const float2* src = reinterpret_cast<const float2*>(p.feats + ((int64_t)blk * p.R + r) * p.D);
      for (int c = lane; c < D2; c += 32) dst[c] = __ldg(src + c);
    } else {
      for (int c = lane; c < D2; c += 32) dst[c] = make_float2(0.f, 0.f);
    }
  }
}

int pack_batch(const PackDesc& d, cudaStream_t st) {
  if (d.N <= 0 || d.T <= 0 || d.T > PACK_MAXT || d.R < 0 || d.D < 0 || (d.D & 1) != 0) {
    set_error("pack_batch: bad shape N=%d T=%d (<= %d) R=%d D=%d (even)", d.N, d.T, PACK_MAXT, d.R, d.D);
    return MODCR_ERR_SHAPE;
  }
  if (d.header == nullptr || d.tokens == nullptr || d.input_ids == nullptr || d.input_mask == nullptr) {
    set_error("pack_batch: header, tokens, input_ids and input_mask are required");
    return MODCR_ERR_SHAPE;
  }
  if (d.chunk_bits != nullptr && d.cwords != (d.T + 31) / 32) {
    set_error("pack_batch: cwords=%d, expected ceil(T/32)=%d", d.cwords, (d.T + 31) / 32);
    return MODCR_ERR_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(d.tokens) & 7) != 0 || (reinterpret_cast<uintptr_t>(d.feats) & 7) != 0 ||
      (reinterpret_cast<uintptr_t>(d.img_feat) & 7) != 0) {
    set_error("pack_batch: tokens / feats / img_feat must be 8-byte aligned");
    return MODCR_ERR_ALIGN;
  }
  PackParams p;
  p.header = d.header; p.tokens = d.tokens; p.feats = d.feats; p.N = d.N; p.T = d.T; p.R = d.R; p.D = d.D;
  p.cwords = d.cwords; p.input_ids = d.input_ids; p.token_type_ids = d.token_type_ids; p.total_label = d.total_label;
  p.align_pos = d.align_pos; p.label = d.label; p.label_f = d.label_f; p.input_mask = d.input_mask;
  p.chunk_id = d.chunk_id; p.chunk_bits = d.chunk_bits; p.chunk_mask = d.chunk_mask; p.img_feat = d.img_feat;
  pack_tokens_kernel<<<d.N, 256, 0, st>>>(p);
  if (int rc = check_launch("pack_batch(tokens)")) return rc;
  if (d.img_feat != nullptr && d.R > 0 && d.D > 0) {
    if (d.feats == nullptr) {
      set_error("pack_batch: img_feat requested without feats");
      return MODCR_ERR_SHAPE;
    }
    const int64_t rows = (int64_t)d.N * d.R;
    int64_t grid = (rows + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    pack_feats_kernel<<<(int)grid, 256, 0, st>>>(p);
    return check_launch("pack_batch(feats)");
  }
  return MODCR_OK;
}

}  // namespace modcr
