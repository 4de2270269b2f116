// Internal C++ interface of the fused prefix-attention kernels (see attn.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace modcr {

// Masking modes.  All additive biases are fp32, NEG = -10000.
//   0  key bias only:  bias[q,k] = kbias[n,k]                    (BertImgModel, seq_enc layers 3..8)
//   1  seq_enc layers 0..2: text q/text k -> chunk bit ? 0 : NEG ; text q/img k -> kbias ;
//                           img q/text k -> NEG ; img q/img k -> kbias
//   2  seq_enc layers 9..11: as mode 1 but img q/img k -> (k == q ? 0 : NEG)
struct AttnDesc {
  int N = 0, S = 0, P = 0, heads = 0;   // head dim is fixed at 64; H = heads * 64
  int T = 0;                            // text length (modes 1/2 and probs_img), else 0
  int mode = 0;
  float scale = 0.125f;
  const void* qkv = nullptr;            // bf16 [N,S,3H]   (Q | K | V)
  const void* pkv = nullptr;            // bf16 [N,P,2H]   (K | V) of the prefix rows, P > 0 only
  const float* kbias = nullptr;         // fp32 [N,P+S]    additive key bias
  const uint32_t* cbits = nullptr;      // [N,T,cwords]    chunk bit mask (modes 1/2)
  int cwords = 0;
  void* ctx = nullptr;                  // bf16 [N,S,H]    forward output
  float* lse = nullptr;                 // fp32 [N,heads,S]
  float* probs_img = nullptr;           // fp32 [N,heads,T,S-T] normalised text->region probabilities (optional)
  const int64_t* row_sel = nullptr;     // int64 [N,T] (align_pos) or null: only rows with a non-zero entry are written to /
                                        // read from probs_img and receive dalign (the alignment loss uses no other row)
  // backward only
  const void* dctx = nullptr;           // bf16 [N,S,H]
  const float* dalign = nullptr;        // fp32 [N,T,S-T] upstream gradient of the head-summed block (optional)
  void* dqkv = nullptr;                 // bf16 [N,S,3H]
  void* dpkv = nullptr;                 // bf16 [N,P,2H]
  float* dq_ws = nullptr;               // fp32 [N*heads, S, 64] workspace (only touched when Kp > 128)
  // training-mode dropout of the attention probabilities: mask = f(*drop_seed, drop_site, (n, head, q, key))
  const uint32_t* drop_seed = nullptr;  // device pointer to the step's seed
  uint32_t drop_site = 0;
  float drop_p = 0.f;
};

int attn_fwd_launch(const AttnDesc& a, cudaStream_t stream);
int attn_bwd_launch(const AttnDesc& a, cudaStream_t stream);
size_t attn_bwd_workspace_bytes(int N, int S, int heads);
// developer builds (-DMODCR_ATTN_TRACE) only: copies the recorded clock64() timeline of attn_bwd_kernel to the host
int attn_trace_read(long long* host, int max_entries);

}  // namespace modcr
