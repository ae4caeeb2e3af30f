// Host input assembly on the device: builds one padded CPT batch — input_ids, segment ids,
// attention mask, [MASK] positions and the padded region-feature tensor — from token-id lists and a packed feature store,
// replacing the per-sample Python of the reference's dataset + collate:
//   tokenize()            Oscar/oscar/datasets/refcoco_zsl_cpt_dataset.py:211-302   ([CLS] a [SEP] b [SEP], pair truncation
//                                                                                    :191-208, segment ids, zero padding,
//                                                                                    mask = 1 on tokens and on the boxes)
//   feature padding       :119-120   (torch.cat([feat, zeros(R - n, 2054)]))
//   [MASK] position       :118       (input_ids.tolist().index(103))
//   test_collate          Oscar/oscar/zeroshot/refcoco_cpt.py:159-172 (torch.stack of the per-row tensors)
// One CTA per row.  The feature rows are copied from the store with 16-byte loads when the row pitch allows it (2054
// floats = 8216 bytes: every second row is only 8-byte aligned, so the copy runs in float2).
#pragma once
#include "ptx.cuh"

namespace cptk {

struct AssembleParams {
  int B, T, R, F;
  const float* store;          // [rows, F] packed region features (device)
  const long long* feat_row0;  // [B] first store row of the sample
  const int* n_boxes;          // [B]
  const int* tok_a;            // flat token ids of text_a (the prompt caption), a_off [B + 1]
  const int* a_off;
  const int* tok_b;            // flat token ids of text_b (object tags), b_off [B + 1]
  const int* b_off;
  const int* has_b;            // [B] text_b was a non-empty string (pair truncation applies even if it tokenises to nothing)
  int cls_id, sep_id, pad_id, mask_id;
  long long* input_ids;        // [B, T]
  long long* segment_ids;      // [B, T]
  long long* input_mask;       // [B, T + R]
  long long* mask_pos;         // [B]
  float* img_feats;            // [B, R, F]
  int* err;                    // 5 = more boxes than R, 6 = no [MASK] token in the row
};

__global__ void __launch_bounds__(256) assemble_inputs_kernel(const AssembleParams p) {
  const int b = blockIdx.x;
  __shared__ int s_la, s_lb, s_mask;
  if (threadIdx.x == 0) {
    int la = p.a_off[b + 1] - p.a_off[b];
    int lb = p.has_b[b] ? p.b_off[b + 1] - p.b_off[b] : 0;
    if (p.has_b[b]) {
      // _truncate_seq_pair(tokens_a, tokens_b, T - 3): pop from the longer list, ties pop from b
      while (la + lb > p.T - 3) {
        if (la > lb) --la;
        else --lb;
      }
    } else if (la > p.T - 2) {
      la = p.T - 2;
    }
    s_la = la;
    s_lb = lb;
    s_mask = 0x7fffffff;
  }
  __syncthreads();
  const int la = s_la, lb = s_lb;
  const int n_tok = 1 + la + 1 + (lb > 0 ? lb + 1 : 0);
  int nb = p.n_boxes[b];
  if (nb > p.R) {
    if (threadIdx.x == 0) atomicExch(p.err, 5);
    nb = p.R;
  }
  const int* ta = p.tok_a + p.a_off[b];
  const int* tb = p.tok_b + p.b_off[b];
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
    int id = p.pad_id, seg = 0;
    if (t == 0) id = p.cls_id;
    else if (t <= la) id = ta[t - 1];
    else if (t == la + 1) id = p.sep_id;
    else if (lb > 0 && t <= la + 1 + lb) { id = tb[t - la - 2]; seg = 1; }
    else if (lb > 0 && t == la + 2 + lb) { id = p.sep_id; seg = 1; }
    p.input_ids[(long long)b * p.T + t] = id;
    p.segment_ids[(long long)b * p.T + t] = seg;
    p.input_mask[(long long)b * (p.T + p.R) + t] = t < n_tok ? 1 : 0;
    if (id == p.mask_id) atomicMin(&s_mask, t);
  }
  for (int r = threadIdx.x; r < p.R; r += blockDim.x) p.input_mask[(long long)b * (p.T + p.R) + p.T + r] = r < nb ? 1 : 0;
  // features: nb rows from the store, then zeros
  const long long total = (long long)p.R * p.F;
  const long long live = (long long)nb * p.F;
  const float* src = p.store + p.feat_row0[b] * p.F;
  float* dst = p.img_feats + (long long)b * total;
  const bool al8 = (((uintptr_t)src | (uintptr_t)dst) & 7) == 0 && (p.F % 2) == 0;
  if (al8) {
    const float2* s2 = reinterpret_cast<const float2*>(src);
    float2* d2 = reinterpret_cast<float2*>(dst);
    const long long n2 = total / 2, l2 = live / 2;
    for (long long i = threadIdx.x; i < n2; i += blockDim.x) d2[i] = i < l2 ? __ldg(s2 + i) : make_float2(0.f, 0.f);
  } else {
    for (long long i = threadIdx.x; i < total; i += blockDim.x) dst[i] = i < live ? __ldg(src + i) : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_mask == 0x7fffffff) {
      atomicExch(p.err, 6);
      p.mask_pos[b] = -1;
    } else {
      p.mask_pos[b] = s_mask;
    }
  }
}

}  // namespace cptk
