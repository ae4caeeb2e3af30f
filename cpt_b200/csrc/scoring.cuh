// CPT decision per query on the device: the colour-logit post-processing the reference runs as a
// Python loop with one device->host synchronisation per image —
//   zero-shot RefCOCO  Oscar/oscar/zeroshot/refcoco_cpt.py:222-254 : per row keep the row's own colour columns
//                      (cur_color_set, a prefix of the palette; the trailing "none" column is dropped), concatenate the
//                      rows of the query, argmax -> the rectangle at that position
//   few-shot RefCOCO   Oscar/oscar/fewshot/refcoco_cpt.py:273-297  : the same with colour / none as the score
//   VCR                Oscar/oscar/fewshot/vcr_nsp_cpt.py:600-604  : score = 1 - softmax(nsp)[:, 1], argmax over the rows
//   hit / miss         zeroshot/refcoco_cpt.py:268-276 + Oscar/oscar/utils/iou.py:1-12 : IoU(pred, gt) > 0.5, in double
//                      precision like the Python floats of the reference
// One warp per query; argmax has torch.argmax's semantics (first maximal value; NaN counts as maximal).
#pragma once
#include "ptx.cuh"

namespace cptk {

enum ScoreMode { SCORE_ZSL = 0, SCORE_FSL = 1, SCORE_VCR = 2 };

struct ScoreParams {
  const float* logits;       // [rows, ld]
  long long ld;
  int K;                     // columns in use: K - 1 colours + "none" (zsl / fsl); NSP classes (vcr)
  int Q;
  const int* row_start;      // [Q + 1]
  const int* col_start;      // [rows + 1] CSR of the VALID colour columns of every row, in collected order (NULL: K - 1 each)
  const double* rects;       // [total valid columns, 4] x1 y1 x2 y2 in collected order (NULL: no rectangle output)
  const double* gt;          // [Q, 4] x y w h (NULL: no IoU)
  int mode;
  int* pick;                 // [Q] index into the query's collected scores (the reference's max_idx)
  double* pick_rect;         // [Q, 4] (NULL ok)
  double* iou;               // [Q] (NULL ok)
  int* correct;              // [Q] iou > 0.5 (NULL ok)
  int* err;                  // device error flag: 4 = malformed rectangle (the reference's assert p[2] > p[0] ...)
};

// (value, index) ordering of torch.argmax: NaN beats everything, ties go to the smaller index
__device__ __forceinline__ bool score_better(float v, int i, float bv, int bi) {
  if (bi < 0) return true;
  if (i < 0) return false;
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(128) score_queries_kernel(const ScoreParams p) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= p.Q) return;
  const int r0 = p.row_start[q], r1 = p.row_start[q + 1];
  float best = 0.f;
  int bidx = -1;
  if (p.mode == SCORE_VCR) {
    for (int r = r0 + lane; r < r1; r += 32) {
      const float* x = p.logits + (long long)r * p.ld;
      float mx = x[0];
      for (int c = 1; c < p.K; ++c) mx = fmaxf(mx, x[c]);
      float sum = 0.f;
      for (int c = 0; c < p.K; ++c) sum += expf(x[c] - mx);
      const float s = 1.0f - expf(x[1] - mx) / sum;
      if (score_better(s, r - r0, best, bidx)) {
        best = s;
        bidx = r - r0;
      }
    }
  } else {
    const int base = p.col_start ? p.col_start[r0] : r0 * (p.K - 1);
    for (int r = r0; r < r1; ++r) {
      const int c0 = p.col_start ? p.col_start[r] : r * (p.K - 1);
      const int nv = p.col_start ? p.col_start[r + 1] - c0 : p.K - 1;
      const float* x = p.logits + (long long)r * p.ld;
      const float none = x[p.K - 1];
      for (int c = lane; c < nv; c += 32) {
        const float s = p.mode == SCORE_FSL ? x[c] / none : x[c];
        const int idx = c0 - base + c;
        if (score_better(s, idx, best, bidx)) {
          best = s;
          bidx = idx;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (score_better(ov, oi, best, bidx)) {
      best = ov;
      bidx = oi;
    }
  }
  if (lane != 0) return;
  p.pick[q] = bidx;
  if (p.rects == nullptr || p.mode == SCORE_VCR || bidx < 0) return;
  const int base = p.col_start ? p.col_start[r0] : r0 * (p.K - 1);
  const double* rc = p.rects + 4ll * (base + bidx);
  const double x1 = rc[0], y1 = rc[1], x2 = rc[2], y2 = rc[3];
  if (p.pick_rect) {
    p.pick_rect[4ll * q + 0] = x1;
    p.pick_rect[4ll * q + 1] = y1;
    p.pick_rect[4ll * q + 2] = x2;
    p.pick_rect[4ll * q + 3] = y2;
  }
  if (p.gt == nullptr) return;
  if (!(x2 > x1 && y2 > y1)) atomicExch(p.err, 4);
  // [x1, y1, x2, y2] -> [x, y, w, h] with the reference's +1 (refcoco_cpt.py:271), then iou.py:1-12
  const double aw = x2 - x1 + 1.0, ah = y2 - y1 + 1.0;
  const double* g = p.gt + 4ll * q;
  const double ix1 = fmax(x1, g[0]), iy1 = fmax(y1, g[1]);
  const double ix2 = fmin(x1 + aw - 1.0, g[0] + g[2] - 1.0), iy2 = fmin(y1 + ah - 1.0, g[1] + g[3] - 1.0);
  // explicit round-to-nearest products / sums: no fused multiply-add, so every intermediate is the double the reference's
  // Python arithmetic produces and the IoU is bit-identical, not just the decision
  double inter = 0.0;
  if (ix1 < ix2 && iy1 < iy2) inter = __dmul_rn(ix2 - ix1 + 1.0, iy2 - iy1 + 1.0);
  const double uni = __dadd_rn(__dadd_rn(__dmul_rn(aw, ah), __dmul_rn(g[2], g[3])), -inter);
  const double v = inter / uni;
  if (p.iou) p.iou[q] = v;
  if (p.correct) p.correct[q] = v > 0.5 ? 1 : 0;
}

}  // namespace cptk
