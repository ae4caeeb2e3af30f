"""CPT decision per query on the device — ctypes binding of cpt_score_queries.

The reference post-processes the colour logits in a Python loop over images with a device->host synchronisation per
image (Oscar/oscar/zeroshot/refcoco_cpt.py:222-254, fewshot/refcoco_cpt.py:273-297, fewshot/vcr_nsp_cpt.py:600-604)
and scores a prediction as a hit when IoU > 0.5 (zeroshot/refcoco_cpt.py:268-276, Oscar/oscar/utils/iou.py:1-12).
`score_queries` does all of it for the whole (gathered) batch in one launch: per row the row's OWN colour set is
honoured (the last proposal set of a query usually has fewer rectangles, hence fewer colours), argmax follows
torch.argmax's tie rules, IoU runs in double precision like the reference's Python floats.  CUDA only.
"""
import ctypes as C

import torch

from . import _lib
from .config import BertConfig

MODES = {"zsl": 0, "fsl": 1, "vcr": 2}
_handles = {}
_index_cache = {}  # (device, fan-outs, colour-set sizes, K) -> (row_start, col_start) device tensors


def _handle(device):
    """A minimal library handle per device (the kernel needs none of the model's state)."""
    h = _handles.get(device)
    if h is None:
        from .engine import Engine
        cfg = BertConfig(vocab_size_or_config_json_file=8, hidden_size=128, num_hidden_layers=0, num_attention_heads=2,
                         intermediate_size=128, max_position_embeddings=8, type_vocab_size=2)
        h = Engine(cfg, device)
        _handles[device] = h
    return h


def _i32(x, device):
    return torch.as_tensor(x, dtype=torch.int32).to(device).contiguous()


def score_queries(logits, fanouts, mode="zsl", n_valid=None, rects=None, gt=None):
    """logits [rows, K] (CUDA fp32; zsl/fsl: K-1 palette colours then "none"; vcr: NSP scores).
    fanouts: rows per query.  n_valid: per row, how many of its colour columns are in use (None: all K-1).
    rects: [sum(n_valid), 4] float64 x1 y1 x2 y2 in collected order (row by row, colour by colour); gt: [Q, 4] float64
    x y w h.  Returns dict(pick=int32 [Q], and — when rects (and gt) are given — rect [Q,4], iou [Q], correct [Q])."""
    if not logits.is_cuda:
        raise _lib.CptError("cpt_b200: score_queries runs on a CUDA device only (no CPU path)")
    dev = logits.device
    eng = _handle(dev)
    lg = logits.to(torch.float32).contiguous()
    rows, K = lg.shape
    Q = len(fanouts)
    ckey = (dev, tuple(int(f) for f in fanouts), None if n_valid is None else tuple(int(x) for x in n_valid), K,
            MODES[mode] != 2)
    cached = _index_cache.get(ckey)
    if cached is None:  # the same split is scored batch after batch: its index tables are uploaded once
        starts = [0]
        for f in fanouts:
            starts.append(starts[-1] + int(f))
        if starts[-1] != rows:
            raise ValueError("score_queries: fan-outs sum to %d but there are %d rows" % (starts[-1], rows))
        row_start = _i32(starts, dev)
        col_start, n_cols = None, 0
        if MODES[mode] != 2:
            nv = [K - 1] * rows if n_valid is None else [int(x) for x in n_valid]
            if len(nv) != rows or any(x < 0 or x > K - 1 for x in nv):
                raise ValueError("score_queries: n_valid must give 0..K-1 colour columns for each of the %d rows" % rows)
            cs = [0]
            for x in nv:
                cs.append(cs[-1] + x)
            col_start, n_cols = _i32(cs, dev), cs[-1]
        if len(_index_cache) > 64:
            _index_cache.clear()
        cached = _index_cache[ckey] = (row_start, col_start, n_cols)
    row_start, col_start, n_cols = cached
    out = {"pick": torch.empty(Q, dtype=torch.int32, device=dev)}
    r = g = None
    if rects is not None:
        r = torch.as_tensor(rects, dtype=torch.float64).to(dev).contiguous()
        if col_start is not None and r.shape[0] != n_cols:
            raise ValueError("score_queries: %d rectangles for %d valid colour columns" % (r.shape[0], n_cols))
        out["rect"] = torch.empty(Q, 4, dtype=torch.float64, device=dev)
        if gt is not None:
            g = torch.as_tensor(gt, dtype=torch.float64).to(dev).contiguous()
            out["iou"] = torch.empty(Q, dtype=torch.float64, device=dev)
            out["correct"] = torch.empty(Q, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(eng.lib.cpt_score_queries(eng._h, C.c_void_p(torch.cuda.current_stream().cuda_stream), p(lg),
                                             lg.stride(0), K, Q, p(row_start), p(col_start), p(r), p(g), MODES[mode],
                                             p(out["pick"]), p(out.get("rect")), p(out.get("iou")),
                                             p(out.get("correct"))))
    return out
