"""GPU: the device-side CPT decision (cpt_score_queries) against the oracle's restatement of the reference's per-image
loops and of Oscar/oscar/utils/iou.py — decisions (argmax index, hit / miss) must be bit-exact."""
import random

import pytest
import torch

from cpt_b200 import comm
from cpt_b200.scoring import score_queries
from oracle import cpt_oracle as O

pytestmark = pytest.mark.gpu


def _case(seed, n_query, palette=5, tie=False, nan=False):
    rnd = random.Random(seed)
    g = torch.Generator().manual_seed(seed)
    fanouts, n_valid, rect_sets, rows, gts = [], [], [], [], []
    for _ in range(n_query):
        sets = rnd.randint(1, 4)
        fanouts.append(sets)
        per_query = []
        for j in range(sets):
            n = palette if j < sets - 1 else rnd.randint(1, palette)   # the last proposal set is usually shorter
            n_valid.append(n)
            rs = []
            for _ in range(n):
                x1, y1 = rnd.uniform(0, 400), rnd.uniform(0, 300)
                rs.append([x1, y1, x1 + rnd.uniform(2, 200), y1 + rnd.uniform(2, 200)])
            per_query.append(rs)
            row = torch.rand(palette + 1, generator=g) * 4 + 0.5
            if tie and n >= 2:
                row[1] = row[0]
            rows.append(row)
        rect_sets.append(per_query)
        x, y = rnd.uniform(0, 400), rnd.uniform(0, 300)
        gts.append([x, y, rnd.uniform(5, 250), rnd.uniform(5, 250)])
    logits = torch.stack(rows)
    if nan:
        logits[0, 0] = float("nan")
    return logits, fanouts, n_valid, rect_sets, gts


@pytest.mark.parametrize("few_shot", [False, True])
@pytest.mark.parametrize("seed,tie,nan", [(1, False, False), (2, True, False), (3, False, True), (4, False, False)])
def test_refcoco_decisions_and_iou_match_the_reference_loop(few_shot, seed, tie, nan):
    logits, fanouts, n_valid, rect_sets, gts = _case(seed, 37, tie=tie, nan=nan)
    flat_rects = [r for q in rect_sets for s in q for r in s]
    out = score_queries(logits.cuda(), fanouts, mode="fsl" if few_shot else "zsl", n_valid=n_valid, rects=flat_rects,
                        gt=gts)
    torch.cuda.synchronize()
    pick, rect, iou, ok = (out[k].cpu() for k in ("pick", "rect", "iou", "correct"))
    r = 0
    for q, sets in enumerate(fanouts):
        idx, want_rect = O.refcoco_decide([logits[r + j] for j in range(sets)], n_valid[r:r + sets], rect_sets[q],
                                          few_shot=few_shot)
        v, hit = O.refcoco_hit(want_rect, gts[q])
        assert int(pick[q]) == idx, (q, int(pick[q]), idx)
        assert rect[q].tolist() == want_rect
        assert float(iou[q]) == v and bool(ok[q]) == hit
        r += sets
    # the same decisions through the comm helper the sharded loop uses
    assert torch.equal(comm.pick_per_query(logits.cuda(), fanouts, "fsl" if few_shot else "zsl", n_valid=n_valid).cpu(),
                       pick)


def test_columns_outside_a_rows_colour_set_never_win():
    # row 1 of the query uses 2 of 5 colours; its strongest logit sits in an unused column
    lg = torch.tensor([[1.0, 2.0, 3.0, 1.5, 0.5, 9.0], [0.1, 0.2, 50.0, 60.0, 70.0, 9.0]])
    out = score_queries(lg.cuda(), [2], mode="zsl", n_valid=[5, 2])
    assert int(out["pick"][0]) == 2
    assert int(score_queries(lg.cuda(), [2], mode="zsl")["pick"][0]) == 5 + 4   # without the per-row sets it would


def test_vcr_choice_matches_oracle():
    g = torch.Generator().manual_seed(11)
    nsp = torch.randn(128, 2, generator=g) * 3
    pick = comm.pick_per_query(nsp.cuda(), [4] * 32, "vcr").cpu()
    s = O.vcr_choice_scores(nsp)
    for q in range(32):
        assert int(pick[q]) == int(s[4 * q:4 * q + 4].argmax())


def test_malformed_rectangle_raises_like_the_reference_assert():
    from cpt_b200.scoring import _handle
    lg = torch.tensor([[3.0, 1.0, 0.5]])
    score_queries(lg.cuda(), [1], mode="zsl", n_valid=[2], rects=[[10.0, 10.0, 5.0, 20.0], [0.0, 0.0, 4.0, 4.0]],
                  gt=[[0.0, 0.0, 10.0, 10.0]])
    with pytest.raises(RuntimeError):
        _handle(torch.device("cuda", torch.cuda.current_device())).check()
