#!/usr/bin/env python
"""Golden fixture for the input path, produced by the reference's OWN functions — run in the
build container (needs /root/reference):
    python tests/golden/make_golden_inputs.py
Writes tests/golden/inputs_tiny.tsv (a predictions.tsv in the extractor's schema, written the way
prompt_feat/maskrcnn_benchmark/engine/inference_ref.py:129-192 writes it) and tests/golden/inputs_tiny.pt with what
Oscar/oscar/datasets/refcoco_zsl_cpt_dataset.py makes of it: decode_features (:161-180), tokenize (:211-302), the feature
padding of __getitem__ (:119-120), the [MASK] position (:118), stacked as test_collate does
(Oscar/oscar/zeroshot/refcoco_cpt.py:159-172).  The tokenizer is a stand-in (whitespace split, fixed word -> id map):
tokenisation proper is not on the path; what is pinned here is everything AFTER it."""
import base64
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/Oscar")
from oscar.datasets import refcoco_zsl_cpt_dataset as D  # noqa: E402
from oscar.utils.tsv_file import TSVFile  # noqa: E402


class StubTokenizer(object):
    SPECIAL = {"[CLS]": 101, "[SEP]": 102, "[MASK]": 103, "[PAD]": 0}

    def tokenize(self, text):
        return text.lower().replace(".", " .").split()

    def convert_tokens_to_ids(self, toks):
        if isinstance(toks, str):
            toks = [toks]
        return [self.SPECIAL.get(t.upper() if t.startswith("[") else t, 1000 + (sum(ord(c) * (i + 7) for i, c in enumerate(t)) % 29000))
                for t in toks]


def main():
    rng = np.random.RandomState(88)
    palette = ["red", "purple", "green", "yellow", "blue"]
    classes = ["man", "woman", "dog", "table", "car", "tree", "sky", "window", "shirt", "plate"]
    rows = []
    for k in range(5):
        w, h = 640, 480
        n_sets = int(rng.randint(1, 4))
        boxlists, colors, rects = [], [], []
        for s in range(n_sets):
            n = 5 if s < n_sets - 1 else int(rng.randint(1, 6))
            bl, rs = [], []
            for i in range(n + int(rng.randint(0, 18))):  # the proposals of the set plus context boxes
                x1, y1 = rng.uniform(0, w - 40), rng.uniform(0, h - 40)
                x2, y2 = x1 + rng.uniform(10, w - x1 - 1), y1 + rng.uniform(10, h - y1 - 1)
                pooled = np.maximum(rng.randn(2048), 0).astype(np.float32)
                geo = np.array([x1 / w, y1 / h, x2 / w, y2 / h, (x2 - x1) / w, (y2 - y1) / h], np.float32)
                feat = np.concatenate([pooled, geo]).astype(np.float32)
                bl.append({"rect": [x1, y1, x2, y2], "bbox_id": i, "class": classes[int(rng.randint(len(classes)))],
                           "conf": float(rng.uniform(0.2, 1.0)), "feature": base64.b64encode(feat).decode("utf-8")})
                if i < n:
                    rs.append([x1, y1, x2, y2])
            boxlists.append(bl)
            colors.append(palette[:n])
            rects.append(rs)
        n_words = [4, 75, 12, 50, 66][k]  # short, and long enough for the pair truncation to cut text_a, text_b or both
        caption = " ".join(["the"] + [classes[int(rng.randint(len(classes)))] for _ in range(n_words)])
        rows.append(("%d" % (10800 + k), json.dumps({"objects": [boxlists, caption, colors, rects]})))
    tsv = os.path.join(HERE, "inputs_tiny.tsv")
    with open(tsv, "w") as f:
        for key, js in rows:
            f.write(key + "\t" + js + "\n")
    feat_tsv = TSVFile(tsv, generate_lineidx=True)
    tok = StubTokenizer()
    T, R = 70, 50
    out = {"T": T, "R": R, "rows": []}
    ids_l, mask_l, seg_l, feat_l, mpos_l = [], [], [], [], []
    for img_idx in range(len(rows)):
        img_name, od_labels, im_feats, caption, colors, rect_lists = D.ZSLColorFinetuneDataset.decode_features(None, feat_tsv, img_idx)
        for j, (labels, feat) in enumerate(zip(od_labels, im_feats)):
            posi = caption.index(" ", 4)  # after the second word: [MASK] early, it survives the truncation of long captions
            text_a = D.template4(caption, [posi])
            text_b = labels if (img_idx + j) % 4 != 3 else ""   # one row in four without object tags
            input_ids, input_mask, segment_ids, _ = D.tokenize(tok, text_a=text_a, text_b=text_b, img_feat=feat,
                                                               max_img_seq_len=R, max_seq_a_len=40, max_seq_len=T,
                                                               cls_token_segment_id=0, pad_token_segment_id=0,
                                                               sequence_a_segment_id=0, sequence_b_segment_id=1)
            padded = torch.cat([feat, torch.zeros([R - feat.size(0), 2054])], 0)
            ids_l.append(input_ids)
            mask_l.append(input_mask)
            seg_l.append(segment_ids)
            feat_l.append(padded)
            mpos_l.append(input_ids.tolist().index(103))
            out["rows"].append({"img": img_idx, "set": j, "key": img_name, "tokens_a": tok.convert_tokens_to_ids(tok.tokenize(text_a)),
                                "tokens_b": tok.convert_tokens_to_ids(tok.tokenize(text_b)) if text_b else None,
                                "od_labels": labels, "n_boxes": int(feat.size(0))})
        out.setdefault("images", []).append({"key": img_name, "caption": caption, "colors": colors, "rects": rect_lists,
                                             "od_labels": od_labels})
    out["input_ids"] = torch.stack(ids_l, 0)
    out["input_mask"] = torch.stack(mask_l, 0)
    out["segment_ids"] = torch.stack(seg_l, 0)
    out["mask_pos"] = torch.tensor(mpos_l, dtype=torch.long)
    # the feature tensor itself is reproduced from the TSV by the tests (12 MB); its checksum pins it
    feats = torch.stack(feat_l, 0)
    out["img_feats_shape"] = tuple(feats.shape)
    out["img_feats_sum"] = feats.double().sum().item()
    out["img_feats_rowsums"] = feats.double().sum(dim=2)
    torch.save(out, os.path.join(HERE, "inputs_tiny.pt"))
    os.remove(os.path.splitext(tsv)[0] + ".lineidx")
    print("wrote", tsv, os.path.getsize(tsv) // 1024, "KB;", len(out["rows"]), "rows")


if __name__ == "__main__":
    main()
