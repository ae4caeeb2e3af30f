"""The input path: extractor wire format -> feature store -> padded batch.  CPU part: the TSV decoder and the store
against the fixture the reference's own decode_features / tokenize produced (tests/golden/make_golden_inputs.py).
GPU part: the device-side assembly (cpt_assemble_inputs) is bit-exact against the same fixture."""
import os

import numpy as np
import pytest
import torch

from cpt_b200 import featstore as FS


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "inputs_tiny.pt")), os.path.join(golden_dir, "inputs_tiny.tsv")


def test_tsv_decoder_matches_reference_decode_features(golden):
    g, tsv = golden
    lines = [l for l in open(tsv) if l.strip()]
    assert len(lines) == len(g["images"])
    row = 0
    for line, want in zip(lines, g["images"]):
        name, labels, feats, caption, colors, rects = FS.decode_prediction_row(line)
        assert name == want["key"] and caption == want["caption"] and colors == want["colors"]
        assert rects == want["rects"] and labels == want["od_labels"]
        for f in feats:
            assert f.dtype == np.float32 and f.shape[1] == 2054
            n = f.shape[0]
            got = torch.from_numpy(f.copy()).double().sum(1)
            assert torch.equal(got, g["img_feats_rowsums"][row, :n])       # bit-exact features, in order
            assert float(g["img_feats_rowsums"][row, n:].abs().sum()) == 0.0
            row += 1
    assert row == g["input_ids"].shape[0]


def test_feature_store_roundtrip(golden, tmp_path):
    g, tsv = golden
    n = FS.build_feature_store(tsv, str(tmp_path))
    assert n == len(g["images"])
    st = FS.FeatureStore(str(tmp_path))
    assert st.rows == sum(r["n_boxes"] for r in g["rows"]) and st.features.shape == (st.rows, 2054)
    lines = [l for l in open(tsv) if l.strip()]
    for i, line in enumerate(lines):
        _, labels, feats, caption, colors, rects = FS.decode_prediction_row(line)
        im = st.images[i]
        assert im["caption"] == caption and im["colors"] == colors and im["rects"] == rects and im["od_labels"] == labels
        for j, f in enumerate(feats):
            assert np.array_equal(st.set_features(i, j), f)
    assert abs(float(np.asarray(st.features, np.float64).sum()) - g["img_feats_sum"]) <= 1e-6 * abs(g["img_feats_sum"])
    with pytest.raises(RuntimeError):   # assembly is a CUDA kernel: no CPU path
        st.assemble([(0, 0)], [[103]], [None])


@pytest.mark.gpu
def test_device_assembly_is_bit_exact_against_the_reference_tokenize_and_collate(golden, tmp_path):
    g, tsv = golden
    FS.build_feature_store(tsv, str(tmp_path))
    st = FS.FeatureStore(str(tmp_path)).to("cuda")
    samples = [(r["img"], r["set"]) for r in g["rows"]]
    out = st.assemble(samples, [r["tokens_a"] for r in g["rows"]], [r["tokens_b"] for r in g["rows"]], T=g["T"], R=g["R"])
    torch.cuda.synchronize()
    assert torch.equal(out["input_ids"].cpu(), g["input_ids"])
    assert torch.equal(out["token_type_ids"].cpu(), g["segment_ids"])
    assert torch.equal(out["attention_mask"].cpu(), g["input_mask"])
    assert torch.equal(out["mask_pos"].cpu(), g["mask_pos"])
    feats = out["img_feats"].cpu()
    assert tuple(feats.shape) == tuple(g["img_feats_shape"])
    assert torch.equal(feats.double().sum(2), g["img_feats_rowsums"])
    for k, r in enumerate(g["rows"]):
        want = torch.from_numpy(np.array(st.set_features(r["img"], r["set"])))
        assert torch.equal(feats[k, :r["n_boxes"]], want) and float(feats[k, r["n_boxes"]:].abs().sum()) == 0.0


@pytest.mark.gpu
def test_device_assembly_flags_a_prompt_without_mask(golden, tmp_path):
    from cpt_b200.scoring import _handle
    g, tsv = golden
    FS.build_feature_store(tsv, str(tmp_path))
    st = FS.FeatureStore(str(tmp_path)).to("cuda")
    out = st.assemble([(0, 0)], [[2000, 2001]], [[3000]])   # the reference's input_ids.index(103) raises ValueError
    assert int(out["mask_pos"][0]) == -1
    with pytest.raises(RuntimeError):
        _handle(torch.device("cuda", torch.cuda.current_device())).check()
