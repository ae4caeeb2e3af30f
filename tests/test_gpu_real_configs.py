"""Parity at the configurations BASELINE.json names — full depth, full vocabulary, the bench's batch sizes — not only at
the reduced geometries of test_gpu_parity.py.  The CUDA path runs the whole batch (so the large-M code paths — the
dataflow chain kernel, 2-CTA tiles, graph replay — are the ones under test); the fp32 CPU oracle runs a strided sample
of the batch's rows (samples are independent), which it finishes in seconds.

  config 2  Oscar-base 12 layers, B=64, S=120 (70 tokens + 50 regions), 2 colour ids     — RefCOCO CPT inference
  config 3  Oscar-base 12 layers, V=30522, S=210, micro-batch 4, loss + every gradient  — GQA few-shot step, fp16 + bf16
  config 4  Oscar-base 12 layers, 128 rows, S=210, NSP head + 1-softmax[:,1]            — VCR q->a inference
  config 5  Oscar-large 24 layers / 1024, B=16, S=200, 2 colour ids                     — RefCOCO CPT, large
Tolerances are those of test_gpu_parity.py / test_gpu_train.py (1e-3 of the row maximum for inference)."""
import pytest
import torch

from cpt_b200 import config as C
from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _build(cfg, sd):
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.modeling_vcr import NSPCPT
    pre = BertImgForPreTraining(cfg)
    missing, unexpected = pre.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    pre.tie_weights()
    pre = pre.to("cuda").eval()
    rec, nsp = REC_MLM_CPT(cfg), NSPCPT(cfg)
    rec.copy_from_pretraining_model(pre)
    nsp.copy_from_pretraining_model(pre)
    return pre, rec.eval(), nsp.eval()


def _sample(b, idx):
    return {k: v[idx] for k, v in b.items()}


@pytest.mark.parametrize("model,B,T,R,K,stride", [("base", 64, 70, 50, 2, 8), ("large", 16, 150, 50, 2, 4)])
def test_cpt_inference_at_bench_configs(model, B, T, R, K, stride):
    """configs 2 and 5: colour logits at [MASK] (gather-first call), the hidden states and the unmodified [B,S,V] call."""
    from oracle import cpt_oracle as O
    cfg = C.oscar_base() if model == "base" else C.oscar_large()
    sd = synth_state_dict(cfg, seed=88)
    b = synth_batch(cfg, B, T, R, seed=88)
    vids = synth_vocab_ids(cfg, K, seed=88)
    pre, rec, nsp = _build(cfg, sd)
    d = {k: v.cuda() for k, v in b.items()}
    with torch.no_grad():
        logits = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                     mask_pos=d["mask_pos"], vocab_ids=vids.cuda())[0]
        again = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                    mask_pos=d["mask_pos"], vocab_ids=vids.cuda())[0]          # the graph replay
        seq = rec.bert(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0]
    rec.bert.engine().check()
    assert torch.equal(logits, again)
    idx = torch.arange(0, B, stride)
    s = _sample(b, idx)
    with torch.no_grad():
        oseq, _, _ = O.bert_img_model(sd, cfg, s["input_ids"], s["token_type_ids"], s["attention_mask"],
                                      img_feats=s["img_feats"])
        rows = O.lm_head(sd, cfg, oseq[torch.arange(len(idx)), s["mask_pos"]])
    assert (seq[idx.cuda()].cpu() - oseq).abs().max().item() <= RTOL * oseq.abs().max().item()
    row_max = rows.abs().max(dim=1, keepdim=True).values
    err = ((logits[idx.cuda()].cpu() - rows[:, vids]).abs() / row_max).max().item()
    assert err <= RTOL, "max rel-to-row-max error %.3e" % err
    if model == "base":
        # the reference's unmodified call on a slice of the batch: full scores, then the caller's gather
        n = 8
        with torch.no_grad():
            scores = rec(d["input_ids"][:n], d["token_type_ids"][:n], d["attention_mask"][:n],
                         img_feats=d["img_feats"][:n])[0]
        assert scores.shape == (n, T + R, cfg.vocab_size)
        got = scores[torch.arange(n), d["mask_pos"][:n]][:, vids.cuda()].cpu()
        assert idx[1] == n and idx[0] == 0
        assert ((got[0] - rows[0, vids]).abs() <= 2 * RTOL * row_max[0]).all()


def test_vcr_nsp_inference_at_bench_config():
    """config 4: 128 answer rows (32 questions x 4 answers), S=210, NSP logits and the answer scores."""
    from oracle import cpt_oracle as O
    cfg = C.oscar_base()
    sd = synth_state_dict(cfg, seed=88)
    B, T, R, stride = 128, 165, 45, 16
    b = synth_batch(cfg, B, T, R, seed=89)
    pre, rec, nsp = _build(cfg, sd)
    d = {k: v.cuda() for k, v in b.items()}
    with torch.no_grad():
        out = nsp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0]
        again = nsp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0]
    nsp.bert.engine().check()
    assert out.shape == (B, cfg.num_contrast_classes) and torch.equal(out, again)
    idx = torch.arange(0, B, stride)
    s = _sample(b, idx)
    with torch.no_grad():
        ref = O.nsp_cpt(sd, cfg, s["input_ids"], s["token_type_ids"], s["attention_mask"], img_feats=s["img_feats"])[0]
    got = out[idx.cuda()].cpu()
    assert (got - ref).abs().max().item() <= RTOL * max(1.0, ref.abs().max().item())
    assert (O.vcr_choice_scores(got) - O.vcr_choice_scores(ref)).abs().max().item() <= RTOL


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_gqa_few_shot_step_at_base_geometry(dtype):
    """config 3: one few-shot step at full geometry — 12 layers, V=30522, S=210 (165 + 45), micro-batch 4, one labelled
    [MASK] per row — loss and EVERY parameter gradient against autograd through the fp32 oracle."""
    from test_gpu_train import LOSS_SCALE, LTOL, build_rec, compare_all, oracle_grads
    cfg = C.oscar_base()
    sd = synth_state_dict(cfg, seed=88)
    B, T, R = 4, 165, 45
    b = synth_batch(cfg, B, T, R, seed=90)
    answers = synth_vocab_ids(cfg, 1853, seed=88)
    labels = torch.full((B, T + R), -1, dtype=torch.long)
    labels[torch.arange(B), b["mask_pos"]] = answers[torch.tensor([3, 500, 1000, 1852])]
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.0
    ref_loss, ref = oracle_grads(cfg, sd, b, labels)
    rec = build_rec(cfg, sd, dtype)
    d = {k: v.cuda() for k, v in b.items()}
    loss, _ = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                  masked_lm_labels=labels.cuda())
    (loss * LOSS_SCALE[dtype]).backward()
    rec.bert.train_engine()[0].check()
    assert abs(loss.item() - ref_loss.item()) <= LTOL[dtype] * abs(ref_loss.item())

    def key_of(k):
        key = k if k.startswith("bert.") else "cls.predictions." + k[len("cls."):]
        return None if key == "cls.predictions.decoder.weight" else key

    compare_all(rec.named_parameters(), ref, key_of, dtype, LOSS_SCALE[dtype], 16 * 12 + 10)
