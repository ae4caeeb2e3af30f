"""Drop-in demonstration: the reference's OWN evaluation loops — zeroshot/refcoco_cpt.py val() (lines 208-288) and
fewshot/gqa_cpt.py evaluate() (lines 558-636) — imported unmodified from the offline install in baseline/_ref
(baseline/install_ref.sh), run with cpt_b200's modules aliased in as INTEGRATION.md §2 shows, on a synthetic loader, on
the GPU; the same loops then run on the reference's own fp32 CPU modules and the outcomes are compared.

The loops call the model the unmodified way (`model(ids, seg, mask, img_feats=f)[0]` -> full [B,S,V] scores), so this is
also the test of the unmodified call.  oracle/ref_shim.py supplies the un-vendored pytorch-transformers 1.x symbols the
reference imports (test infrastructure)."""
import logging
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

pytestmark = pytest.mark.gpu

COLOURS = ["red", "purple", "green", "yellow", "blue", "none"]


class _Tokenizer(object):
    """Word -> id table standing in for BertTokenizer (no vocab.txt in this container)."""

    def __init__(self, words, vocab_size, seed=5):
        g = torch.Generator().manual_seed(seed)
        ids = (torch.randperm(vocab_size - 1000, generator=g)[:len(words)] + 1000).tolist()
        self.table = dict(zip(words, ids))

    def tokenize(self, text):
        return text.split()

    def convert_tokens_to_ids(self, toks):
        return [self.table[t] for t in toks]


@pytest.fixture(scope="module")
def dropin():
    """Aliases cpt_b200's classes into the reference's module namespace BEFORE the task scripts are imported (the
    INTEGRATION.md §2 recipe), keeps the reference's own classes for the CPU arm, restores everything afterwards."""
    if not os.path.isfile(os.path.join(REF, "oscar", "zeroshot", "refcoco_cpt.py")):
        pytest.skip("baseline/_ref has no task scripts (run baseline/install_ref.sh in the build container)")
    from oracle import ref_shim
    ref_shim.install(REF)
    import oscar.modeling.modeling_bert as ref_mb
    import oscar.modeling.modeling_rec as ref_mr
    import cpt_b200.modeling_bert as mb
    import cpt_b200.modeling_rec as mr
    keep = dict(pre=ref_mb.BertImgForPreTraining, model=ref_mb.BertImgModel, heads=ref_mb.BertPreTrainingHeads,
                rec=ref_mr.REC_MLM_CPT)

    def alias(on):
        for name, k in (("BertImgModel", "model"), ("BertImgForPreTraining", "pre"), ("BertPreTrainingHeads", "heads")):
            setattr(ref_mb, name, getattr(mb, name) if on else keep[k])
        ref_mr.REC_MLM_CPT = mr.REC_MLM_CPT if on else keep["rec"]
    alias(True)
    for m in ("oscar.zeroshot.refcoco_cpt", "oscar.fewshot.gqa_cpt"):
        sys.modules.pop(m, None)
    import oscar.zeroshot.refcoco_cpt as zs
    import oscar.fewshot.gqa_cpt as gq
    assert zs.REC_MLM_CPT is mr.REC_MLM_CPT and zs.BertImgForPreTraining is mb.BertImgForPreTraining
    assert gq.REC_MLM_CPT is mr.REC_MLM_CPT and gq.BertImgForPreTraining is mb.BertImgForPreTraining
    zs.logger = logging.getLogger("dropin")          # main() sets this global (refcoco_cpt.py:291ff)
    yield types.SimpleNamespace(zs=zs, gq=gq, ref=keep, shim=ref_shim, alias=alias)
    alias(False)
    for m in ("oscar.zeroshot.refcoco_cpt", "oscar.fewshot.gqa_cpt"):
        sys.modules.pop(m, None)


def _build_models(d, tmp_path, layers=2):
    """Both arms built the way the scripts build theirs (refcoco_cpt.py:437-451): a checkpoint directory ->
    from_pretrained -> REC_MLM_CPT(config).copy_from_pretraining_model(tmp) -> .to(device)."""
    from cpt_b200 import config as C
    from cpt_b200.synthetic import synth_state_dict
    cfg = C.oscar_base(num_hidden_layers=layers)
    sd = synth_state_dict(cfg)
    ck = str(tmp_path / "checkpoint")
    os.makedirs(ck, exist_ok=True)
    cfg.save_pretrained(ck)
    torch.save(sd, os.path.join(ck, "pytorch_model.bin"))
    # GPU arm: the names the script itself resolved at import
    config = d.zs.BertImgForPreTraining.config_class.from_pretrained(ck)
    tmp = d.zs.BertImgForPreTraining.from_pretrained(ck, config=config)
    ours = d.zs.REC_MLM_CPT(config)
    ours.copy_from_pretraining_model(tmp)
    ours.to("cuda")
    # CPU arm: the reference's own classes (their `super(BertImgForPreTraining, self)` resolves the module-level name,
    # so the aliases are lifted while they are constructed)
    d.alias(False)
    dd = cfg.to_dict()
    v = dd.pop("vocab_size")
    rcfg = d.shim.BertConfig(v, **dd)
    pre = d.ref["pre"](rcfg)
    missing, unexpected = pre.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    pre.tie_weights()
    ref = d.ref["rec"](rcfg)
    ref.copy_from_pretraining_model(pre)
    d.alias(True)
    return cfg, ours, ref


class _RefcocoLoader(object):
    """What make_data_loader + test_collate (refcoco_cpt.py:159-205) hand to val(): (img_keys, (img_feats, input_ids,
    input_mask, segment_ids, mask_pos, colors, rects)) with one model row per colour set, and dataset.anns_dic."""

    def __init__(self, cfg, n_batches=3, images_per_batch=4, T=70, R=50, seed=11):
        from cpt_b200.synthetic import synth_batch
        g = torch.Generator().manual_seed(seed)
        self.batches, anns, key = [], {}, 0
        for b in range(n_batches):
            keys, colors, rects = [], [], []
            for _ in range(images_per_batch):
                sets = int(torch.randint(1, 4, (1,), generator=g))
                cl, rl = [], []
                for _s in range(sets):
                    k = int(torch.randint(2, 6, (1,), generator=g))
                    cl.append(COLOURS[:k])
                    xy = torch.randint(0, 300, (k, 2), generator=g)
                    wh = torch.randint(20, 200, (k, 2), generator=g)
                    rl.append([[int(x), int(y), int(x + w), int(y + h)] for (x, y), (w, h) in zip(xy.tolist(), wh.tolist())])
                keys.append(str(key))
                colors.append(cl)
                rects.append(rl)
                # ground truth = one of the candidate rectangles (xywh), so that hits and misses both occur
                allr = [r for s in rl for r in s]
                r = allr[int(torch.randint(0, len(allr), (1,), generator=g))]
                anns[str(key)] = {"bbox": [r[0], r[1], r[2] - r[0] + 1, r[3] - r[1] + 1], "file_name": "%d.jpg" % key,
                                  "caption": "synthetic"}
                key += 1
            rows = sum(len(c) for c in colors)
            sb = synth_batch(cfg, rows, T=T, R=R, seed=seed + b)
            self.batches.append((keys, (sb["img_feats"], sb["input_ids"], sb["attention_mask"], sb["token_type_ids"],
                                        sb["mask_pos"], colors, rects)))
        self.dataset = types.SimpleNamespace(anns_dic=anns)

    def __iter__(self):
        return iter(self.batches)


def _capture_all_gather(module):
    seen = []

    def all_gather(data):        # single process: comm.all_gather returns [data] (utils/comm.py:102-104)
        seen.append(data)
        return [data]
    module.all_gather = all_gather
    return seen


def test_reference_zeroshot_val_loop_runs_on_the_dropin_and_agrees_with_the_cpu_reference(dropin, tmp_path):
    cfg, ours, ref = _build_models(dropin, tmp_path)
    tok = _Tokenizer(COLOURS, cfg.vocab_size)
    loader = _RefcocoLoader(cfg)
    seen = _capture_all_gather(dropin.zs)
    acc_gpu = dropin.zs.val(types.SimpleNamespace(device=torch.device("cuda")), loader, ours, tok)
    pred_gpu, saved_gpu = seen[0], seen[1]
    del seen[:]
    acc_cpu = dropin.zs.val(types.SimpleNamespace(device=torch.device("cpu")), loader, ref, tok)
    pred_cpu, saved_cpu = seen[0], seen[1]
    assert set(pred_gpu) == set(pred_cpu) and len(pred_gpu) == 12
    decided, agree = 0, 0
    for k in pred_cpu:
        s_cpu = torch.tensor(saved_cpu[k]["scores"])
        s_gpu = torch.tensor(saved_gpu[k]["scores"])
        # the scores the loop gathered: 16-bit operands against the fp32 reference, relative to the row's largest score
        assert float((s_gpu - s_cpu).abs().max()) <= 1e-3 * float(s_cpu.abs().max()) + 1e-3
        top = s_cpu.sort(descending=True).values
        clear = len(top) < 2 or float(top[0] - top[1]) > 2e-3 * float(s_cpu.abs().max()) + 2e-3
        if clear:                  # a decision the CPU reference makes by more than the 16-bit margin must be ours too
            decided += 1
            agree += int(pred_gpu[k] == pred_cpu[k] and saved_gpu[k]["max_idx"] == saved_cpu[k]["max_idx"])
    assert decided >= 8 and agree == decided
    if decided == len(pred_cpu):
        assert acc_gpu == acc_cpu


class _GqaDataset(torch.utils.data.Dataset):
    """What GQADataset.tensorize_example returns (gqa_cpt.py:198-205) + the attributes evaluate() reads."""

    def __init__(self, cfg, n=20, n_labels=40, T=70, R=50, seed=3):
        from cpt_b200.synthetic import synth_batch
        self.labels = ["answer%d" % i for i in range(n_labels)]
        sb = synth_batch(cfg, n, T=T, R=R, seed=seed)
        g = torch.Generator().manual_seed(seed)
        self.sb, self.n = sb, n
        self.gt = torch.randint(0, n_labels, (n,), generator=g)
        self.eval_dic = {str(1000 + i): [int(self.gt[i])] for i in range(n)}

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        sb = self.sb
        return (sb["input_ids"][i], sb["attention_mask"][i], sb["token_type_ids"][i], self.gt[i:i + 1],
                torch.zeros(len(self.labels)), sb["img_feats"][i], torch.tensor([1000 + i]),
                (sb["input_ids"][i] == 103).nonzero(as_tuple=False).squeeze(1).tolist())


def test_reference_gqa_evaluate_loop_runs_on_the_dropin_and_agrees_with_the_cpu_reference(dropin, tmp_path):
    cfg, ours, ref = _build_models(dropin, tmp_path)
    ds = _GqaDataset(cfg)
    tok = _Tokenizer(ds.labels, cfg.vocab_size)

    def args(device):
        return types.SimpleNamespace(task_name="gqa", output_dir=str(tmp_path / "out"), local_rank=-1, n_gpu=1,
                                     per_gpu_eval_batch_size=8, workers=0, device=device, model_type="bert",
                                     img_feature_dim=cfg.img_feature_dim, result_dir=str(tmp_path / "rst"))
    res_gpu = dropin.gq.evaluate(args(torch.device("cuda")), ours, eval_dataset=ds, tokenizer=tok)
    res_cpu = dropin.gq.evaluate(args(torch.device("cpu")), ref, eval_dataset=ds, tokenizer=tok)
    assert len(res_gpu) == len(res_cpu) == len(ds)
    decided = 0
    for a, b in zip(res_gpu, res_cpu):
        assert a["question_id"] == b["question_id"]
        la, lb = torch.from_numpy(a["logits"]), torch.from_numpy(b["logits"])
        scale = float(lb.abs().max())
        assert float((la - lb).abs().max()) <= 1e-3 * scale + 1e-3
        top = lb.sort(descending=True).values
        if float(top[0] - top[1]) > 2e-3 * scale + 2e-3:
            decided += 1
            assert a["answer"] == b["answer"] and a["correct"] == b["correct"]
    assert decided >= len(ds) // 2
    assert os.path.isfile(str(tmp_path / "rst" / "val_results.pk"))   # the loop's own side effect (gqa_cpt.py:628-630)
