"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/cpt_b200.h declares;
the host layer refuses to run without a CUDA device (no CPU fallback exists)."""
import ctypes
import os
import re

import pytest
import torch

from cpt_b200 import _lib
from cpt_b200 import config as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "cpt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cpt_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cpt_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    hdr = open(os.path.join(ROOT, "include", "cpt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    def body(name):
        return re.search(r"typedef struct \{([^}]*)\} %s;" % name, hdr, re.S).group(1)

    assert tuple(re.findall(r"\*(\w+)", body("cpt_layer_weights"))) == _lib.LAYER_FIELDS
    assert tuple(re.findall(r"\*(\w+)", body("cpt_weights"))) == _lib.GLOBAL_FIELDS + ("layers",)
    names = [n for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body("cpt_config"))
             for n in re.findall(r"\w+", decl)]
    assert names == [f[0] for f in _lib.Config._fields_]
    # training / optimizer structs
    assert tuple(re.findall(r"\*(\w+)", body("cpt_layer_grads"))) == _lib.LAYER_FIELDS
    assert tuple(re.findall(r"\*(\w+)", body("cpt_grads"))) == _lib.GRAD_GLOBAL_FIELDS + ("layers",)
    drop = [n for decl in re.findall(r"(?:float|uint64_t|const uint64_t)\s+([^;]+);", body("cpt_dropout"))
            for n in re.findall(r"\w+", decl)]
    assert drop == [f[0] for f in _lib.Dropout._fields_]
    assert ctypes.sizeof(_lib.Dropout) == 24 and ctypes.sizeof(_lib.Grads) == 8 * (len(_lib.GRAD_GLOBAL_FIELDS) + 1)
    from cpt_b200 import optimization as OPT
    adam = [n for decl in re.findall(r"(?:const float|float|int64_t)\s+([^;]+);", body("cpt_adam_tensor"))
            for n in re.findall(r"\w+", decl)]
    assert adam == list(OPT._TENSOR.names) and OPT._TENSOR.itemsize == 56
    chunk = [n for decl in re.findall(r"(?:int32_t|int64_t)\s+([^;]+);", body("cpt_adam_chunk"))
             for n in re.findall(r"\w+", decl)]
    assert chunk == list(OPT._CHUNK.names) and OPT._CHUNK.itemsize == 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    from cpt_b200.engine import CptError, Engine
    from cpt_b200.modeling_rec import REC_MLM_CPT
    with pytest.raises(CptError):
        Engine(C.oscar_tiny(), "cpu")
    with pytest.raises(CptError):
        Engine(C.oscar_tiny(), "cuda:0")
    m = REC_MLM_CPT(C.oscar_tiny()).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, dtype=torch.long))


def test_module_state_dict_keys_match_reference(golden_dir):
    from cpt_b200.modeling_bert import BertImgForPreTraining
    g = torch.load(os.path.join(golden_dir, "tiny_s120.pt"))
    d = dict(g["cfg"])
    v = d.pop("vocab_size")
    m = BertImgForPreTraining(C.BertConfig(v, **d))
    assert sorted(m.state_dict().keys()) == g["state_dict_keys"]
    assert m.cls.predictions.decoder.weight is m.bert.embeddings.word_embeddings.weight


def test_from_pretrained_roundtrip(tmp_path):
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    import copy
    cfg = C.oscar_tiny()
    m = BertImgForPreTraining(cfg)
    m.save_pretrained(str(tmp_path))
    sd = torch.load(os.path.join(str(tmp_path), "pytorch_model.bin"))
    legacy = {k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta"): v
              for k, v in sd.items()}
    torch.save(legacy, os.path.join(str(tmp_path), "pytorch_model.bin"))
    m2 = BertImgForPreTraining.from_pretrained(str(tmp_path), config=cfg)
    assert not m2.training
    for (k1, v1), (k2, v2) in zip(sorted(m.state_dict().items()), sorted(m2.state_dict().items())):
        assert k1 == k2 and torch.equal(v1, v2)
    assert m2.cls.predictions.decoder.weight is m2.bert.embeddings.word_embeddings.weight
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(m2)
    assert rec.bert is m2.bert and rec.cls is m2.cls.predictions
    rec2 = copy.deepcopy(rec)   # fewshot/gqa_cpt.py:384 deep-copies the model
    assert rec2.cls.decoder.weight is rec2.bert.embeddings.word_embeddings.weight
    names = [n for n, _ in rec.named_parameters()]
    assert any("LayerNorm.weight" in n for n in names) and any(n.endswith("bias") for n in names)


def test_exchange_argument_checks_need_no_gpu():
    """cpt_exchange_create validates before touching CUDA: more than 8 ranks, bad rank, empty shapes -> error string."""
    import ctypes as C
    from cpt_b200 import _lib
    lib = _lib.load()
    ex, handle = C.c_void_p(), (C.c_ubyte * 64)()
    for rank, world, rows, K in ((0, 9, 4, 2), (2, 2, 4, 2), (0, 2, 0, 2), (0, 2, 4, 0)):
        assert lib.cpt_exchange_create(0, rank, world, rows, K, C.byref(ex), handle) != 0
        assert b"cpt_exchange_create" in lib.cpt_last_error()
    assert lib.cpt_exchange_connect(None, None) != 0
    assert lib.cpt_exchange_destroy(None) == 0
