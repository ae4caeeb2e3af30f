"""Host-side engine: owns one C-ABI handle per (model, device), feeds it torch tensors' device pointers and the
current CUDA stream.  PyTorch is plumbing here (device memory, streams); all arithmetic happens in
cpt_b200/csrc.  No CPU path exists: tensors must live on a CUDA device of compute capability 10.x.
"""
import ctypes as C
import os
import weakref

import torch

from . import _lib
from ._lib import CptError  # noqa: F401  (re-exported)

DTYPES = {"fp16": 0, "float16": 0, "half": 0, "bf16": 1, "bfloat16": 1}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_tensor(name, t, dtype, device, shape=None):
    if t.device != device:
        raise CptError("cpt_b200: %s is on %s but the model is on %s" % (name, t.device, device))
    if t.dtype != dtype:
        raise CptError("cpt_b200: %s must be %s, got %s" % (name, dtype, t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise CptError("cpt_b200: %s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t.contiguous()


def _numel(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def layer_keys(i):
    p = "bert.encoder.layer.%d." % i
    return {"q_w": p + "attention.self.query.weight", "q_b": p + "attention.self.query.bias",
            "k_w": p + "attention.self.key.weight", "k_b": p + "attention.self.key.bias",
            "v_w": p + "attention.self.value.weight", "v_b": p + "attention.self.value.bias",
            "ao_w": p + "attention.output.dense.weight", "ao_b": p + "attention.output.dense.bias",
            "ao_ln_g": p + "attention.output.LayerNorm.weight", "ao_ln_b": p + "attention.output.LayerNorm.bias",
            "i_w": p + "intermediate.dense.weight", "i_b": p + "intermediate.dense.bias",
            "o_w": p + "output.dense.weight", "o_b": p + "output.dense.bias",
            "o_ln_g": p + "output.LayerNorm.weight", "o_ln_b": p + "output.LayerNorm.bias"}


GLOBAL_KEYS = {"word_emb": "bert.embeddings.word_embeddings.weight",
               "pos_emb": "bert.embeddings.position_embeddings.weight",
               "type_emb": "bert.embeddings.token_type_embeddings.weight",
               "emb_ln_g": "bert.embeddings.LayerNorm.weight", "emb_ln_b": "bert.embeddings.LayerNorm.bias",
               "img_w": "bert.img_embedding.weight", "img_b": "bert.img_embedding.bias",
               "img_ln_g": "bert.LayerNorm.weight", "img_ln_b": "bert.LayerNorm.bias",
               "pooler_w": "bert.pooler.dense.weight", "pooler_b": "bert.pooler.dense.bias",
               "mlm_dense_w": "cls.predictions.transform.dense.weight",
               "mlm_dense_b": "cls.predictions.transform.dense.bias",
               "mlm_ln_g": "cls.predictions.transform.LayerNorm.weight",
               "mlm_ln_b": "cls.predictions.transform.LayerNorm.bias",
               "mlm_bias": "cls.predictions.bias",
               "nsp_w": "cls.seq_relationship.weight", "nsp_b": "cls.seq_relationship.bias"}
OPTIONAL = {"img_w", "img_b", "img_ln_g", "img_ln_b", "pooler_w", "pooler_b", "mlm_dense_w", "mlm_dense_b",
            "mlm_ln_g", "mlm_ln_b", "mlm_bias", "nsp_w", "nsp_b"}


class Engine(object):
    """One handle on one device.  `load_state_dict` takes tensors keyed like BertImgForPreTraining's
    state_dict; fp32, on `device`."""

    def __init__(self, cfg, device, dtype="fp16", train=False):
        self.lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise CptError("cpt_b200 runs on a CUDA device (sm_100a) only; got device '%s' — no CPU path exists"
                           % device)
        if not torch.cuda.is_available():
            raise CptError("cpt_b200: no CUDA device is available")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.cfg = cfg
        self.dtype = dtype
        c = _lib.Config(cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size,
                        cfg.vocab_size, cfg.max_position_embeddings, cfg.type_vocab_size,
                        getattr(cfg, "img_feature_dim", 2054), int(getattr(cfg, "use_img_layernorm", 0) or 0),
                        int(getattr(cfg, "num_contrast_classes", 2)), float(cfg.layer_norm_eps),
                        float(getattr(cfg, "img_layer_norm_eps", cfg.layer_norm_eps)), DTYPES[dtype])
        h = C.c_void_p()
        _lib.check(self.lib.cpt_create(C.byref(c), device.index, C.byref(h)))
        self._h = h
        self._ws = None
        self.train = bool(train)
        self.weights_version = 0  # bumped by load_state_dict; a training tape is only valid for the version it saw
        if train:
            _lib.check(self.lib.cpt_train_enable(h, 1))
        self._keep = None  # tensors the handle references in place (embedding tables)
        # CUDA-graph cache of the fused encoder+head call (one graph per distinct set of input buffers): the 91
        # launches of a forward cost ~1.8 ms of host time enqueued one by one, 3 us replayed
        self.use_graphs = os.environ.get("CPT_B200_GRAPHS", "1") != "0"
        self._graphs, self._seen, self._replayed_launches = {}, {}, 0
        self._sgraphs, self._sseen, self.graph_replays = {}, {}, 0  # shape-keyed graphs over staging buffers
        self.use_train_graphs = os.environ.get("CPT_B200_TRAIN_GRAPHS", "1") != "0"
        self._tgraphs, self._tseen, self._ptr_sig = {}, {}, None
        self.grad_sync_group = None  # set by comm.enable_overlapped_grad_sync: all-reduce gradients inside the backward
        self.grad_sync_dtype = None  # None: exchange fp32; torch.bfloat16 / torch.float16: exchange 16-bit copies
        self.grad_sync_skip = False  # True inside no_sync(): gradients stay local, their sum is owed to the next exchange
        self.grad_sync_graphs = os.environ.get("CPT_B200_SYNC_GRAPHS", "1") != "0"  # NCCL inside the captured backward
        self._unsynced = None        # (layout, fp32 slab): local gradient sums of the passes run under no_sync()
        self._sync16 = None
        self._progress_cb = None
        self._profiling = False

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cpt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        dev = self.device
        keep = []

        def get(key, optional):
            t = sd.get(key)
            if t is None:
                if optional:
                    return None
                raise CptError("cpt_b200: state_dict is missing '%s'" % key)
            t = t.detach()
            if t.device != dev or t.dtype != torch.float32:
                raise CptError("cpt_b200: '%s' must be a float32 tensor on %s (got %s on %s)"
                               % (key, dev, t.dtype, t.device))
            t = t.contiguous()
            keep.append(t)
            return t

        w = _lib.Weights()
        for f, key in GLOBAL_KEYS.items():
            setattr(w, f, _ptr(get(key, f in OPTIONAL)))
        L = self.cfg.num_hidden_layers
        layers = (_lib.LayerWeights * max(L, 1))()
        for i in range(L):
            for f, key in layer_keys(i).items():
                setattr(layers[i], f, _ptr(get(key, False)))
        w.layers = C.cast(layers, C.POINTER(_lib.LayerWeights))
        self._graphs.clear()  # captured graphs hold the old 16-bit weight buffers
        self._seen.clear()
        for st in self._sgraphs.values():  # staging buffers stay, their graphs are re-captured
            st["graph"] = None
        ptr_sig = tuple(t.data_ptr() for t in keep)
        if ptr_sig != self._ptr_sig:  # a tensor moved: the handle may reallocate, captured training graphs are stale
            self._tgraphs.clear()
            self._tseen.clear()
            self._ptr_sig = ptr_sig
        with torch.cuda.device(dev):
            _lib.check(self.lib.cpt_set_weights(self._h, C.byref(w), _stream()))
        self.weights_version += 1
        # only the embedding tables are referenced in place by the handle; keep them alive
        self._keep = [sd[GLOBAL_KEYS[k]] for k in ("word_emb", "pos_emb", "type_emb")]

    # ------------------------------------------------------------------ forward
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    def encoder_forward(self, input_ids, token_type_ids=None, attention_mask=None, position_ids=None,
                        img_feats=None, want_pooled=True, want_hidden=False):
        dev = self.device
        if input_ids.dim() != 2:
            raise CptError("cpt_b200: input_ids must be [B,T]")
        B, T = input_ids.shape
        R = 0 if img_feats is None else img_feats.shape[1]
        S = T + R
        H = self.cfg.hidden_size
        i64 = torch.int64
        ids = _chk_tensor("input_ids", input_ids, i64, dev)
        seg = None if token_type_ids is None else _chk_tensor("token_type_ids", token_type_ids, i64, dev, (B, T))
        msk = None if attention_mask is None else _chk_tensor("attention_mask", attention_mask, i64, dev, (B, S))
        pos = None
        if position_ids is not None:
            pos = _chk_tensor("position_ids", position_ids.expand(B, T) if position_ids.dim() == 2 else position_ids,
                              i64, dev, (B, T))
        img = None
        if img_feats is not None:
            img = _chk_tensor("img_feats", img_feats, torch.float32, dev, (B, R, self.cfg.img_feature_dim))
        with torch.cuda.device(dev):
            seq = torch.empty(B, S, H, dtype=torch.float32, device=dev)
            pooled = torch.empty(B, H, dtype=torch.float32, device=dev) if want_pooled else None
            hidden = (torch.empty(self.cfg.num_hidden_layers + 1, B, S, H, dtype=torch.float32, device=dev)
                      if want_hidden else None)
            nbytes = self.lib.cpt_workspace_bytes(self._h, B, T, R)
            ws = self._workspace(nbytes)
            _lib.check(self.lib.cpt_encoder_forward(self._h, _stream(), _ptr(ids), _ptr(seg), _ptr(msk), _ptr(pos),
                                                    _ptr(img), B, T, R, _ptr(ws), ws.numel(), _ptr(seq),
                                                    _ptr(pooled), _ptr(hidden)))
        return seq, pooled, hidden

    # ------------------------------------------------------------------ training step (MLM loss)
    def _train_inputs(self, input_ids, token_type_ids, attention_mask, position_ids, img_feats):
        dev, i64 = self.device, torch.int64
        if input_ids.dim() != 2:
            raise CptError("cpt_b200: input_ids must be [B,T]")
        B, T = input_ids.shape
        R = 0 if img_feats is None else img_feats.shape[1]
        ids = _chk_tensor("input_ids", input_ids, i64, dev)
        seg = None if token_type_ids is None else _chk_tensor("token_type_ids", token_type_ids, i64, dev, (B, T))
        msk = None if attention_mask is None else _chk_tensor("attention_mask", attention_mask, i64, dev, (B, T + R))
        pos = None
        if position_ids is not None:
            pos = _chk_tensor("position_ids", position_ids.expand(B, T) if position_ids.dim() == 2 else position_ids,
                              i64, dev, (B, T))
        img = None
        if img_feats is not None:
            img = _chk_tensor("img_feats", img_feats, torch.float32, dev, (B, R, self.cfg.img_feature_dim))
        return B, T, R, ids, seg, msk, pos, img

    def _grads_struct(self, head, grads):
        dev = self.device
        keep = []

        def gp(key, optional=False):
            t = grads.get(key)
            if t is None:
                if optional:
                    return C.c_void_p(0)
                raise CptError("cpt_b200: gradient buffer for '%s' is missing" % key)
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise CptError("cpt_b200: gradient buffer '%s' must be contiguous float32 on %s" % (key, dev))
            keep.append(t)
            return _ptr(t)

        g = _lib.Grads()
        unused = ("pooler_w", "pooler_b", "nsp_w", "nsp_b") if head == "mlm" else \
            ("mlm_dense_w", "mlm_dense_b", "mlm_ln_g", "mlm_ln_b", "mlm_bias")
        for f in _lib.GRAD_GLOBAL_FIELDS:
            setattr(g, f, gp(GLOBAL_KEYS[f], f.startswith("img_") or f in unused))
        L = self.cfg.num_hidden_layers
        layers = (_lib.LayerGrads * max(L, 1))()
        for i in range(L):
            for f, key in layer_keys(i).items():
                setattr(layers[i], f, gp(key))
        g.layers = C.cast(layers, C.POINTER(_lib.LayerGrads))
        return g, (layers, keep)

    def _raw_forward(self, s):
        fwd = self.lib.cpt_train_forward_mlm if s["head"] == "mlm" else self.lib.cpt_train_forward_nsp
        _lib.check(fwd(self._h, _stream(), _ptr(s["ids"]), _ptr(s["seg"]), _ptr(s["msk"]), _ptr(s["pos"]),
                       _ptr(s["img"]), s["B"], s["T"], s["R"], _ptr(s["rows"]), _ptr(s["targets"]), s["n"],
                       C.byref(s["drop"]), _ptr(s["tape"]), s["tape"].numel(), _ptr(s["loss"])))

    def _raw_backward(self, s, gl, g):
        bwd = self.lib.cpt_train_backward_mlm if s["head"] == "mlm" else self.lib.cpt_train_backward_nsp
        _lib.check(bwd(self._h, _stream(), _ptr(s["ids"]), _ptr(s["seg"]), _ptr(s["pos"]), s["B"], s["T"], s["R"],
                       _ptr(s["rows"]), _ptr(s["targets"]), s["n"], C.byref(s["drop"]), _ptr(gl), _ptr(s["tape"]),
                       s["tape"].numel(), C.byref(g)))

    def train_forward(self, head, input_ids, token_type_ids, attention_mask, position_ids, img_feats, rows, targets,
                      dropout=None):
        """head "mlm": loss of REC_MLM_CPT.forward(masked_lm_labels=...) (modeling_rec.py:146-149) at the labelled
        positions `rows` (flat b*S+s indices) with labels `targets`; head "nsp": loss of
        NSPCPT.forward(next_sentence_label=...) (modeling_vcr.py:120-127), rows = b*S of the labelled samples.
        dropout = (p_hidden, p_attn, seed) or None.  Returns (loss, saved) — `saved` feeds train_backward.

        A step is ~400 small launches; for few-shot batches the launch sequence, not the arithmetic, sets the time.
        The second time a step SHAPE (head, B, T, R, number of labelled rows, which optional inputs are present,
        dropout probabilities) is seen, the forward — and then its backward — are captured into CUDA graphs over
        persistent staging buffers and replayed from then on (inputs are copied into the staging buffers, the
        dropout seed is read from device memory).  CPT_B200_TRAIN_GRAPHS=0 turns this off."""
        if not self.train:
            raise CptError("cpt_b200: this engine was not created with train=True")
        B, T, R, ids, seg, msk, pos, img = self._train_inputs(input_ids, token_type_ids, attention_mask,
                                                               position_ids, img_feats)
        dev = self.device
        rows = _chk_tensor("rows", rows, torch.int64, dev)
        targets = _chk_tensor("targets", targets, torch.int64, dev, tuple(rows.shape))
        n = int(rows.numel())
        p_h, p_a, seed = (0.0, 0.0, 0) if dropout is None else (float(dropout[0]), float(dropout[1]), int(dropout[2]))
        key = (head, B, T, R, n, seg is None, msk is None, pos is None, img is None, p_h, p_a)
        st = None
        if (self.use_train_graphs and not self._profiling and (self.grad_sync_group is None or self.grad_sync_graphs)
                and not torch.cuda.is_current_stream_capturing()):
            st = self._tgraphs.get(key)
            if st is None:
                if len(self._tseen) > 4096:  # ever-changing shapes: do not let the sightings table grow without bound
                    self._tseen.clear()
                c = self._tseen.get(key, 0) + 1
                self._tseen[key] = c
                if c >= 2 and len(self._tgraphs) < 8:
                    st = self._new_train_state(key, head, B, T, R, n, seg, msk, pos, img, p_h, p_a)
                    self._tgraphs[key] = st
            if st is not None and st["pending"]:
                st["pending"] = False  # an un-backpropagated forward still owns the tape: run this one eagerly
                st = None
        if st is None:
            with torch.cuda.device(dev):
                nbytes = self.lib.cpt_train_tape_bytes(self._h, B, T, R, n)
                s = dict(head=head, B=B, T=T, R=R, n=n, ids=ids, seg=seg, msk=msk, pos=pos, img=img, rows=rows,
                         targets=targets, drop=_lib.Dropout(p_h, p_a, seed & 0xFFFFFFFFFFFFFFFF, None),
                         tape=torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=dev),
                         loss=torch.empty((), dtype=torch.float32, device=dev), version=self.weights_version,
                         graph=None)
                self._raw_forward(s)
            return s["loss"], s
        # graph path: refresh the staging buffers, replay
        with torch.cuda.device(dev):
            for name, t in (("ids", ids), ("seg", seg), ("msk", msk), ("pos", pos), ("img", img), ("rows", rows),
                            ("targets", targets)):
                if t is not None:
                    st[name].copy_(t, non_blocking=True)
            st["seed"].fill_(seed)
            if st["fwd"] is None:
                # warm-up run + captured run are both counted by the library, and exactly two runs execute
                l0 = self.lib.cpt_launch_count(self._h)
                st["fwd"] = self._capture(lambda: self._raw_forward(st))
                st["fwd_launches"] = int(self.lib.cpt_launch_count(self._h) - l0) // 2
            else:
                st["fwd"].replay()
                self._replayed_launches += st["fwd_launches"]
            st["pending"] = True
            st["gen"] += 1
            saved = dict(graph=st, gen=st["gen"], head=head, version=self.weights_version)
            return st["loss"].clone(), saved

    def _capture(self, fn):
        """Capture fn's launches on a side stream into a CUDA graph (fn runs once, un-captured, first: lazy one-time
        initialisation inside the library must not land in the graph), replay it once, return it."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        fn()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                fn()
        cur.wait_stream(side)
        g.replay()
        return g

    def _new_train_state(self, key, head, B, T, R, n, seg, msk, pos, img, p_h, p_a):
        dev, i64 = self.device, torch.int64
        with torch.cuda.device(dev):
            st = dict(head=head, B=B, T=T, R=R, n=n, pending=False, gen=0, fwd=None, bwd=None, slab=None, layout=None,
                      ids=torch.zeros(B, T, dtype=i64, device=dev),
                      seg=None if seg is None else torch.zeros(B, T, dtype=i64, device=dev),
                      msk=None if msk is None else torch.zeros(B, T + R, dtype=i64, device=dev),
                      pos=None if pos is None else torch.zeros(B, T, dtype=i64, device=dev),
                      img=None if img is None else torch.zeros(B, R, self.cfg.img_feature_dim, device=dev),
                      rows=torch.zeros(n, dtype=i64, device=dev), targets=torch.zeros(n, dtype=i64, device=dev),
                      seed=torch.zeros(1, dtype=i64, device=dev), loss=torch.zeros((), device=dev),
                      grad_loss=torch.ones((), device=dev))
            st["drop"] = _lib.Dropout(p_h, p_a, 0, st["seed"].data_ptr())
            nbytes = self.lib.cpt_train_tape_bytes(self._h, B, T, R, n)
            st["tape"] = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=dev)
            st["fwd_launches"] = st["bwd_launches"] = 0
        return st

    def grad_buffers(self, saved, keys, shapes):
        """fp32 gradient buffers for `keys` (zero-filled): views of one slab.  On the graph path the slab is persistent
        (its addresses are baked into the captured backward) and train_backward hands back a copy."""
        sizes = [(_numel(s) + 63) // 64 * 64 for s in shapes]
        st = saved.get("graph")
        if st is not None and st["slab"] is not None and st["layout"] == (tuple(keys), tuple(shapes)):
            slab = st["slab"]
        else:
            slab = torch.zeros(sum(sizes), dtype=torch.float32, device=self.device)
            if st is not None:
                st["slab"], st["layout"], st["bwd"], st["bwd_sync"] = slab, (tuple(keys), tuple(shapes)), None, None
        return self._views(slab, keys, shapes, sizes)

    @staticmethod
    def _views(slab, keys, shapes, sizes):
        out, off = {}, 0
        for k, s, n in zip(keys, shapes, sizes):
            out[k] = slab[off:off + _numel(s)].view(s)
            off += n
        return out

    def train_backward(self, saved, grad_loss, grads):
        """Adds d(grad_loss * loss)/d(param) into `grads` (from grad_buffers).  Returns the dict holding the results
        (on the graph path a copy of the persistent slab, so that the caller owns what it accumulates)."""
        if saved["version"] != self.weights_version:
            raise CptError("cpt_b200: weights were reloaded between the training forward and its backward")
        dev = self.device
        gl = _chk_tensor("grad_loss", grad_loss.reshape(()), torch.float32, dev)
        st = saved.get("graph")
        sync = self.grad_sync_group is not None and bool(saved.get("groups"))
        if st is None:
            g, keep = self._grads_struct(saved["head"], grads)
            with torch.cuda.device(dev):
                if sync and not self.grad_sync_skip and self._unsynced is None:
                    self._backward_with_grad_sync(saved, gl, g, grads, saved["groups"])
                else:
                    self._raw_backward(saved, gl, g)
                    if sync:
                        self._settle_unsynced(self._slab_of(grads), tuple(grads))
            return grads
        if saved["gen"] != st["gen"]:
            raise CptError("cpt_b200: the tape of this forward was overwritten by a later forward of the same shape "
                           "before backward() ran; set CPT_B200_TRAIN_GRAPHS=0 for this usage pattern")
        with torch.cuda.device(dev):
            st["grad_loss"].copy_(gl, non_blocking=True)
            # two captured variants of the backward: local ("bwd"), and with the gradient exchange inside ("bwd_sync":
            # one NCCL all-reduce per gradient group on NCCL's stream, forked / joined inside the graph)
            in_graph = sync and not self.grad_sync_skip and self._unsynced is None
            which = "bwd_sync" if in_graph else "bwd"
            if st.get(which) is None:
                g, keep = self._grads_struct(saved["head"], grads)
                st["_keep_" + which] = (g, keep)
                groups = saved.get("groups")

                def run():
                    st["slab"].zero_()
                    if in_graph:
                        self._backward_with_grad_sync(st, st["grad_loss"], g, grads, groups)
                    else:
                        self._raw_backward(st, st["grad_loss"], g)

                l0 = self.lib.cpt_launch_count(self._h)
                st[which] = self._capture(run)
                st[which + "_launches"] = int(self.lib.cpt_launch_count(self._h) - l0) // 2
            else:
                st[which].replay()
                self._replayed_launches += st[which + "_launches"]
            if sync and not in_graph:
                self._settle_unsynced(st["slab"], tuple(grads))
            st["pending"] = False
            keys, shapes = st["layout"]
            sizes = [(_numel(s) + 63) // 64 * 64 for s in shapes]
            return self._views(st["slab"].clone(), keys, shapes, sizes)

    @staticmethod
    def _slab_of(grads):
        t = next(iter(grads.values()))
        return (t._base if t._base is not None else t).view(-1)

    def _settle_unsynced(self, slab, layout):
        """Gradient accumulation over several backward passes with the exchange only on the last one (DDP's no_sync(),
        gqa_cpt.py:441-462 with gradient_accumulation_steps > 1).  Under no_sync() the pass stays local and its
        gradients are added to `_unsynced`.  The first exchanging pass afterwards averages (this pass + everything
        owed) over the ranks and hands autograd  average - owed,  so that  .grad = owed + (average - owed)  is what
        DDP would have left there: the average of the accumulated gradients."""
        import torch.distributed as dist
        slab = slab.view(-1)
        if self.grad_sync_skip:
            if self._unsynced is None:
                self._unsynced = (layout, torch.zeros_like(slab))
            if self._unsynced[0] != layout or self._unsynced[1].numel() != slab.numel():
                raise CptError("cpt_b200: the set of trained parameters changed between backward passes under no_sync()")
            self._unsynced[1].add_(slab)
            return
        owed = self._unsynced
        self._unsynced = None
        if owed is None:
            return
        if owed[0] != layout or owed[1].numel() != slab.numel():
            raise CptError("cpt_b200: the set of trained parameters changed between backward passes under no_sync()")
        slab.add_(owed[1])
        dist.all_reduce(slab, op=dist.ReduceOp.AVG, group=self.grad_sync_group)
        slab.sub_(owed[1])

    def release_sync_graphs(self):
        """Destroy the captured backward graphs that hold NCCL operations.  NCCL keeps a communicator alive — and
        ncclCommDestroy / process exit waiting — until every CUDA graph that captured one of its collectives is gone,
        so this must run before torch.distributed.destroy_process_group() (comm.enable_overlapped_grad_sync arranges it)."""
        n = 0
        for st in self._tgraphs.values():
            if st.get("bwd_sync") is not None:
                n += 1
            st["bwd_sync"] = None
            st.pop("_keep_bwd_sync", None)
        if n:
            torch.cuda.synchronize(self.device)
        return n

    def drop_unsynced_gradients(self):
        """zero_grad() between no_sync() passes and the exchanging pass: forget what was owed."""
        self._unsynced = None

    def _backward_with_grad_sync(self, saved, gl, g, grads, groups):
        """Data-parallel backward: each gradient group (loss head, layer L-1, ..., layer 0, embeddings) is averaged over
        the ranks by an asynchronous NCCL all-reduce issued the moment its last kernel has been enqueued
        (cpt_train_set_progress_callback), so the exchange overlaps the rest of the backward instead of following it
        (the reference's DDP reducer does the same per 25 MB bucket).  Gradients of a group are one
        contiguous range of the slab (training.trainable_groups).  With `grad_sync_dtype` set, a 16-bit copy of the
        range travels instead (half the bytes; the average is rounded once to that type).  Runs eagerly or inside a
        CUDA-graph capture (NCCL's stream is forked from and joined to the capturing stream by the Work objects)."""
        import torch.distributed as dist
        group = self.grad_sync_group
        ranges = []
        for keys in groups:
            ts = [grads[k] for k in keys if k in grads]
            if not ts:
                ranges.append(None)
                continue
            base = ts[0]._base if ts[0]._base is not None else ts[0]
            lo = min(t.storage_offset() for t in ts)
            hi = max(t.storage_offset() + t.numel() for t in ts)
            ranges.append((base.view(-1)[lo:hi], lo, hi))
        d16 = self.grad_sync_dtype
        if d16 is not None:
            total = max((r[2] for r in ranges if r is not None), default=0)
            if self._sync16 is None or self._sync16.numel() < total or self._sync16.dtype != d16:
                if torch.cuda.is_current_stream_capturing():
                    raise CptError("cpt_b200: internal: the 16-bit exchange buffer must exist before capture")
                self._sync16 = torch.empty(total, dtype=d16, device=self.device)
        works, err = [], []

        def on_stage(_user, stage):
            try:
                if stage < len(ranges) and ranges[stage] is not None:
                    r, lo, hi = ranges[stage]
                    if d16 is None:
                        works.append((dist.all_reduce(r, op=dist.ReduceOp.AVG, group=group, async_op=True), None, None))
                    else:
                        b = self._sync16[lo:hi]
                        b.copy_(r)
                        works.append((dist.all_reduce(b, op=dist.ReduceOp.AVG, group=group, async_op=True), r, b))
            except Exception as e:  # never let an exception cross the C frame
                err.append(e)

        cb = _lib.PROGRESS_FN(on_stage)
        _lib.check(self.lib.cpt_train_set_progress_callback(self._h, cb, None))
        try:
            self._raw_backward(saved, gl, g)
        finally:
            _lib.check(self.lib.cpt_train_set_progress_callback(self._h, None, None))
        if err:
            raise err[0]
        for w, r, b in works:
            w.wait()  # stream-level wait: the current stream sees the averaged gradients
            if r is not None:
                r.copy_(b)

    def mlm_gather(self, seq_out, mask_pos, vocab_ids=None, exchange=None):
        """logits[b, k] = scores[b, mask_pos[b], vocab_ids[k]] (gather-first MLM head).  exchange (a
        comm.LogitsExchange): the decoder kernel stores into every rank's gather buffer instead and the result is the
        gathered [world * B, K] block."""
        dev = self.device
        B, S, H = seq_out.shape
        seq = _chk_tensor("sequence_output", seq_out, torch.float32, dev)
        mp = _chk_tensor("mask_pos", mask_pos, torch.int64, dev, (B,))
        if vocab_ids is None:
            K, vid = self.cfg.vocab_size, None
        else:
            vid = _chk_tensor("vocab_ids", vocab_ids, torch.int64, dev)
            K = vid.numel()
        with torch.cuda.device(dev):
            ws = self._workspace(B * H * 4 + 512)
            if exchange is not None:
                out = torch.empty(exchange.world * B, K, dtype=torch.float32, device=dev)
                _lib.check(self.lib.cpt_mlm_gather_exchange(self._h, exchange._ex, _stream(), _ptr(seq), B, S, _ptr(mp),
                                                            _ptr(vid), K, _ptr(ws), ws.numel(), _ptr(out)))
                return out
            out = torch.empty(B, K, dtype=torch.float32, device=dev)
            _lib.check(self.lib.cpt_mlm_gather_forward(self._h, _stream(), _ptr(seq), B, S, _ptr(mp), _ptr(vid), K,
                                                       _ptr(ws), ws.numel(), _ptr(out)))
        return out

    def cpt_logits(self, input_ids, token_type_ids, attention_mask, position_ids, img_feats, mask_pos, vocab_ids,
                   exchange=None):
        """encoder + gathered MLM head in one call: logits[b,k] = scores[b, mask_pos[b], vocab_ids[k]].

        The launch sequence is replayed from CUDA graphs in two ways:
          * the same input BUFFERS (addresses + shapes) seen twice -> a graph over those buffers, zero-copy (a caller that
            keeps its own device staging buffers, like bench.py);
          * otherwise, the second time an input SHAPE is seen -> a graph over persistent staging buffers owned by the
            engine; every call copies its (fresh) input tensors into them and replays.  This is what the reference's
            loop hits: it moves every batch to the device anew (Oscar/oscar/zeroshot/refcoco_cpt.py:212-219), so
            addresses never repeat — without the staging graph each forward paid ~1.8 ms of host enqueue time for
            ~40 launches' worth of GPU work."""
        if exchange is not None and getattr(exchange, "_ex", None) is None:
            raise CptError("cpt_b200: this LogitsExchange has been closed")

        def run(t):
            seq, _, _ = self.encoder_forward(t[0], t[1], t[2], t[3], t[4], want_pooled=False)
            return self.mlm_gather(seq, t[5], t[6], exchange)

        return self._graphed("mlm" if exchange is None else ("mlm", id(exchange)), (input_ids, token_type_ids, attention_mask, position_ids, img_feats, mask_pos,
                                     vocab_ids), run)

    def nsp_scores(self, input_ids, token_type_ids, attention_mask, position_ids, img_feats):
        """encoder + pooler + seq_relationship head in one (graph-replayed) call: NSPCPT.forward without labels
        (Oscar/oscar/modeling/modeling_vcr.py:115-123), the VCR inference step (fewshot/vcr_nsp_cpt.py:597-600)."""
        def run(t):
            _, pooled, _ = self.encoder_forward(t[0], t[1], t[2], t[3], t[4], want_pooled=True)
            return self.nsp(pooled)

        return self._graphed("nsp", (input_ids, token_type_ids, attention_mask, position_ids, img_feats), run)

    def _graphed(self, kind, ts, run):
        """Run `run(ts)` (a fixed launch sequence over the tensors `ts`), replaying CUDA graphs as described in
        cpt_logits."""
        if not self.use_graphs or self._profiling or torch.cuda.is_current_stream_capturing():
            return run(ts)
        if any(t is not None and not t.is_contiguous() for t in ts):
            return run(ts)
        key = (kind,) + tuple((t.data_ptr(), tuple(t.shape), t.dtype) if t is not None else None for t in ts)
        hit = self._graphs.get(key)
        if hit is not None:
            hit[0].replay()
            self._replayed_launches += hit[2]
            self.graph_replays += 1
            return hit[1].clone()
        # a pointer-keyed graph is for PERSISTENT input buffers (the caller's own staging tensors).  Fresh tensors get
        # recycled addresses from the caching allocator, so the same key can recur without any buffer being reused:
        # count a sighting only if the first tensor is the very object seen before.
        first = next((t for t in ts if t is not None), None)
        seen = self._seen.get(key)
        if seen is not None and seen[1]() is first:
            seen[0] += 1
        else:
            seen = [1, weakref.ref(first)]
            self._seen[key] = seen
        n = seen[0]
        if len(self._seen) > 4096:
            self._seen.clear()
        if n >= 2 and len(self._graphs) < 64:
            g, out, n_launch = self._capture_counted(lambda: run(ts))
            self._graphs[key] = (g, out, n_launch, ts)  # ts keeps the input buffers (and their addresses) alive
            return out.clone()
        # fresh buffers: the shape-keyed graph over the engine's staging buffers
        skey = (kind,) + tuple((tuple(t.shape), t.dtype) if t is not None else None for t in ts)
        st = self._sgraphs.get(skey)
        if st is None:
            c = self._sseen.get(skey, 0) + 1
            self._sseen[skey] = c
            if len(self._sseen) > 4096:
                self._sseen.clear()
            if c < 2 or len(self._sgraphs) >= 16:
                return run(ts)
            with torch.cuda.device(self.device):
                bufs = tuple(None if t is None else torch.empty_like(t) for t in ts)
            st = dict(bufs=bufs, graph=None, out=None, launches=0)
            self._sgraphs[skey] = st
        for buf, t in zip(st["bufs"], ts):
            if buf is not None:
                buf.copy_(t, non_blocking=True)
        if st["graph"] is None:
            bufs = st["bufs"]
            st["graph"], st["out"], st["launches"] = self._capture_counted(lambda: run(bufs))
        else:
            st["graph"].replay()
            self._replayed_launches += st["launches"]
            self.graph_replays += 1
        return st["out"].clone()

    def _capture_counted(self, fn):
        """Capture fn (which must have run eagerly before for this shape: schedules and lazy one-time state are built
        outside capture) into a graph on a side stream, replay it once; returns (graph, output, launches per replay)."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        l0 = self.lib.cpt_launch_count(self._h)
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                out = fn()
        cur.wait_stream(side)
        n_launch = int(self.lib.cpt_launch_count(self._h) - l0)
        g.replay()
        self._replayed_launches += n_launch
        self.graph_replays += 1
        return g, out, n_launch

    def mlm_scores(self, seq_out):
        dev = self.device
        seq = _chk_tensor("sequence_output", seq_out, torch.float32, dev)
        rows = seq.numel() // seq.shape[-1]
        with torch.cuda.device(dev):
            out = torch.empty(tuple(seq.shape[:-1]) + (self.cfg.vocab_size,), dtype=torch.float32, device=dev)
            # the encoder workspace may still be in flight on this stream: that is fine, same-stream order
            nbytes = self.lib.cpt_mlm_scores_workspace_bytes(self._h, rows)
            ws = self._workspace(nbytes)
            _lib.check(self.lib.cpt_mlm_scores_forward(self._h, _stream(), _ptr(seq), rows, _ptr(ws), ws.numel(),
                                                       _ptr(out)))
        return out

    def nsp(self, pooled):
        dev = self.device
        p = _chk_tensor("pooled_output", pooled, torch.float32, dev)
        B = p.shape[0]
        with torch.cuda.device(dev):
            out = torch.empty(B, int(getattr(self.cfg, "num_contrast_classes", 2)), dtype=torch.float32, device=dev)
            _lib.check(self.lib.cpt_nsp_forward(self._h, _stream(), _ptr(p), B, _ptr(out)))
        return out

    def head_linear(self, pooled, weight, bias):
        """out[B,C] = pooled W^T + b with caller-held fp32 weights (VCRQAR_NSPCPT's per-call head choice)."""
        dev = self.device
        p = _chk_tensor("pooled_output", pooled, torch.float32, dev)
        w = _chk_tensor("head weight", weight.detach(), torch.float32, dev, (weight.shape[0], self.cfg.hidden_size))
        b = None if bias is None else _chk_tensor("head bias", bias.detach(), torch.float32, dev, (weight.shape[0],))
        with torch.cuda.device(dev):
            out = torch.empty(p.shape[0], w.shape[0], dtype=torch.float32, device=dev)
            _lib.check(self.lib.cpt_head_linear(self._h, _stream(), _ptr(p), p.shape[0], _ptr(w), _ptr(b), w.shape[0],
                                                _ptr(out)))
        return out

    def check(self):
        """Synchronise and surface device-side input errors (out-of-range ids, the reference's IndexError)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cpt_check_async_error(self._h, _stream()))

    # ------------------------------------------------------------------ launch accounting / profiling
    def launch_count(self):
        """kernels launched by this engine: counted by the library, plus graph replays x launches per graph"""
        return int(self.lib.cpt_launch_count(self._h)) + self._replayed_launches

    def profile(self, on=True):
        self._profiling = bool(on)  # graph replays carry no per-kernel events: profile the eager launch sequence
        _lib.check(self.lib.cpt_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kernel class: (device ms, launches)} accumulated since profile(True) / the previous read."""
        ms = (C.c_double * _lib.K_COUNT)()
        n = (C.c_longlong * _lib.K_COUNT)()
        _lib.check(self.lib.cpt_profile_read(self._h, ms, n))
        return {self.lib.cpt_kernel_name(i).decode(): (ms[i], int(n[i])) for i in range(_lib.K_COUNT) if n[i]}

    # ------------------------------------------------------------------ kernel-level hooks (tests / bench)
    def _t16(self):
        return torch.float16 if DTYPES[self.dtype] == 0 else torch.bfloat16

    def gemm(self, A, W, bias=None, resid=None, epi=0, out_fp32=False, tile_cfg=0, trans=False, accumulate_into=None,
             ksplit=0):
        """out = epi(A . W^T); trans: A is [K,M], W is [K,N] and out = A^T . W; accumulate_into: fp32 [M,N] tensor
        the result is added to (returned)."""
        if trans == "b":  # out = A . W with W given as [K, N]
            M, K = A.shape
            N = W.shape[1]
            epi |= 0x400
        elif trans:
            K, M = A.shape
            N = W.shape[1]
            epi |= 0x100
        else:
            M, K = A.shape
            N = W.shape[0]
        if accumulate_into is not None:
            out, out_fp32 = accumulate_into, True
            epi |= 0x200 | (int(ksplit) << 12)
        else:
            out = torch.empty(M, N, dtype=torch.float32 if out_fp32 else self._t16(), device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cpt_gemm(self._h, _stream(), _ptr(A), A.stride(0), _ptr(W), W.stride(0), M, N, K,
                                         _ptr(bias), _ptr(resid), 0 if resid is None else resid.stride(0), epi,
                                         1 if out_fp32 else 0, _ptr(out), out.stride(0), tile_cfg))
        return out

    def chain(self, stages):
        """Run dependent stages as ONE dataflow launch (cpt_chain_run).  `stages`: dicts with kind "gemm"
        (A, W, bias, out, gelu, accumulate, ksplit, dep) or "ln" (x, gamma, beta, eps, out32, out16, dep); `dep` is the
        index of the stage whose output the stage reads (or None)."""
        arr = (_lib.ChainStage * len(stages))()
        for i, s in enumerate(stages):
            c = arr[i]
            c.dep_stage = -1 if s.get("dep") is None else int(s["dep"])
            if s["kind"] == "ln":
                x = s["x"]
                c.kind, c.M, c.N = 1, x.shape[0], x.shape[1]
                c.ln_in, c.gamma, c.beta, c.eps = x.data_ptr(), s["gamma"].data_ptr(), s["beta"].data_ptr(), s["eps"]
                c.out32 = 0 if s.get("out32") is None else s["out32"].data_ptr()
                c.out16 = 0 if s.get("out16") is None else s["out16"].data_ptr()
            elif s.get("resid") is not None:  # dense + bias + residual (+ LayerNorm in the tile epilogue, or deferred)
                A, W, r = s["A"], s["W"], s["resid"]
                c.kind, c.M, c.K, c.N, c.ksplit = 0, A.shape[0], A.shape[1], W.shape[0], 1
                c.ln = 2 if s.get("part") is not None else 1
                c.A, c.lda, c.W, c.ldw = A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0)
                c.bias = 0 if s.get("bias") is None else s["bias"].data_ptr()
                c.resid, c.ldr = r.data_ptr(), r.stride(0)
                c.gamma = 0 if s.get("gamma") is None else s["gamma"].data_ptr()
                c.beta = 0 if s.get("beta") is None else s["beta"].data_ptr()
                c.eps = s.get("eps", 0.0)
                c.part = 0 if s.get("part") is None else s["part"].data_ptr()
                c.rpart = 0 if s.get("rpart") is None else s["rpart"].data_ptr()
                c.out32 = 0 if s.get("out32") is None else s["out32"].data_ptr()
                c.out16 = 0 if s.get("out16") is None else s["out16"].data_ptr()
            else:
                A, W, out = s["A"], s["W"], s["out"]
                c.kind, c.M, c.K, c.N = 0, A.shape[0], A.shape[1], W.shape[0]
                c.gelu, c.out_fp32, c.ksplit = int(bool(s.get("gelu"))), int(out.dtype == torch.float32), int(s.get("ksplit", 1))
                c.A, c.lda, c.W, c.ldw = A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0)
                c.bias = 0 if s.get("bias") is None else s["bias"].data_ptr()
                c.out, c.ldo = out.data_ptr(), out.stride(0)
                if s.get("apart") is not None:  # A = raw rows of a deferred LayerNorm, W = gamma-folded weight
                    c.apart, c.gvec, c.eps = s["apart"].data_ptr(), s["gvec"].data_ptr(), s["eps"]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cpt_chain_run(self._h, _stream(), arr, len(stages)))

    def chain_trace(self):
        """Event log of the last chain launch: (header [pairs,2], events [pairs,pitch,16]) as nested lists."""
        n = 1 << 22
        buf = (C.c_longlong * n)()
        pairs, pitch = C.c_int(), C.c_int()
        _lib.check(self.lib.cpt_chain_trace(self._h, buf, n, C.byref(pairs), C.byref(pitch)))
        P, W = pairs.value, pitch.value
        hdr = [list(buf[2 * i:2 * i + 2]) for i in range(P)]
        off = 2 * P
        ev = [[list(buf[off + (i * W + j) * 16:off + (i * W + j) * 16 + 16]) for j in range(W)] for i in range(P)]
        return hdr, ev

    def gemm_trace(self, n=148):
        buf = (C.c_longlong * (16 * n))()
        _lib.check(self.lib.cpt_gemm_trace(self._h, buf, n))
        return [list(buf[16 * i:16 * i + 16]) for i in range(n)]

    def attention(self, qkv, ext_mask, B, S, impl=0):
        H = self.cfg.hidden_size
        ctx = torch.empty(B * S, H, dtype=self._t16(), device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cpt_attention(self._h, _stream(), _ptr(qkv), _ptr(ext_mask), B, S, _ptr(ctx), impl))
        return ctx

    def attention_backward(self, qkv, dctx, ext_mask, B, S, impl=-1):
        dqkv = torch.zeros_like(qkv)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cpt_attention_backward(self._h, _stream(), _ptr(qkv), _ptr(dctx), _ptr(ext_mask), B, S,
                                                       _ptr(dqkv), impl))
        return dqkv

    def layernorm(self, x, gamma, beta, eps):
        M, H = x.shape
        o32 = torch.empty_like(x)
        o16 = torch.empty(M, H, dtype=self._t16(), device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cpt_layernorm(self._h, _stream(), _ptr(x), M, _ptr(gamma), _ptr(beta), eps,
                                              _ptr(o32), _ptr(o16)))
        return o32, o16
