"""Multi-GPU plumbing of the CPT path: one process per GPU over torch.distributed (NCCL on the GPUs; gloo in the CPU
tests).  The path shards by ROWS — samples are independent — so there is no collective inside the
encoder; the only exchange is the final [B_local, K] logits, gathered with ONE collective instead of the
reference's two pickle-over-NCCL all_gathers of Python dicts (/root/reference/Oscar/oscar/utils/comm.py:102-142,
called from oscar/zeroshot/refcoco_cpt.py:256,262).

The small helpers keep the reference's names and meaning (comm.py:14-46) so its callers read the same.
"""
import torch
import torch.distributed as dist


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def is_main_process():
    return get_rank() == 0


def synchronize():
    if get_world_size() > 1:
        dist.barrier()


def shard_queries(fanouts, rank=None, world=None):
    """Partition queries over ranks, contiguously and balanced by ROW count, never splitting a query's fan-out rows
    (RefCOCO: one row per proposal set, zeroshot/refcoco_cpt.py:224-246; VCR: 4 answer rows, vcr_nsp_cpt.py:602-604)
    so the per-query argmax stays local.  Returns (first_query, end_query, first_row, end_row) for `rank`."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    total = sum(fanouts)
    bounds, acc, q = [0], 0, 0
    for r in range(1, world):
        target = total * r / world
        while q < len(fanouts) and acc + fanouts[q] / 2.0 <= target:
            acc += fanouts[q]
            q += 1
        bounds.append(q)
    bounds.append(len(fanouts))
    q0, q1 = bounds[rank], bounds[rank + 1]
    r0 = sum(fanouts[:q0])
    return q0, q1, r0, r0 + sum(fanouts[q0:q1])


def shard_rows(fanouts, world=None):
    """Rows held by every rank under `shard_queries` — computable locally on each rank, so the gather below needs no
    size exchange."""
    world = get_world_size() if world is None else world
    out = []
    for r in range(world):
        _, _, r0, r1 = shard_queries(fanouts, r, world)
        out.append(r1 - r0)
    return out


def all_gather_logits(local, sizes=None):
    """[B_local, K] -> [sum_r B_r, K] on every rank, rank order preserved, with ONE collective and NO device->host
    synchronisation on the fast path.

    sizes: rows held by each rank (`shard_rows(fanouts)`; every rank can compute it from the fan-outs, or pass
    [B] * world for equal shards).  Then the exchange is a single fixed-shape all_gather_into_tensor —
    asynchronous on the stream, so a replayed CUDA graph of the encoder is never serialised against it
    (the reference's two pickle gathers, Oscar/oscar/utils/comm.py:102-142, synchronise twice per call).
    sizes=None keeps the general ragged form: sizes are exchanged first (one host sync for all ranks' sizes)."""
    world = get_world_size()
    if world == 1:
        return local
    if sizes is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        got = torch.empty(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(got, n)
        sizes = got.tolist()  # the one synchronisation of the ragged form
    sizes = [int(x) for x in sizes]
    if len(sizes) != world or sizes[get_rank()] != local.shape[0]:
        raise ValueError("all_gather_logits: sizes %r do not describe this rank's %d rows" % (sizes, local.shape[0]))
    mx = max(sizes)
    tail = tuple(local.shape[1:])
    send = local.contiguous()
    if local.shape[0] != mx:  # ragged last shard: pad to the common shape
        send = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
        send[:local.shape[0]].copy_(local)
    out = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)  # caching allocator: no sync
    dist.all_gather_into_tensor(out, send)
    if all(x == mx for x in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + x] for r, x in enumerate(sizes)], 0)


class LogitsExchange(object):
    """Peer-memory all-gather of the per-rank [rows_per_rank, K] logits, fused into the kernel that computes them
    (include/cpt_b200.h, cpt_mlm_gather_exchange): every rank's decoder kernel stores its logits straight into every
    rank's gather buffer over NVLink and raises one flag per peer; a small second kernel waits for the peers and hands
    out the gathered [world * rows_per_rank, K] block.  No NCCL call, no host synchronisation, and both launches are
    part of the CUDA graph the forward is replayed from.  One node, one process per GPU, equal shards; every rank must
    make the same sequence of calls.

        ex = cpt_b200.comm.LogitsExchange(rows_per_rank=B, K=len(color_ids))          # once per shape
        logits = model(ids, seg, mask, img_feats=f, mask_pos=mp, vocab_ids=color_ids, gather=ex)[0]   # [world*B, K]

    (replaces `all_gather(predictions)` of Oscar/oscar/utils/comm.py:102-142 as used in zeroshot/refcoco_cpt.py:256
    for the scores; `rows()` does the same for rows some other kernel produced, e.g. NSP scores)."""

    def __init__(self, rows_per_rank, K, device=None, group=None):
        import ctypes as C
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("LogitsExchange needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.rows_per_rank, self.K = int(rows_per_rank), int(K)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = _lib.load()
        handle = (C.c_ubyte * 64)()
        ex = C.c_void_p()
        err = None
        try:
            _lib.check(self.lib.cpt_exchange_create(self.device.index, self.rank, self.world, self.rows_per_rank,
                                                    self.K, C.byref(ex), handle))
        except _lib.CptError as e:
            err, ex = e, C.c_void_p()
        self._ex = ex if ex.value else None
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if err is None else None, group=self.group)   # 64-byte IPC handles
        if err is None and all(h is not None for h in handles):
            try:
                _lib.check(self.lib.cpt_exchange_connect(self._ex, b"".join(handles)))
            except _lib.CptError as e:
                err = e
        elif err is None:
            err = _lib.CptError("cpt_b200: a peer could not create its exchange buffer")
        # every rank learns whether EVERY rank is connected (this all-reduce is also the barrier after which peers may
        # store: all flags are zeroed); on failure all ranks raise together, so callers can fall back to NCCL in step
        ok = torch.tensor([0 if err is not None else 1], device=self.device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            if self._ex is not None:
                self.lib.cpt_exchange_destroy(self._ex)
                self._ex = None
            raise _lib.CptError("cpt_b200: peer-memory exchange unavailable on this node (%s)"
                                % (err if err is not None else "a peer failed to map the buffers"))

    def rows(self, engine, local):
        """[rows_per_rank, K] fp32 rows of this rank -> [world * rows_per_rank, K] on every rank (rank order)."""
        from . import _lib
        from .engine import _ptr, _stream
        if tuple(local.shape) != (self.rows_per_rank, self.K) or local.dtype != torch.float32 or not local.is_contiguous():
            raise ValueError("LogitsExchange.rows: expected a contiguous float32 [%d, %d] tensor"
                             % (self.rows_per_rank, self.K))
        with torch.cuda.device(self.device):
            out = torch.empty(self.world * self.rows_per_rank, self.K, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.cpt_exchange_rows(engine._h, self._ex, _stream(), _ptr(local), self.rows_per_rank,
                                                  _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_ex", None):
            # not a collective: once this rank's last exchange has completed it has seen every peer's flag for it, so no
            # peer store into this buffer is outstanding
            torch.cuda.synchronize(self.device)
            self.lib.cpt_exchange_destroy(self._ex)
            self._ex = None


def merge_by_key(dicts):
    """De-duplicate per-rank result dicts the way the reference does after its gather (DistributedSampler pads by
    repeating samples): the same key must carry the same value on every rank (zeroshot/refcoco_cpt.py:256-260)."""
    merged = {}
    for d in dicts:
        for k, v in d.items():
            assert (k not in merged) or (merged[k] == v), "rank results disagree for key %r" % (k,)
            merged[k] = v
    return merged


def pick_per_query(logits, fanouts, mode="zsl", n_valid=None):
    """CPT decision per query from the gathered colour logits [rows, K] (last column = the "none" token) — ONE native
    launch for the whole batch (cpt_b200/scoring.py -> cpt_score_queries) instead of the reference's per-image loop:
      zsl: argmax over the rows' colour columns                        (zeroshot/refcoco_cpt.py:242-246)
      fsl: argmax of colour / none                                     (fewshot/refcoco_cpt.py:291-294)
      vcr: logits are NSP scores [rows, C]; score = 1 - softmax[:, 1]  (fewshot/vcr_nsp_cpt.py:600-604)
    n_valid: per row, the size of the row's own colour set (the reference gathers `cur_color_set + ["none"]` per row;
    the last proposal set of a query is usually shorter) — columns beyond it never win.  Returns an int32 tensor
    [n_queries] of indices into each query's concatenated scores (the reference's max_idx).  CUDA tensors only."""
    from .scoring import score_queries
    return score_queries(logits, fanouts, mode=mode, n_valid=n_valid)["pick"]


_SYNC_SLOTS = []        # weak references to the engine slots whose backward graphs hold NCCL operations
_SHUTDOWN_HOOKED = [False]


def release_sync_graphs():
    """Destroy every captured backward that holds NCCL operations (Engine.release_sync_graphs).  Runs automatically
    before torch.distributed.destroy_process_group() and at interpreter exit once enable_overlapped_grad_sync was used:
    NCCL does not let go of a communicator while a CUDA graph that captured it exists."""
    for ref in list(_SYNC_SLOTS):
        slot = ref()
        eng = getattr(slot, "train_engine", None) if slot is not None else None
        if eng is not None:
            try:
                eng.release_sync_graphs()
            except Exception:
                pass


def _register_for_shutdown(slot):
    import atexit
    import functools
    import weakref
    if not any(r() is slot for r in _SYNC_SLOTS):
        _SYNC_SLOTS.append(weakref.ref(slot))
    if _SHUTDOWN_HOOKED[0]:
        return
    _SHUTDOWN_HOOKED[0] = True
    atexit.register(release_sync_graphs)
    orig = dist.destroy_process_group

    @functools.wraps(orig)
    def destroy_process_group(*a, **k):
        release_sync_graphs()
        return orig(*a, **k)
    dist.destroy_process_group = destroy_process_group


def enable_overlapped_grad_sync(ddp_model, process_group=None, exchange_dtype="auto"):
    """Training over several GPUs: average the gradients INSIDE the native backward, group by group as
    they become final (loss head, layer L-1, ..., layer 0, embeddings), overlapping NCCL with the remaining backward
    kernels — inside the CUDA graph the backward is replayed from — and switch DistributedDataParallel's own reducer off
    (it would copy every gradient into its buckets and walk the autograd graph every step for nothing).

        model = DistributedDataParallel(model, device_ids=[rank], find_unused_parameters=True)   # reference code
        cpt_b200.comm.enable_overlapped_grad_sync(model)                                          # one extra line

    exchange_dtype: "fp32" (what DDP's all-reduce carries), "bf16" (half the bytes; the averaged gradient is rounded
    once to 8 significant bits, far below the 16-bit operand rounding already inside it), or "auto" = bf16 when the
    training handle computes in bf16, fp32 otherwise (fp16 would need the caller's loss scale to be safe).
    `with model.no_sync():` keeps working: passes inside it stay local, and the next pass outside exchanges the
    accumulated sum (engine._settle_unsynced), as DDP does for gradient accumulation (gqa_cpt.py:441-462).
    Without this call DDP's own bucketed all-reduce runs after the whole native backward (correct, not overlapped).
    Every parameter of the wrapped module must get its gradient from the native step (true for REC_MLM_CPT / NSPCPT).
    """
    import contextlib
    import torch
    module = ddp_model.module if hasattr(ddp_model, "module") else ddp_model
    bert = getattr(module, "bert", module)
    group = process_group if process_group is not None else dist.group.WORLD
    if exchange_dtype not in ("auto", "fp32", "bf16"):
        raise ValueError("exchange_dtype must be 'auto', 'fp32' or 'bf16'")
    slot = bert._slot
    slot.grad_sync_group = group
    slot.grad_sync_dtype = exchange_dtype
    if slot.train_engine is not None:
        slot.apply_grad_sync(slot.train_engine)
    _register_for_shutdown(slot)
    if hasattr(ddp_model, "require_backward_grad_sync"):
        ddp_model.require_backward_grad_sync = False      # DDP's reducer never arms itself again

        @contextlib.contextmanager
        def no_sync():
            eng = slot.train_engine
            if eng is None:
                raise RuntimeError("cpt_b200: no_sync() before the first training forward")
            old = eng.grad_sync_skip
            eng.grad_sync_skip = True
            try:
                yield
            finally:
                eng.grad_sync_skip = old
        ddp_model.no_sync = no_sync
    return ddp_model
