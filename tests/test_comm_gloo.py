"""CPU, world_size 2 over gloo: the N>1 host logic of the path — row sharding that keeps a query's fan-out rows on
one rank, the single logits all-gather (ragged shards), and the reference-style de-duplicating merge."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpt_b200 import comm
from oracle import cpt_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fanouts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        full = torch.randn(sum(fanouts), 3, generator=g)           # what one process would have computed
        q0, q1, r0, r1 = comm.shard_queries(fanouts)
        local = full[r0:r1].clone()                                # this rank's rows
        gathered = comm.all_gather_logits(local)
        ok = torch.equal(gathered, full)
        # fast path: every rank derives all ranks' row counts from the fan-outs -> no size exchange, no host sync
        sizes = comm.shard_rows(fanouts)
        ok = ok and sizes[rank] == r1 - r0 and sum(sizes) == sum(fanouts)
        ok = ok and torch.equal(comm.all_gather_logits(local, sizes=sizes), full)
        # the per-query decision itself is a CUDA kernel (tests/test_gpu_scoring.py); here the oracle stands in for it
        picks = torch.tensor([O.refcoco_zsl_pick(gathered[sum(fanouts[:i]):sum(fanouts[:i + 1])])
                              for i in range(len(fanouts))])
        # per-rank dicts with an overlapping (duplicated) key, as DistributedSampler padding produces
        mine = {int(i): int(picks[i]) for i in range(q0, q1)}
        mine[0] = int(picks[0])
        objs = [None] * world
        dist.all_gather_object(objs, mine)
        merged = comm.merge_by_key(objs)
        ok = ok and merged == {i: int(picks[i]) for i in range(len(fanouts))}
        comm.synchronize()
        q.put((rank, bool(ok), (q0, q1, r0, r1)))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_gather_merge():
    fanouts = [3, 1, 4, 2, 2, 5, 1]   # ragged: rank row counts differ
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fanouts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    (a0, a1, ar0, ar1), (b0, b1, br0, br1) = res[0][2], res[1][2]
    assert a0 == 0 and a1 == b0 and b1 == len(fanouts) and ar1 == br0 and br1 == sum(fanouts)


def test_shard_queries_never_splits_a_query_and_balances_rows():
    fan = [4] * 32                                       # VCR: 4 answer rows per question, 8 GPUs
    spans = [comm.shard_queries(fan, r, 8) for r in range(8)]
    assert all(r1 - r0 == 16 for _, _, r0, r1 in spans)
    fan = [1, 7, 2, 9, 3, 3, 8, 1, 1, 6]
    spans = [comm.shard_queries(fan, r, 4) for r in range(4)]
    assert spans[0][0] == 0 and spans[-1][1] == len(fan)
    for (q0, q1, r0, r1), nxt in zip(spans, spans[1:]):
        assert q1 == nxt[0] and r1 == nxt[2] and r1 - r0 == sum(fan[q0:q1])


def test_single_process_helpers_and_no_cpu_pick():
    g = torch.Generator().manual_seed(7)
    lg = torch.rand(6, 4, generator=g) + 0.5
    assert comm.get_world_size() == 1 and comm.is_main_process()
    assert comm.all_gather_logits(lg) is lg
    assert comm.shard_rows([2, 1, 3], world=2) == [3, 3]
    try:  # the per-query decision is a CUDA kernel: there is no CPU path to fall back to
        comm.pick_per_query(lg, [2, 1, 3], "zsl")
        raised = False
    except RuntimeError:
        raised = True
    assert raised
