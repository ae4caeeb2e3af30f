"""Worker of tests/test_gpu_ddp.py::test_peer_memory_logits_exchange: one rank of the fused head + peer-memory
all-gather (cpt_b200.comm.LogitsExchange) against NCCL's all_gather of the same logits, over several steps (eager run,
graph capture, graph replays, fresh input tensors)."""
import functools
import os
import sys

import torch
import torch.distributed as dist

print = functools.partial(print, flush=True)  # noqa: A001
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from cpt_b200 import comm, config as C  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    cfg = C.oscar_tiny(num_hidden_layers=2)
    sd = synth_state_dict(cfg, seed=3)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(pre.cuda())
    rec.eval()
    ok = True
    for B, K in ((6, 5), (16, 40)):
        vids = synth_vocab_ids(cfg, K, seed=5).cuda()
        ex = comm.LogitsExchange(B, K)
        worst = 0.0
        for step in range(7):
            b = synth_batch(cfg, B, 40, 24, seed=100 * step + rank)
            d = {k: v.cuda() for k, v in b.items()}   # fresh tensors every step, as the reference's loop makes them
            with torch.no_grad():
                local = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                            mask_pos=d["mask_pos"], vocab_ids=vids)[0]
                fused = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                            mask_pos=d["mask_pos"], vocab_ids=vids, gather=ex)[0]
            want = comm.all_gather_logits(local, [B] * world)
            again = ex.rows(rec.bert.engine(), local)
            rec.bert.engine().check()
            ok = ok and tuple(fused.shape) == (world * B, K)
            worst = max(worst, (fused - want).abs().max().item(), (again - want).abs().max().item())
        eng = rec.bert.engine()
        if rank == 0:
            print("B=%d K=%d: 7 steps, fused exchange vs NCCL all_gather max |diff| %.3e; graph replays so far %d"
                  % (B, K, worst, eng.graph_replays))
        ok = ok and worst == 0.0 and eng.graph_replays > 0
        ex.close()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("EXCHANGE_OK" if t.item() == 1.0 else "EXCHANGE_FAIL")
    dist.destroy_process_group()
    sys.exit(0)


if __name__ == "__main__":
    main()
