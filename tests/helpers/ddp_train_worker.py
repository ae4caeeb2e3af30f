"""Worker of tests/test_gpu_ddp.py: one rank of a DistributedDataParallel few-shot training step
(fewshot/refcoco_cpt.py:300-317 wraps the model in DDP(find_unused_parameters=True); :243-248 is the step).

Every rank holds the same weights and a different shard of one global batch.  After backward, DDP has averaged the
gradients over ranks; rank 0 compares them with the gradients of the whole batch computed by a plain (non-DDP) copy of
the model, then checks that an optimizer step leaves all ranks with identical parameters.
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from cpt_b200 import config as C  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    cfg = C.oscar_tiny(num_hidden_layers=2)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.0
    sd = synth_state_dict(cfg, seed=3)
    per = 3
    B, T, R = per * world, 40, 24
    b = synth_batch(cfg, B, T, R, seed=17)
    labels = torch.full((B, T + R), -1, dtype=torch.long)
    labels[torch.arange(B), b["mask_pos"]] = torch.arange(B) % 11 + 5
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(pre.cuda())
    rec.train()
    whole = copy.deepcopy(rec)  # gqa_cpt.py:384 deep-copies the model; the native handle must not be shared
    ddp = torch.nn.parallel.DistributedDataParallel(rec, device_ids=[torch.cuda.current_device()],
                                                    find_unused_parameters=True)
    overlap = len(sys.argv) > 1 and sys.argv[1] == "overlap"
    if overlap:  # gradients averaged inside the native backward, group by group; DDP's own reducer switched off
        from cpt_b200 import comm
        comm.enable_overlapped_grad_sync(ddp)
    sl = slice(rank * per, (rank + 1) * per)
    d = {k: v[sl].cuda() for k, v in b.items()}
    loss, _ = ddp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                  masked_lm_labels=labels[sl].cuda())
    loss.backward()
    ok = True
    if rank == 0:
        f = {k: v.cuda() for k, v in b.items()}
        wl, _ = whole(f["input_ids"], f["token_type_ids"], f["attention_mask"], img_feats=f["img_feats"],
                      masked_lm_labels=labels.cuda())
        wl.backward()
        ref = dict(whole.named_parameters())
        worst = 0.0
        for k, p in rec.named_parameters():
            r = ref[k].grad
            if p.grad is None:
                ok = ok and r is None
                continue
            if k.endswith("attention.self.key.bias"):
                continue
            e = (p.grad - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
            worst = max(worst, e)
        print("worst relative gradient difference %s vs whole batch: %.3e"
              % ("overlapped in-backward all-reduce" if overlap else "DDP-averaged", worst))
        ok = ok and worst < 3e-2
    opt = torch.optim.AdamW([p for p in ddp.parameters() if p.requires_grad], lr=1e-3)
    opt.step()
    # all ranks must hold identical parameters after the step
    flat = torch.cat([p.detach().flatten() for p in rec.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all().item())
    # second step runs on the updated weights (the handle must refresh its 16-bit copies)
    l2, _ = ddp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                masked_lm_labels=labels[sl].cuda())
    t = torch.tensor([loss.item(), l2.item()], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("mean loss before %.5f after %.5f; parameters identical across ranks: %s" % (t[0] / world, t[1] / world, same))
        ok = ok and same and (t[1] < t[0]).item()
        print("DDP_TRAIN_OK" if ok else "DDP_TRAIN_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
