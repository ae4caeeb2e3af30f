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

import functools
print = functools.partial(print, flush=True)  # noqa: A001  (a hung worker must still show how far it got)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from cpt_b200 import config as C  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)   # a hang shows where, and ends the process
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
    mode = sys.argv[1] if len(sys.argv) > 1 else "ddp"
    overlap = mode.startswith("overlap")
    if overlap:  # gradients averaged inside the native backward, group by group; DDP's own reducer switched off
        from cpt_b200 import comm
        comm.enable_overlapped_grad_sync(ddp, exchange_dtype="fp32" if mode == "overlap32" else "auto")
    sl = slice(rank * per, (rank + 1) * per)
    d = {k: v[sl].cuda() for k, v in b.items()}
    f = {k: v.cuda() for k, v in b.items()}
    wl, _ = whole(f["input_ids"], f["token_type_ids"], f["attention_mask"], img_feats=f["img_feats"],
                  masked_lm_labels=labels.cuda())
    wl.backward()
    ref = dict(whole.named_parameters())

    def worst_difference(scale=1.0):
        worst, same_none = 0.0, True
        for k, p in rec.named_parameters():
            r = ref[k].grad
            if p.grad is None:
                same_none = same_none and r is None
                continue
            if k.endswith("attention.self.key.bias"):
                continue
            e = (p.grad - scale * r).abs().max().item() / max(scale * r.abs().max().item(), 1e-30)
            worst = max(worst, e)
        return worst, same_none

    ok = True
    # pass 0 runs eagerly, pass 1 captures the forward, pass 2 the backward (with NCCL inside when overlapped), pass 3
    # replays both graphs: every one of them must leave the whole-batch gradient behind
    for it in range(4 if overlap else 1):
        rec.zero_grad()
        loss, _ = ddp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                      masked_lm_labels=labels[sl].cuda())
        loss.backward()
        worst, same_none = worst_difference()
        if rank == 0:
            print("pass %d: worst relative gradient difference %s vs whole batch: %.3e"
                  % (it, "overlapped in-backward all-reduce (%s)" % mode if overlap else "DDP-averaged", worst))
        ok = ok and same_none and worst < 3e-2
    if overlap:
        eng = rec.bert.train_engine()[0]
        if rank == 0:
            print("graph replays of the training step: forward %s backward-with-exchange %s; exchange dtype %s"
                  % (any(st["fwd"] is not None for st in eng._tgraphs.values()),
                     any(st.get("bwd_sync") is not None for st in eng._tgraphs.values()), eng.grad_sync_dtype))
        ok = ok and any(st.get("bwd_sync") is not None for st in eng._tgraphs.values())
        # gradient accumulation: two local passes under no_sync(), then one exchanging pass = 3 x the whole-batch gradient
        rec.zero_grad()
        for k in range(3):
            import contextlib
            with (ddp.no_sync() if k < 2 else contextlib.nullcontext()):
                l3, _ = ddp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                            masked_lm_labels=labels[sl].cuda())
                l3.backward()
        worst, _ = worst_difference(3.0)
        if rank == 0:
            print("accumulated over 2 no_sync() passes + 1 exchanging pass: worst difference vs 3 x whole batch %.3e" % worst)
        ok = ok and worst < 3e-2 and eng._unsynced is None
        rec.zero_grad()
        loss, _ = ddp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                      masked_lm_labels=labels[sl].cuda())
        loss.backward()
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
    sys.stdout.flush()
    dist.destroy_process_group()   # comm.enable_overlapped_grad_sync made this release the NCCL-holding graphs first
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
