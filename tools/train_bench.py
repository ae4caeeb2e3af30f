"""Time the native training step (forward with tape + backward + AdamW) of REC_MLM_CPT on one B200 and print the
per-kernel-class split.  Oscar-base geometry, synthetic weights/data.
    python tools/train_bench.py [--batch 16] [--T 70] [--R 50] [--steps 10] [--dropout 0.1]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--T", type=int, default=70)
    ap.add_argument("--R", type=int, default=50)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--ddp", default="", choices=["", "ddp", "overlap"],
                    help="under torchrun: wrap in DistributedDataParallel; 'overlap' = comm.enable_overlapped_grad_sync")
    a = ap.parse_args()
    rank = 0
    if a.ddp:
        import torch.distributed as dist
        rank = int(os.environ["RANK"])
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl")
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    cfg = C.oscar_base(num_hidden_layers=a.layers)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = a.dropout
    cfg.cpt_b200_train_dtype = a.dtype
    sd = synth_state_dict(cfg, seed=1)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(pre.cuda())
    rec.train()
    net = rec
    if a.ddp:
        net = torch.nn.parallel.DistributedDataParallel(rec, device_ids=[torch.cuda.current_device()],
                                                        find_unused_parameters=True)
        if a.ddp == "overlap":
            from cpt_b200 import comm
            comm.enable_overlapped_grad_sync(net)
    B, T, R = a.batch, a.T, a.R
    b = synth_batch(cfg, B, T, R, seed=2)
    d = {k: v.cuda() for k, v in b.items()}
    labels = torch.full((B, T + R), -1, dtype=torch.long)
    labels[torch.arange(B), b["mask_pos"]] = torch.arange(B) % 7 + 1000
    labels = labels.cuda()
    params = [p for p in rec.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-5, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss, _ = net(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                      masked_lm_labels=labels)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if a.ddp:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"workload": "REC_MLM_CPT train step under DDP (%s), Oscar-base B=%d/GPU S=%d dropout=%g, "
                                          "%d GPUs" % (a.ddp, a.batch, a.T + a.R, a.dropout, dist.get_world_size()),
                              "ms_per_step_max_over_ranks": round(t.item(), 3),
                              "samples_per_s": round(a.batch * dist.get_world_size() / t.item() * 1e3, 1)}))
        dist.destroy_process_group()
        return
    # fwd / bwd / optimizer split with events
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    opt.zero_grad(set_to_none=True)
    ev[0].record()
    loss, _ = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                  masked_lm_labels=labels)
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    eng = rec.bert.train_engine()[0]
    eng.profile(True)
    opt.zero_grad(set_to_none=True)
    loss, _ = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                  masked_lm_labels=labels)
    fwd = eng.profile_read()
    loss.backward()
    bwd = eng.profile_read()
    eng.profile(False)
    out = {"workload": "REC_MLM_CPT train step, Oscar-base L=%d B=%d S=%d dropout=%g %s" % (a.layers, B, T + R, a.dropout, a.dtype),
           "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1),
           "split_ms": {"forward(+weight refresh)": round(ev[0].elapsed_time(ev[1]), 3),
                        "backward": round(ev[1].elapsed_time(ev[2]), 3), "adamw": round(ev[2].elapsed_time(ev[3]), 3)},
           "loss": round(loss.item(), 4),
           "forward_kernels_ms": {k: [round(v[0], 3), v[1]] for k, v in fwd.items()},
           "backward_kernels_ms": {k: [round(v[0], 3), v[1]] for k, v in bwd.items()}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
