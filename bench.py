#!/usr/bin/env python
"""CPT hot-path benchmark (BASELINE.json metric: CPT samples/sec, Oscar-base, RefCOCO shape).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

A "step" is one pass of the hot path over one batch of synthetic CPT queries: BertImgModel encoder (text +
region embedding -> 12 layers) + the masked-colour-token head gathered at the [MASK] rows and colour ids
(what zeroshot/refcoco_cpt.py:217-219,234-235 consumes).  Workload = BASELINE.json configs[1]: RefCOCO CPT
inference, Oscar-base, batch 64 per GPU, T=70 text tokens + R=50 regions x 2054-d.

Printed JSON (one line, rank 0): `value` = device-timed whole-job samples/s with inputs resident in HBM;
`e2e` = the same through the public module API with HOST (pinned) inputs copied in and logits copied out every
step; `roofline` = the dominant kernel (the FFN GEMM) against the measured bf16/fp16 tensor peak in
MEASURED_PEAKS.json; `cpu_baseline` = the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cpt_b200 import comm  # noqa: E402
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids  # noqa: E402

T_TEXT, R_REG, K_IDS = 70, 50, 2
METRIC = "CPT samples/sec (Oscar-base, RefCOCO CPT inference, 50 regions x 2054-d + 70 text tokens, S=120)"
UNIT = "samples/s"


def flops_per_sample(cfg, T, R, K):
    """Algorithmic forward FLOPs of one row (SURVEY.md 8d / BASELINE.md 3)."""
    S, H, L, F = T + R, cfg.hidden_size, cfg.num_hidden_layers, cfg.img_feature_dim
    return L * (24 * S * H * H + 4 * S * S * H) + 2 * R * F * H + (2 * H * H + 2 * H * K)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        busy = sorted(x for x in sm if x > 0)
        # median over the upper half of the samples (= under load; idle samples before/after are dropped)
        med = busy[len(busy) // 2:][len(busy[len(busy) // 2:]) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step(sd, cfg, b, vids):
    """The reference's own call sequence: full [B,S,V] scores, then the caller's gathers."""
    from oracle import cpt_oracle as O
    with torch.no_grad():
        scores = O.rec_mlm_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                               img_feats=b["img_feats"])[0]
        return scores[torch.arange(scores.size(0)), b["mask_pos"]][:, vids]


def make_cpu_reference(cfg, sd, vids):
    """Returns (step_fn(batch) -> [B,K] logits, kind, description).  Preferred: the reference's OWN modules
    (oscar.modeling.modeling_rec.REC_MLM_CPT from the offline install in baseline/_ref, its un-vendored
    pytorch-transformers dependency supplied by oracle/ref_shim.py) called exactly as
    oscar/zeroshot/refcoco_cpt.py:217-219 does; otherwise the oracle port."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    try:
        if not os.path.isfile(os.path.join(ref_root, "oscar", "modeling", "modeling_rec.py")):
            raise ImportError("baseline/_ref/oscar is not installed")
        from oracle import ref_shim
        ref_shim.install(ref_root)
        from oscar.modeling.modeling_bert import BertImgForPreTraining as RefPre
        from oscar.modeling.modeling_rec import REC_MLM_CPT as RefRec
        d = cfg.to_dict()
        v = d.pop("vocab_size")
        rcfg = ref_shim.BertConfig(v, **d)
        pre = RefPre(rcfg)
        missing, unexpected = pre.load_state_dict(sd, strict=False)
        assert not missing and not unexpected
        pre.tie_weights()
        rec = RefRec(rcfg)
        rec.copy_from_pretraining_model(pre)
        rec.eval()

        def step(b):
            with torch.no_grad():
                out = rec(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[0]
                return out[torch.arange(out.size(0)), b["mask_pos"]][:, vids]
        return step, "reference", ("the reference's own oscar.modeling.modeling_rec.REC_MLM_CPT (unmodified, installed "
                                   "offline into baseline/_ref; pytorch-transformers 1.x blocks from oracle/ref_shim.py), "
                                   "fp32 CPU, called as zeroshot/refcoco_cpt.py:217-219")
    except Exception as ex:  # noqa: BLE001
        why = "%s: %s" % (type(ex).__name__, str(ex)[:80])
        return (lambda b: oracle_step(sd, cfg, b, vids)), "port", \
            "oracle/cpt_oracle.py (torch fp32 CPU restatement; reference modules unavailable: %s)" % why


def cpu_train_leg(cfg, sd, batch=8, iters=2):
    """One few-shot training step (forward with labels + backward, no optimizer) of the reference's own REC_MLM_CPT on
    the host cores, on a bounded sample — the CPU baseline of the `train_step` leg.  Falls back to autograd through the
    oracle port when baseline/_ref is absent."""
    torch.set_num_threads(os.cpu_count() or 1)
    b = synth_batch(cfg, batch, T_TEXT, R_REG, seed=88)
    labels = torch.full((batch, T_TEXT + R_REG), -1, dtype=torch.long)
    labels[torch.arange(batch), b["mask_pos"]] = 2000 + torch.arange(batch) % 7
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    try:
        if not os.path.isfile(os.path.join(ref_root, "oscar", "modeling", "modeling_rec.py")):
            raise ImportError("baseline/_ref/oscar is not installed")
        from oracle import ref_shim
        ref_shim.install(ref_root)
        from oscar.modeling.modeling_bert import BertImgForPreTraining as RefPre
        from oscar.modeling.modeling_rec import REC_MLM_CPT as RefRec
        d = cfg.to_dict()
        v = d.pop("vocab_size")
        rcfg = ref_shim.BertConfig(v, **d)
        pre = RefPre(rcfg)
        pre.load_state_dict(sd, strict=False)
        pre.tie_weights()
        rec = RefRec(rcfg)
        rec.copy_from_pretraining_model(pre)
        rec.train()

        def step():
            rec.zero_grad()
            loss = rec(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"],
                       masked_lm_labels=labels)[0]
            loss.backward()
        kind = "reference"
    except Exception:  # noqa: BLE001
        from oracle import cpt_oracle as O
        leaf = {k: t.clone().requires_grad_(True) for k, t in sd.items()}

        def step():
            for t in leaf.values():
                t.grad = None
            O.rec_mlm_cpt(leaf, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"], masked_lm_labels=labels,
                          img_feats=b["img_feats"], training=True)[0].backward()
        kind = "port"
    step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    per = (time.perf_counter() - t0) / iters
    return {"value": batch / per, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%d timed forward+backward passes (1 warm-up) of a %d-row batch, %.2f s per pass, fp32 CPU, no "
                      "optimizer step" % (iters, batch, per)}


def cpu_leg(cfg, sd, vids, batch, iters, warmup=1):
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, what = make_cpu_reference(cfg, sd, vids)
    b = synth_batch(cfg, batch, T_TEXT, R_REG, seed=88)
    for _ in range(warmup):
        step(b)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        step(b)
        ts.append(time.perf_counter() - t0)
    return batch * len(ts) / sum(ts), sum(ts) / len(ts), kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = C.oscar_base()
    sd = synth_state_dict(cfg, seed=88)
    vids = synth_vocab_ids(cfg, K_IDS, seed=88)
    bs = args.ref_batch
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, what = make_cpu_reference(cfg, sd, vids)
    b = synth_batch(cfg, bs, T_TEXT, R_REG, seed=88)
    for _ in range(args.warmup):
        step(b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(b)
    dt = time.perf_counter() - t0
    val = bs * args.steps / dt
    cores = torch.get_num_threads()
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "RefCOCO CPT inference (BASELINE.json configs[1] shape): Oscar-base, T=70 R=50 F=2054, "
                                  "K=2 colour ids; each step = a bounded sample of %d rows (full-vocab head over all S, "
                                  "then gather, as the reference runs it)" % bs, "batch_per_step": bs},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": "%d steps x %d rows; %s" % (args.steps, bs, what)},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def train_leg(cfg, model, batch, dev, steps=10, warmup=3):
    """REC_MLM_CPT training step (native forward with tape + backward + torch's fused AdamW) on one resident batch of
    the bench shape, dropout 0.1 as the reference's few-shot runs (fewshot/refcoco_cpt.py:243-248)."""
    orig = (cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.1
    B, S = batch["input_ids"].shape[0], batch["attention_mask"].shape[1]
    labels = torch.full((B, S), -1, dtype=torch.long, device=dev)
    labels[torch.arange(B, device=dev), batch["mask_pos"]] = 2000 + torch.arange(B, device=dev) % 7
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-6, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model(batch["input_ids"], batch["token_type_ids"], batch["attention_mask"],
                     img_feats=batch["img_feats"], masked_lm_labels=labels)[0]
        loss.backward()
        opt.step()
        return loss

    model.train()
    try:
        with torch.enable_grad():
            return _train_leg_timed(model, step, steps, warmup, B, S)
    finally:
        model.eval()
        opt.zero_grad(set_to_none=True)
        cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob = orig


def _train_leg_timed(model, step, steps, warmup, B, S):
    for _ in range(warmup):
        step()
    eng = model.bert.train_engine()[0]
    l0 = eng.launch_count()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": "REC_MLM_CPT few-shot training step: forward with tape + backward + fused AdamW, Oscar-base, "
                        "batch %d, S=%d, dropout 0.1, bf16 GEMM operands / fp32 master weights" % (B, S),
            "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "steps": steps, "warmup": warmup,
            "gpu_launches_per_step": (eng.launch_count() - l0) / steps, "loss": float(loss.detach())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="rows per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=16, help="rows per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="run warmup+steps once without the extra legs (ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    real_stdout = None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        # ... and whatever a library still writes to fd 1 (NCCL's version banner) goes to stderr: the JSON line is
        # written to the saved descriptor at the end
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group(backend="nccl", device_id=dev)

    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT

    cfg = C.oscar_base()
    sd = synth_state_dict(cfg, seed=88)
    vids_cpu = synth_vocab_ids(cfg, K_IDS, seed=88)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    pre = pre.to(dev).eval()
    model = REC_MLM_CPT(cfg)
    model.copy_from_pretraining_model(pre)
    model.eval()
    vids = vids_cpu.to(dev)
    B = args.batch
    NROT = 8  # distinct input batches: 8 x 26 MB of region features > 126 MB L2, so inputs are never L2-hot
    host = [synth_batch(cfg, B, T_TEXT, R_REG, seed=1000 + rank * 100 + i, dense=(i % 2 == 1)) for i in range(NROT)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def step_resident(i):
        b = devb[i % NROT]
        return model(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"],
                     mask_pos=b["mask_pos"], vocab_ids=vids)[0]


    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        step_resident(0)  # builds the engine, converts weights
        eng = model.bert.engine()
        model.bert.freeze_engine_weights(True)
        # every rotating batch is seen twice so that its CUDA graph is captured before the timed region
        for i in range(args.warmup if args.profile_only else max(args.warmup, 2 * NROT)):
            out = step_resident(i)
            if world > 1:
                comm.all_gather_logits(out, sizes=[B] * world)
        if args.profile_only:
            for i in range(args.steps):
                step_resident(i)
            torch.cuda.synchronize()
            return

        # ---------------- device-timed leg: inputs resident in HBM
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        l0 = eng.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            out = step_resident(i)
            if world > 1:
                comm.all_gather_logits(out, sizes=[B] * world)  # the path's only collective: the final [B,K] logits
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())

        # ---------------- end-to-end leg: host (pinned) inputs in, logits out, every step
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()
        slots = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
        slot_ready = [torch.cuda.Event() for _ in range(2)]
        slot_free = [torch.cuda.Event() for _ in range(2)]
        out_host = [torch.empty(B, K_IDS).pin_memory() for _ in range(2)]
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        d2h = B * K_IDS * 4

        def upload(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(slot_free[s])
                for k, v in host[i % NROT].items():
                    slots[s][k].copy_(v, non_blocking=True)
                slot_ready[s].record(copy_stream)

        def e2e_loop(n):
            for s in range(2):
                slot_free[s].record(main_stream)
            upload(0)
            for i in range(n):
                s = i % 2
                if i + 1 < n:
                    upload(i + 1)  # overlaps with this step's compute
                main_stream.wait_event(slot_ready[s])
                b = slots[s]
                o = model(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"],
                          mask_pos=b["mask_pos"], vocab_ids=vids)[0]
                slot_free[s].record(main_stream)
                if world > 1:
                    comm.all_gather_logits(o, sizes=[B] * world)
                out_host[s].copy_(o, non_blocking=True)
            torch.cuda.synchronize()

        e2e_loop(max(6, args.warmup))  # both upload slots get their CUDA graph captured
        barrier()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())

        # ---------------- roofline leg: per-kernel-class CUDA-event timing of the same steps (rank 0)
        roof, kernels = None, None
        if rank == 0:
            eng.profile(True)  # eager launches bracketed by CUDA events (graph replay is bypassed while profiling)
            for i in range(args.steps):
                step_resident(i)
            prof = eng.profile_read()
            eng.profile(False)
            M, H, I = B * (T_TEXT + R_REG), cfg.hidden_size, cfg.intermediate_size
            gflop = {"gemm_qkv": 2.0 * M * 3 * H * H, "gemm_attn_out": 2.0 * M * H * H, "gemm_ffn_up": 2.0 * M * I * H,
                     "gemm_ffn_down": 2.0 * M * H * I, "gemm_img": 2.0 * B * R_REG * cfg.img_feature_dim * H,
                     "attention": 4.0 * B * (T_TEXT + R_REG) ** 2 * H}
            kernels = {}
            for name, (kms, n) in prof.items():
                d = {"ms_per_step": kms / args.steps, "launches_per_step": n / args.steps, "us_per_launch": 1e3 * kms / n}
                if name in gflop:
                    d["tflops"] = gflop[name] * n / (kms * 1e-3) / 1e12
                kernels[name] = d
            dom = max((k for k in kernels if k in ("gemm_ffn_up", "gemm_ffn_down", "gemm_qkv", "gemm_attn_out")),
                      key=lambda k: kernels[k]["ms_per_step"])
            burst, sustained, hbm, how = peaks()
            ach = kernels[dom]["tflops"]
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(dom)
            roof = {"kernel": "gemm_kernel (%s)" % dom, "bound": "tensor", "achieved": ach, "peak": sustained,
                    "unit": "TFLOP/s", "frac": ach / sustained, "frac_of_burst_peak": ach / burst,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s); kernel timed inside a long step"
                                   % how,
                    "flops_per_launch": gflop[dom], "us_per_launch": kernels[dom]["us_per_launch"],
                    "traffic": traffic}

        # the few-shot training step: reported next to the headline, not part of it
        train = None
        if rank == 0 and world == 1:
            try:
                train = train_leg(cfg, model, devb[0], dev)
            except Exception as e:  # never let the extra leg take the headline line down
                train = {"error": str(e)[:200]}
            if "error" not in train and not args.no_cpu_baseline:
                try:
                    with torch.enable_grad():
                        train["cpu_baseline"] = cpu_train_leg(cfg, sd)
                except Exception as e:
                    train["cpu_baseline"] = {"error": str(e)[:200]}

        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            v, per, kind, what = cpu_leg(cfg, sd, vids_cpu, 32, 4)
            cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                   "sample": "4 timed passes (1 warm-up) of a 32-row batch of the same workload (full [B,S,V] head "
                             "then gather), %.2f s per pass; %s" % (per, what)}

    if rank == 0:
        total = B * world * args.steps
        value = total / (ms * 1e-3)
        fl = flops_per_sample(cfg, T_TEXT, R_REG, K_IDS)
        burst, sustained, hbm, how = peaks()
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f16", "data": "synthetic",
               "config": {"workload": "RefCOCO CPT inference (BASELINE.json configs[1]): Oscar-base, batch %d per GPU, "
                                      "T=70 text tokens + R=50 regions x 2054-d, K=2 colour ids, encoder + gathered "
                                      "masked-colour-token head" % B,
                          "batch_per_gpu": B, "seq_len": T_TEXT + R_REG,
                          "l2": "inputs rotate over %d distinct batches (%d MB) and one step streams ~170 MB of "
                                "16-bit weights + ~200 MB of activations: larger than the 126 MB L2" %
                                (NROT, NROT * h2d // 2 ** 20),
                          "parallelism": "dp%d (rows sharded, NCCL all-gather of the [B,K] logits)" % world},
               "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": launches,
               "algorithmic_gflop_per_sample": fl / 1e9,
               "model_tflops": value * fl / 1e12,
               "model_frac_of_sustained_peak": value * fl / 1e12 / (sustained * world),
               "clocks": clocks, "roofline": roof, "kernels": kernels}
        if cpu:
            out["cpu_baseline"] = cpu
        if train:
            out["train_step"] = train
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(out) + "\n").encode())
        else:
            print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
