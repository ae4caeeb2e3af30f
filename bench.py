#!/usr/bin/env python
"""CPT hot-path benchmark (BASELINE.json metric: CPT samples/sec on the Oscar cross-modal BERT path).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, BASELINE.json configs[1]
    python bench.py --config {2,3,4,5} ...                   # the other BASELINE.json configs (numbered from 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's own modules on the host CPU cores

Workloads (BASELINE.json `configs`; synthetic data, random-init weights — no datasets / checkpoints offline):
  2 (default)  RefCOCO CPT inference, Oscar-base, batch 64 per GPU, T=70 text tokens + R=50 regions x 2054-d, K=2 colour
               ids; step = encoder + gathered masked-colour-token head on one batch; weak scaling, one logits all-gather.
  3            GQA CPT few-shot fine-tune step, Oscar-base, T=165 + R=45 (S=210), micro-batch 4 per GPU, dropout
               0.3 / 0.1, DDP gradient all-reduce inside the native backward, fused global-norm clip + native AdamW;
               step = forward + backward + all-reduce + optimizer on one micro-batch.
  4            VCR q->a inference, Oscar-base, S=210, 32 questions x 4 answer rows = 128 rows per step over ALL GPUs
               (strong scaling), NSP head + per-question argmax on the device.
  5            RefCOCO CPT inference, Oscar-large (24 layers, H=1024), batch 256 per GPU, T=150 + R=50 (S=200).

Printed JSON (one line, rank 0): `value` = device-timed whole-job samples/s with inputs resident in HBM; `e2e` = the
same through the public module API with HOST (pinned) inputs copied in and results copied out every step; `roofline` =
the dominant kernel against the measured tensor peak in MEASURED_PEAKS.json; `parity` = the timed batch's outputs
against the fp32 oracle on a row sample; `cpu_baseline` = the reference's own modules on this box's host cores on a
bounded sample.
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

UNIT = "samples/s"
WORKLOADS = {
    2: dict(name="RefCOCO CPT inference (BASELINE.json configs[1])", model="base", T=70, R=50, K=2, batch=64, head="mlm",
            kind="infer", scaling="weak", metric="CPT samples/sec (Oscar-base, RefCOCO CPT inference, 50 regions x "
                                                 "2054-d + 70 text tokens, S=120)"),
    3: dict(name="GQA CPT few-shot fine-tune step (BASELINE.json configs[2])", model="base", T=165, R=45, K=1853, batch=4,
            head="mlm", kind="train", scaling="weak",
            metric="CPT training samples/sec (Oscar-base, GQA few-shot step, S=210, micro-batch 4 per GPU)"),
    4: dict(name="VCR q->a CPT inference (BASELINE.json configs[3])", model="base", T=165, R=45, K=0, batch=128, head="nsp",
            kind="infer", scaling="strong",
            metric="CPT answer rows/sec (Oscar-base, VCR q->a, S=210, 32 questions x 4 answers per step over all GPUs)"),
    5: dict(name="RefCOCO CPT inference, Oscar-large (BASELINE.json configs[4])", model="large", T=150, R=50, K=2,
            batch=256, head="mlm", kind="infer", scaling="weak",
            metric="CPT samples/sec (Oscar-large 24L/1024H, RefCOCO CPT inference, S=200)"),
}


def model_cfg(w):
    return C.oscar_large() if w["model"] == "large" else C.oscar_base()


def flops_per_sample(cfg, T, R, K):
    """Algorithmic forward FLOPs of one row (SURVEY.md 8d / BASELINE.md 3)."""
    S, H, L, F = T + R, cfg.hidden_size, cfg.num_hidden_layers, cfg.img_feature_dim
    return L * (24 * S * H * H + 4 * S * S * H) + 2 * R * F * H + (2 * H * H + 2 * H * max(K, 1))


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


# ---------------------------------------------------------------------------------------------------- CPU reference legs
def _ref_modules(cfg, sd):
    """The reference's OWN modules (offline install in baseline/_ref; the un-vendored pytorch-transformers 1.x blocks
    come from oracle/ref_shim.py), or None when baseline/_ref is absent."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_root, "oscar", "modeling", "modeling_rec.py")):
        return None
    from oracle import ref_shim
    ref_shim.install(ref_root)
    from oscar.modeling.modeling_bert import BertImgForPreTraining as RefPre
    from oscar.modeling.modeling_rec import REC_MLM_CPT as RefRec
    from oscar.modeling.modeling_vcr import NSPCPT as RefNsp
    d = cfg.to_dict()
    v = d.pop("vocab_size")
    rcfg = ref_shim.BertConfig(v, **d)
    pre = RefPre(rcfg)
    missing, unexpected = pre.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    pre.tie_weights()
    rec, nsp = RefRec(rcfg), RefNsp(rcfg)
    rec.copy_from_pretraining_model(pre)
    nsp.copy_from_pretraining_model(pre)
    return rec, nsp


def make_cpu_reference(cfg, sd, vids, head):
    """(step_fn(batch) -> outputs, kind, description): the reference called exactly as its scripts call it — full
    [B,S,V] scores then the caller's gather (Oscar/oscar/zeroshot/refcoco_cpt.py:217-219) for the MLM workloads,
    model(**inputs)[0] (Oscar/oscar/fewshot/vcr_nsp_cpt.py:597) for VCR; the oracle port when the modules are absent."""
    try:
        mods = _ref_modules(cfg, sd)
        if mods is None:
            raise ImportError("baseline/_ref/oscar is not installed")
        rec, nsp = mods
        rec.eval()
        nsp.eval()

        def step(b):
            with torch.no_grad():
                if head == "nsp":
                    return nsp(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[0]
                out = rec(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[0]
                return out[torch.arange(out.size(0)), b["mask_pos"]][:, vids]
        return step, "reference", ("the reference's own oscar.modeling modules (unmodified, installed offline into "
                                   "baseline/_ref; pytorch-transformers 1.x blocks from oracle/ref_shim.py), fp32 CPU")
    except Exception as ex:  # noqa: BLE001
        why = "%s: %s" % (type(ex).__name__, str(ex)[:80])
        from oracle import cpt_oracle as O

        def step(b):
            with torch.no_grad():
                if head == "nsp":
                    return O.nsp_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                     img_feats=b["img_feats"])[0]
                s = O.rec_mlm_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                  img_feats=b["img_feats"])[0]
                return s[torch.arange(s.size(0)), b["mask_pos"]][:, vids]
        return step, "port", "oracle/cpt_oracle.py (torch fp32 CPU restatement; reference modules unavailable: %s)" % why


def make_cpu_train_step(cfg, sd, b, labels):
    """One few-shot training step (forward with labels + backward; no optimizer) on the host cores."""
    try:
        mods = _ref_modules(cfg, sd)
        if mods is None:
            raise ImportError
        rec = mods[0]
        rec.train()

        def step():
            rec.zero_grad()
            rec(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"],
                masked_lm_labels=labels)[0].backward()
        return step, "reference"
    except Exception:  # noqa: BLE001
        from oracle import cpt_oracle as O
        leaf = {k: t.clone().requires_grad_(True) for k, t in sd.items()}

        def step():
            for t in leaf.values():
                t.grad = None
            O.rec_mlm_cpt(leaf, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"], masked_lm_labels=labels,
                          img_feats=b["img_feats"], training=True)[0].backward()
        return step, "port"


def train_labels(b, vids, S):
    B = b["input_ids"].shape[0]
    labels = torch.full((B, S), -1, dtype=torch.long)
    labels[torch.arange(B), b["mask_pos"]] = vids[torch.arange(B) % vids.numel()]
    return labels


def cpu_leg(w, cfg, sd, vids, batch, iters, warmup=1):
    """Bounded sample of the workload on the host cores: `iters` timed passes of `batch` rows."""
    torch.set_num_threads(os.cpu_count() or 1)
    b = synth_batch(cfg, batch, w["T"], w["R"], seed=88)
    if w["kind"] == "train":
        labels = train_labels(b, vids, w["T"] + w["R"])
        with torch.enable_grad():
            step, kind = make_cpu_train_step(cfg, sd, b, labels)
            for _ in range(warmup):
                step()
            ts = []
            for _ in range(iters):
                t0 = time.perf_counter()
                step()
                ts.append(time.perf_counter() - t0)
        what = "forward + backward of the reference's REC_MLM_CPT (no optimizer step), fp32 CPU"
    else:
        step, kind, what = make_cpu_reference(cfg, sd, vids, w["head"])
        for _ in range(warmup):
            step(b)
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            step(b)
            ts.append(time.perf_counter() - t0)
    return batch * len(ts) / sum(ts), sum(ts) / len(ts), kind, what


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_cfg(w)
    sd = synth_state_dict(cfg, seed=88)
    vids = synth_vocab_ids(cfg, max(w["K"], 1), seed=88)
    bs = args.ref_batch if args.ref_batch > 0 else {2: 16, 3: 4, 4: 8, 5: 4}[args.config]
    v, per, kind, what = cpu_leg(w, cfg, sd, vids, bs, args.steps, warmup=max(1, min(args.warmup, 2)))
    cores = torch.get_num_threads()
    out = {"impl": "reference", "metric": w["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per, "higher_is_better": True,
           "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s: T=%d R=%d F=2054; each step = a bounded sample of %d rows on the host cores, "
                                  "called as the reference's scripts call it" % (w["name"], w["T"], w["R"], bs),
                      "batch_per_step": bs},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": "%d steps x %d rows; %s" % (args.steps, bs, what)},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------- GPU legs
def build_models(cfg, sd, dev, dtype=None):
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.modeling_vcr import NSPCPT
    if dtype:
        cfg.cpt_b200_dtype = dtype
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    pre = pre.to(dev).eval()
    rec, nsp = REC_MLM_CPT(cfg), NSPCPT(cfg)
    rec.copy_from_pretraining_model(pre)
    nsp.copy_from_pretraining_model(pre)
    return rec.eval(), nsp.eval()


def parity_block(w, cfg, sd, vids_cpu, host_batch, got, rows=4):
    """Max error of the first `rows` rows of a timed batch against the fp32 oracle (test infrastructure, used here as
    the checker only): colour logits relative to the row's largest logit (the tolerance definition of the parity tests),
    NSP scores absolute."""
    from oracle import cpt_oracle as O
    b = {k: v[:rows] for k, v in host_batch.items()}
    with torch.no_grad():
        if w["head"] == "nsp":
            ref = O.nsp_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[0]
            err = (got[:rows].float().cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
            what = "max |nsp score - oracle| / max(1, max|oracle|)"
        else:
            seq = O.bert_img_model(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                   img_feats=b["img_feats"])[0]
            full = O.lm_head(sd, cfg, seq[torch.arange(rows), b["mask_pos"]])
            err = ((got[:rows].float().cpu() - full[:, vids_cpu]).abs() / full.abs().max(dim=1, keepdim=True).values).max().item()
            what = "max |logit - oracle| / max|oracle row| over the gathered colour ids"
    return {"max_rel_err": err, "rows_checked": rows, "definition": what, "tolerance": 1e-3, "ok": bool(err <= 1e-3)}


def infer_leg(args, w, cfg, sd, dev, rank, world, dist, dtype, want_extras):
    """Device-timed + end-to-end legs of an inference workload for one operand dtype."""
    T, R, K = w["T"], w["R"], w["K"]
    B = args.batch if args.batch > 0 else w["batch"]
    if w["scaling"] == "strong":
        if B % world:
            raise SystemExit("--config %d shards %d rows over the GPUs: --gpus must divide it" % (args.config, B))
        B //= world
    vids_cpu = synth_vocab_ids(cfg, max(K, 1), seed=88)
    rec, nsp = build_models(cfg, sd, dev, dtype)
    model = nsp if w["head"] == "nsp" else rec
    vids = vids_cpu.to(dev)
    NROT = 8 if cfg.hidden_size <= 768 else 4  # distinct input batches, larger than the 126 MB L2 together
    host = [synth_batch(cfg, B, T, R, seed=1000 + rank * 100 + i, dense=(i % 2 == 1)) for i in range(NROT)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    fan = [4] * (B // 4) if w["head"] == "nsp" else None

    # the path's only collective, the final logits: by default fused into the head kernel (peer-memory stores over
    # NVLink + one flag per peer, comm.LogitsExchange); --collective nccl = one fixed-shape NCCL all-gather instead
    ex = None
    if world > 1 and args.collective == "peer":
        try:   # raises on every rank together when CUDA IPC peer mapping is not possible on this node
            ex = comm.LogitsExchange(B, int(cfg.num_contrast_classes) if w["head"] == "nsp" else max(K, 1))
        except RuntimeError as e:
            if rank == 0:
                print("bench: %s — falling back to the NCCL all-gather" % e, file=sys.stderr)
            args.collective = "nccl"

    def call(b, gather=None):
        if w["head"] == "nsp":
            return model(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[0]
        return model(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"],
                     mask_pos=b["mask_pos"], vocab_ids=vids, gather=gather)[0]

    def gathered(b):
        if world == 1:
            return call(b)
        if ex is None:
            return comm.all_gather_logits(call(b), sizes=[B] * world)
        if w["head"] == "nsp":
            return ex.rows(model.bert.engine(), call(b))
        return call(b, gather=ex)

    def step_resident(i):
        out = gathered(devb[i % NROT])
        if fan is not None:
            comm.pick_per_query(out, fan * world, "vcr")           # per-question argmax on the device
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = {"batch_per_gpu": B}
    with torch.no_grad():
        call(devb[0])
        eng = rec.bert.engine()
        for i in range(args.warmup if args.profile_only else max(args.warmup, 2 * NROT)):
            step_resident(i)  # every rotating batch is seen twice: its CUDA graph exists before the timed region
        if args.profile_only:
            for i in range(args.steps):
                step_resident(i)
            torch.cuda.synchronize()
            if ex is not None:
                ex.close()
            return None

        sampler = ClockSampler(dev.index) if rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        l0 = eng.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        th0 = time.perf_counter()
        for i in range(args.steps):
            out = step_resident(i)
        host_s = time.perf_counter() - th0
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        res["launches"] = eng.launch_count() - l0
        res["clocks"] = sampler.stop() if sampler else None
        res["host_us_per_step"] = 1e6 * host_s / args.steps  # host time to ENQUEUE a step (graph replay + collective)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res["ms"] = ms
        res["last_out"] = (out[rank * B:(rank + 1) * B] if world > 1 else out).clone()
        res["last_batch"] = host[(args.steps - 1) % NROT]

        # ---------------- end-to-end leg: host (pinned) inputs in, results out, every step
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()
        slots = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
        slot_ready = [torch.cuda.Event() for _ in range(2)]
        slot_free = [torch.cuda.Event() for _ in range(2)]
        out_host = [torch.empty(tuple(out[:B].shape), dtype=out.dtype).pin_memory() for _ in range(2)]
        res["h2d"] = sum(v.numel() * v.element_size() for v in host[0].values())
        res["d2h"] = out_host[0].numel() * out_host[0].element_size()

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
                o = gathered(slots[s])
                slot_free[s].record(main_stream)
                if world > 1:
                    o = o[rank * B:(rank + 1) * B]
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
        res["e2e_s"] = e2e_s
        if ex is not None:
            ex.close()

        if want_extras and rank == 0:
            # the reference's loop as it is written: fresh device tensors every step (zeroshot/refcoco_cpt.py:212-219)
            n_fresh = min(args.steps, 20)
            for i in range(8):   # warm-up in the loop's own steady state: shape-keyed graph captured, allocator blocks recycled
                call({k: v.to(dev, non_blocking=True) for k, v in host[i % NROT].items()})
            torch.cuda.synchronize()
            r0 = eng.graph_replays
            t0 = time.perf_counter()
            for i in range(n_fresh):
                call({k: v.to(dev, non_blocking=True) for k, v in host[i % NROT].items()})
            torch.cuda.synchronize()
            res["fresh_tensor_loop"] = {"samples_per_s": B * n_fresh / (time.perf_counter() - t0),
                                        "graph_replays": eng.graph_replays - r0, "steps": n_fresh,
                                        "what": "every step moves its batch to NEW device tensors, as the reference's loop "
                                                "does; the engine copies them into its staging buffers and replays a graph"}
            if w["head"] == "mlm" and cfg.hidden_size <= 768:
                # the UNMODIFIED reference call: model(ids, seg, mask, img_feats=f)[0] -> [B,S,V], then the caller's gather
                rows = min(B, 16)
                sub = {k: v[:rows].contiguous() for k, v in devb[0].items()}

                def ref_call():
                    s = rec(sub["input_ids"], sub["token_type_ids"], sub["attention_mask"], img_feats=sub["img_feats"])[0]
                    return s[torch.arange(rows, device=dev), sub["mask_pos"]][:, vids]
                for _ in range(2):
                    ref_call()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    ref_call()
                torch.cuda.synchronize()
                res["unmodified_call"] = {"samples_per_s": rows * 5 / (time.perf_counter() - t0), "rows_per_step": rows,
                                          "what": "model(ids, seg, mask, img_feats=f)[0] -> full [B,S,V] scores, then the "
                                                  "caller's gathers (zeroshot/refcoco_cpt.py:217-219,234-235): what a "
                                                  "maintainer gets without the two-keyword edit"}
        res["eng"], res["vids_cpu"] = eng, vids_cpu
        res["call_local"] = lambda i: call(devb[i % NROT])  # no collective: the roofline leg runs on rank 0 alone
    return res


def roofline_leg(args, w, cfg, res):
    """Per-kernel-class CUDA-event timing of the same steps (eager launches, rank 0); the dominant kernel against the
    measured sustained tensor peak."""
    eng, B, S = res["eng"], res["batch_per_gpu"], w["T"] + w["R"]
    with torch.no_grad():
        eng.profile(True)  # graph replay is bypassed while profiling: launches are bracketed by CUDA events
        for i in range(args.steps):
            res["call_local"](i)
        prof = eng.profile_read()
        eng.profile(False)
    M, H, I, L = B * S, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    gflop = {"gemm_qkv": 2.0 * M * 3 * H * H, "gemm_attn_out": 2.0 * M * H * H, "gemm_ffn_up": 2.0 * M * I * H,
             "gemm_ffn_down": 2.0 * M * H * I, "gemm_img": 2.0 * B * w["R"] * cfg.img_feature_dim * H,
             "attention": 4.0 * B * S * S * H,
             # one chain launch = attention-out + FFN-up + FFN-down + (all but the last layer) the next QKV projection
             "chain": 2.0 * M * (H * H + 2 * H * I + 3 * H * H * (L - 1) / L)}
    kernels = {}
    for name, (kms, n) in prof.items():
        d = {"ms_per_step": kms / args.steps, "launches_per_step": n / args.steps, "us_per_launch": 1e3 * kms / n}
        if name in gflop:
            d["tflops"] = gflop[name] * n / (kms * 1e-3) / 1e12
        kernels[name] = d
    cand = [k for k in kernels if k in ("chain", "gemm_ffn_up", "gemm_ffn_down", "gemm_qkv", "gemm_attn_out")]
    dom = max(cand, key=lambda k: kernels[k]["ms_per_step"])
    burst, sustained, hbm, how = peaks()
    ach = kernels[dom]["tflops"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom)
    kname = {"chain": "chain2_kernel (attention-out + FFN-up + FFN-down + next QKV: one dataflow launch per layer)"}
    roof = {"kernel": kname.get(dom, "gemm_kernel (%s)" % dom), "bound": "tensor", "achieved": ach, "peak": sustained,
            "unit": "TFLOP/s", "frac": ach / sustained, "frac_of_burst_peak": ach / burst,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s); kernel timed inside a long step" % how,
            "flops_per_launch": gflop[dom], "us_per_launch": kernels[dom]["us_per_launch"], "traffic": traffic}
    return roof, kernels


def run_infer(args, w, rank, local_rank, world, dist, dev):
    cfg = model_cfg(w)
    sd = synth_state_dict(cfg, seed=88)
    res = infer_leg(args, w, cfg, sd, dev, rank, world, dist, args.dtype, want_extras=not args.no_extras)
    if res is None or rank != 0:
        return None
    B, T, R, K = res["batch_per_gpu"], w["T"], w["R"], w["K"]
    total = B * world * args.steps
    value = total / (res["ms"] * 1e-3)
    fl = flops_per_sample(cfg, T, R, K)
    burst, sustained, hbm, how = peaks()
    roof, kernels = roofline_leg(args, w, cfg, res)
    out = {"metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": res["ms"] / args.steps, "higher_is_better": True,
           "scaling": w["scaling"], "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16"}[args.dtype],
           "data": "synthetic",
           "config": {"workload": "%s: %s, %d rows per GPU per step, T=%d text tokens + R=%d regions x 2054-d%s"
                                  % (w["name"], "Oscar-large" if w["model"] == "large" else "Oscar-base", B, T, R,
                                     ", K=%d gathered vocabulary ids" % K if K else ", NSP head + per-question argmax"),
                      "batch_per_gpu": B, "seq_len": T + R,
                      "l2": "inputs rotate over distinct batches (%d MB together) and one step streams the 16-bit weights "
                            "plus ~3 MB of activations per row: larger than the 126 MB L2"
                            % ((8 if cfg.hidden_size <= 768 else 4) * res["h2d"] // 2 ** 20),
                      "parallelism": ("dp%d (rows sharded; the logits all-gather is fused into the head kernel: P2P stores into every "
                                      "rank's buffer over NVLink + one flag per peer, inside the replayed graph)" % world
                                      if args.collective == "peer" else
                                      "dp%d (rows sharded, one fixed-shape NCCL all-gather of the logits, no host sync)" % world)},
           "e2e": {"value": total / res["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                   "d2h_bytes_per_step": res["d2h"]},
           "gpu_launches": res["launches"],
           "host_us_per_step": res["host_us_per_step"],
           "algorithmic_gflop_per_sample": fl / 1e9,
           "model_tflops": value * fl / 1e12,
           "model_frac_of_sustained_peak": value * fl / 1e12 / (sustained * world),
           "clocks": res["clocks"], "roofline": roof, "kernels": kernels}
    for k in ("fresh_tensor_loop", "unmodified_call"):
        if k in res:
            out[k] = res[k]
    try:
        out["parity"] = parity_block(w, cfg, sd, res["vids_cpu"], res["last_batch"], res["last_out"])
    except Exception as e:  # noqa: BLE001
        out["parity"] = {"error": str(e)[:200]}
    # the other operand type beside the headline one (BASELINE.json says bf16; see DESIGN.md "Precision")
    if world == 1 and not args.no_extras:
        other = "bf16" if args.dtype == "fp16" else "fp16"
        try:
            del res
            torch.cuda.empty_cache()
            r2 = infer_leg(args, w, model_cfg(w), sd, dev, rank, world, dist, other, want_extras=False)
            o = {"dtype": {"fp16": "f16", "bf16": "bf16"}[other], "value": B * args.steps / (r2["ms"] * 1e-3),
                 "ms_per_step": r2["ms"] / args.steps,
                 "e2e": {"value": B * args.steps / r2["e2e_s"], "unit": UNIT}}
            o["parity"] = parity_block(w, cfg, sd, r2["vids_cpu"], r2["last_batch"], r2["last_out"])
            out["other_dtype"] = o
        except Exception as e:  # noqa: BLE001
            out["other_dtype"] = {"error": str(e)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        vids = synth_vocab_ids(cfg, max(K, 1), seed=88)
        bs, it = {2: (32, 4), 4: (8, 3), 5: (4, 2)}[args.config]
        v, per, kind, what = cpu_leg(w, cfg, sd, vids, bs, it)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                               "sample": "%d timed passes (1 warm-up) of a %d-row batch of the same workload, %.2f s per "
                                         "pass; %s" % (it, bs, per, what)}
    return out


class _null(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run_train(args, w, rank, local_rank, world, dist, dev):
    """configs[2]: the GQA few-shot step (Oscar/oscar/fewshot/gqa_cpt.py:428-462, cmds/gqa/_cpt_fsl_base.sh:18-32)."""
    from cpt_b200.optimization import AdamW, WarmupLinearSchedule
    cfg = model_cfg(w)
    cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob = 0.3, 0.1  # _cpt_fsl_base.sh: --drop_out 0.3
    cfg.cpt_b200_train_dtype = args.dtype
    sd = synth_state_dict(cfg, seed=88)
    T, R, K = w["T"], w["R"], w["K"]
    S = T + R
    B = args.batch if args.batch > 0 else w["batch"]
    model, _ = build_models(cfg, sd, dev)
    ddp = None
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel
        ddp = DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)
        comm.enable_overlapped_grad_sync(ddp)  # gradient all-reduce inside the native backward
    net = ddp if ddp is not None else model
    no_decay = ["bias", "LayerNorm.weight"]
    groups = [{"params": [p for n, p in model.named_parameters() if not any(nd in n for nd in no_decay)], "weight_decay": 0.05},
              {"params": [p for n, p in model.named_parameters() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = AdamW(groups, lr=5e-5, eps=1e-8)
    sched = WarmupLinearSchedule(opt, warmup_steps=0, t_total=10 ** 6)
    vids = synth_vocab_ids(cfg, K, seed=88)
    NROT = 8
    host = [synth_batch(cfg, B, T, R, seed=2000 + rank * 100 + i) for i in range(NROT)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    labels = [train_labels(b, vids, S).to(dev) for b in host]

    def step(i):
        b = devb[i % NROT]
        net.train()
        loss = net(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["attention_mask"],
                   img_feats=b["img_feats"], masked_lm_labels=labels[i % NROT])[0]
        loss.backward()
        opt.step(max_grad_norm=1.0)  # clip_grad_norm_(1.0) fused into the native AdamW launch (gqa_cpt.py:454-456)
        sched.step()
        model.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            loss = step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, loss

    with torch.enable_grad():
        for i in range(max(args.warmup, 4)):
            step(i)
        if args.profile_only:
            for i in range(args.steps):
                step(i)
            torch.cuda.synchronize()
            return None
        eng = model.bert.train_engine()[0]
        sampler = ClockSampler(dev.index) if rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        l0 = eng.launch_count()
        ms, loss = timed(args.steps)
        launches = eng.launch_count() - l0
        clocks = sampler.stop() if sampler else None
        n2 = max(3, args.steps // 2)
        ms_nosync = None
        if world > 1:   # the same step with the exchange switched off (every rank keeps its local gradients)
            grp, eng.grad_sync_group = eng.grad_sync_group, None
            for i in range(3):
                step(i)
            ms_nosync = timed(n2)[0] / n2
            eng.grad_sync_group = grp
        # end to end: host (pinned) batch in, loss value out, every step
        hp = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        lp = [x.cpu().pin_memory() for x in labels]
        h2d = sum(v.numel() * v.element_size() for v in hp[0].values()) + lp[0].numel() * 8
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            devb[i % NROT] = {k: v.to(dev, non_blocking=True) for k, v in hp[i % NROT].items()}
            labels[i % NROT] = lp[i % NROT].to(dev, non_blocking=True)
            float(step(i).detach())  # the loss value is read on the host, as the reference's logging does
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
    if rank != 0:
        return None
    total = B * world * args.steps
    value = total / (ms * 1e-3)
    fl = 3 * flops_per_sample(cfg, T, R, K) + 2 * 3 * cfg.hidden_size * cfg.vocab_size  # fwd + dgrad + wgrad (+ full-vocab row)
    burst, sustained, hbm, how = peaks()
    out = {"metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": {"fp16": "f16", "bf16": "bf16"}[args.dtype], "data": "synthetic",
           "config": {"workload": "%s: Oscar-base, micro-batch %d per GPU, T=%d + R=%d (S=%d), K=%d answer ids, dropout "
                                  "0.3 / 0.1; step = native forward with tape + backward (gradient all-reduce inside it) + "
                                  "fused global-norm clip + native AdamW (cpt_adamw_step)" % (w["name"], B, T, R, S, K),
                      "batch_per_gpu": B, "seq_len": S,
                      "l2": "a step streams 223 MB of 16-bit weights, 447 MB of fp32 master weights and optimizer state: "
                            "far larger than the 126 MB L2",
                      "parallelism": "dp%d (DistributedDataParallel semantics; one NCCL all-reduce per gradient group, "
                                     "issued from inside the backward)" % world},
           "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
           "gpu_launches": launches, "loss": float(loss.detach()),
           "algorithmic_gflop_per_sample": fl / 1e9, "model_tflops": value * fl / 1e12,
           "model_frac_of_sustained_peak": value * fl / 1e12 / (sustained * world), "clocks": clocks}
    if ms_nosync is not None:
        out["nccl"] = {"ms_per_step_without_gradient_exchange": ms_nosync,
                       "share_of_step": max(0.0, 1.0 - ms_nosync / (ms / args.steps)),
                       "exchange_dtype": str(eng.grad_sync_dtype or "torch.float32"),
                       "how": "the same step with the in-backward all-reduce switched off (local gradients)"}
    out["roofline"] = {"kernel": "whole training step (forward + backward + optimizer)", "bound": "tensor",
                       "achieved": value * fl / 1e12 / world, "peak": sustained, "unit": "TFLOP/s",
                       "frac": value * fl / 1e12 / (sustained * world),
                       "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % how,
                       "flops_per_launch": None, "us_per_launch": None, "traffic": None,
                       "note": "a micro-batch of 4 x 210 rows is latency-bound (about 300 launches on 840-row operands per "
                               "step): the fraction is reported against the algorithmic 3 x forward FLOPs"}
    if world == 1 and not args.no_cpu_baseline:
        v, per, kind, what = cpu_leg(w, cfg, sd, vids, 4, 2)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                               "sample": "2 timed passes (1 warm-up) of a 4-row micro-batch, %.2f s per pass; %s" % (per, what)}
    return out


def train_leg_quick(rec, cfg, batch, dev, steps=10, warmup=3):
    """RefCOCO few-shot step at the bench shape (B=64, dropout 0.1): reported next to the inference headline."""
    from cpt_b200.optimization import AdamW
    orig = (cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.1
    B, S = batch["input_ids"].shape[0], batch["attention_mask"].shape[1]
    labels = torch.full((B, S), -1, dtype=torch.long, device=dev)
    labels[torch.arange(B, device=dev), batch["mask_pos"]] = 2000 + torch.arange(B, device=dev) % 7
    opt = AdamW([p for p in rec.parameters() if p.requires_grad], lr=1e-6, eps=1e-8, torch_semantics=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = rec(batch["input_ids"], batch["token_type_ids"], batch["attention_mask"], img_feats=batch["img_feats"],
                   masked_lm_labels=labels)[0]
        loss.backward()
        opt.step()
        return loss

    rec.train()
    try:
        with torch.enable_grad():
            for _ in range(warmup):
                step()
            eng = rec.bert.train_engine()[0]
            l0 = eng.launch_count()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return {"workload": "REC_MLM_CPT few-shot training step: native forward with tape + backward + native AdamW "
                                "(cpt_adamw_step), Oscar-base, batch %d, S=%d, dropout 0.1, bf16 GEMM operands / fp32 "
                                "master weights" % (B, S),
                    "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "steps": steps, "warmup": warmup,
                    "gpu_launches_per_step": (eng.launch_count() - l0) / steps, "loss": float(loss.detach())}
    finally:
        rec.eval()
        opt.zero_grad(set_to_none=True)
        cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob = orig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config, numbered from 1 (2 = configs[1], the headline)")
    ap.add_argument("--dtype", default=None, choices=["fp16", "bf16"],
                    help="tensor-core operand type (default: fp16 for inference, bf16 for the training step)")
    ap.add_argument("--batch", type=int, default=0, help="rows per GPU per step (0 = the config's own)")
    ap.add_argument("--ref-batch", type=int, default=0, help="rows per step of the CPU reference arm (0 = per config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the second dtype / fresh-tensor / unmodified-call legs")
    ap.add_argument("--collective", choices=["peer", "nccl"], default="peer",
                    help="N > 1 inference: how the per-rank logits are gathered (fused peer-memory exchange | NCCL)")
    ap.add_argument("--profile-only", action="store_true", help="run warmup+steps once without the extra legs (ncu)")
    args = ap.parse_args()
    w = WORKLOADS[args.config]
    if args.dtype is None:
        args.dtype = "bf16" if w["kind"] == "train" else "fp16"
    if args.impl == "reference":
        return run_reference(args, w)

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
        # ... and whatever a library still writes to fd 1 goes to stderr: the JSON line is written to the saved descriptor
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group(backend="nccl", device_id=dev)

    if w["kind"] == "train":
        out = run_train(args, w, rank, local_rank, world, dist, dev)
    else:
        out = run_infer(args, w, rank, local_rank, world, dist, dev)
        if out is not None and args.config == 2 and world == 1 and not args.no_extras:
            # the few-shot training step at the bench shape: reported next to the headline
            try:
                cfg = model_cfg(w)
                rec, _ = build_models(cfg, synth_state_dict(cfg, seed=88), dev)
                b = {k: v.to(dev) for k, v in synth_batch(cfg, w["batch"], w["T"], w["R"], seed=1000).items()}
                out["train_step"] = train_leg_quick(rec, cfg, b, dev)
            except Exception as e:  # never let the extra leg take the headline line down
                out["train_step"] = {"error": str(e)[:200]}
    if rank == 0 and out is not None:
        line = json.dumps(out) + "\n"
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, line.encode())
        else:
            sys.stdout.write(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
