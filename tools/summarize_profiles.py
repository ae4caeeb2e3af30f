"""Build profiles/rNN_summary.md and profiles/traffic.json from the files tools/gpu/profiles.sh leaves in gpurun_out/.
    python tools/summarize_profiles.py [r01]
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")

RAW_METRICS = [("duration", "gpu__time_duration.sum"), ("SM cycles", "sm__cycles_elapsed.avg"),
               ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
               ("L2 hit %", "lts__t_sector_hit_rate.pct"), ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
               ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               ("tensor pipe active % (of elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               ("regs/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
               ("dyn smem/CTA", "launch__shared_mem_per_block_dynamic"), ("warp instructions", "sm__inst_executed.sum"),
               ("L2->SM bytes", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum"),
               ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               ("L2 bytes (all traffic through the LTS)", "lts__t_bytes.sum"),
               ("L2 sectors read", "lts__t_sectors_op_read.sum"), ("L2 sectors written", "lts__t_sectors_op_write.sum"),
               ("L1/TEX+TMA -> SM throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("cptk::", "")


def launch_table(path):
    agg = OrderedDict()
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(row["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", "")) / 1e3
    return agg


def md_launches(agg, skip=()):
    total = sum(v[1] for k, v in agg.items() if not any(s in k for s in skip))
    out = ["| kernel | launches | total us | us each | share |", "|---|---|---|---|---|"]
    n = 0
    for k, (c, us) in agg.items():
        if any(s in k for s in skip):
            continue
        out.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, c, us, us / c, 100 * us / total))
        n += c
    out.append("| sum | %d | %.1f | | |" % (n, total))
    return "\n".join(out)


def raw_table(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None, []
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["| metric | " + " | ".join("k%d" % i for i in range(len(data))) + " |", "|---|" + "---|" * len(data)]
    out.append("| kernel | " + " | ".join("`%s`" % short(r[idx["Kernel Name"]]) for r in data) + " |")
    for label, m in RAW_METRICS:
        if m not in idx:
            continue
        u = units[idx[m]]
        out.append("| %s%s | " % (label, " (%s)" % u if u else "") + " | ".join(r[idx[m]] for r in data) + " |")
    return "\n".join(out), [(r, idx, units) for r in data]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(val.replace(",", "")) * mult


def main():
    out = ["# profiles/%s — evidence of this round (B200, one GPU; raw files: gpurun_out/%s_*, produced by "
           "`tools/gpu/profiles.sh` / `tools/gpu/r2_profiles.sh`, summarised by `tools/summarize_profiles.py`)\n" % (TAG, TAG)]
    bench = json.load(open(os.path.join(G, TAG + "_bench.json")))
    kernels = bench.pop("kernels", {})
    bench.pop("config", None)
    out.append("## bench.py (`python bench.py --steps 200 --warmup 10`)\n\n```json\n%s\n```\n" % json.dumps(bench, indent=1))
    refp = os.path.join(G, TAG + "_bench_reference.json")
    if os.path.exists(refp):
        ref = json.load(open(refp))
        out.append("Reference arm (`--impl reference --steps %d --warmup %d`): %.1f samples/s on %d host cores, kind `%s`.\n"
                   % (ref["steps"], ref["warmup"], ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["kind"]))
    out.append("Per-kernel-class CUDA-event timing inside bench.py (eager launches, `cpt_profile_*`):\n")
    out.append("| kernel class | ms/step | launches/step | us/launch | TFLOP/s |\n|---|---|---|---|---|")
    for k, v in kernels.items():
        out.append("| %s | %.3f | %d | %.1f | %s |" % (k, v["ms_per_step"], v["launches_per_step"], v["us_per_launch"],
                                                      "%.0f" % v["tflops"] if "tflops" in v else ""))
    out.append("\n## ncu launch list of `bench.py --profile-only --steps 1 --warmup 1` (3 forwards; `ncu --metrics gpu__time_duration.sum --clock-control none`, graphs "
               "off; cold-cache, serialised: compare SHARES)\n")
    out.append(md_launches(launch_table(os.path.join(G, TAG + "_launches.csv")), skip=("cast_weight", "fold_weight")))
    out.append("\n## ncu --set full (one capture per kernel; `--clock-control none`; ncu flushes caches between replays, "
               "so DRAM traffic is the cold-cache figure)\n")
    traffic = {}
    for title, name in (("dataflow chain kernel: attention-out + residual, FFN-up (LayerNorm folded) + GELU, FFN-down + "
                         "residual, next layer's QKV (LayerNorm folded) in ONE launch; two consecutive layers", "chain"),
                        ("GEMM kernels of one layer (QKV, attention-out, FFN-up, FFN-down)", "gemm"),
                        ("attention forward", "attn"), ("LayerNorm", "ln"),
                        ("attention backward (training, S=120)", "attn_bwd"),
                        ("backward GEMMs of one encoder layer at B=64 (in launch order: FFN-down wgrad [trans=3, "
                         "split-K] and dgrad [trans=2], FFN-up wgrad and dgrad, attention-out wgrad and dgrad, merged QKV "
                         "wgrad and dgrad; captured one commit before the tile choice became split-K aware — k4 ran 64-wide tiles here, "
                         "192-wide since)", "bwd_gemm")):
        p = os.path.join(G, "%s_%s_raw.csv" % (TAG, name))
        if not os.path.exists(p):
            continue
        tbl, data = raw_table(p)
        if tbl is None:
            continue
        out.append("### %s\n\n%s\n" % (title, tbl))
        if name == "chain" and data:
            r, idx, units = data[0]
            rd, wr = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"]
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            traffic = json.load(open(tp)) if os.path.exists(tp) else {}
            traffic["chain"] = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
        if name == "gemm":
            for cls, (r, idx, units) in zip(("gemm_qkv", "gemm_attn_out", "gemm_ffn_up", "gemm_ffn_down"), data):
                rd, wr = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"]
                traffic[cls] = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
    if traffic:
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
        out.append("`profiles/traffic.json` (dram read+write bytes per launch, read by bench.py for `roofline.traffic`): "
                   "%s\n" % json.dumps(traffic))
    tb = os.path.join(G, TAG + "_train_bench.jsonl")
    if os.path.exists(tb):
        out.append("## Training step (`tools/train_bench.py`: forward with tape + backward + fused AdamW, Oscar-base, "
                   "synthetic data; CUDA events)\n")
        out.append("| workload | ms/step | samples/s | forward (+16-bit weight refresh) | backward | AdamW |\n|---|---|---|---|---|---|")
        recs = [json.loads(l) for l in open(tb) if l.strip().startswith("{")]
        for r in recs:
            s = r["split_ms"]
            out.append("| %s | %.2f | %.0f | %.2f | %.2f | %.2f |" % (r["workload"].replace("REC_MLM_CPT train step, ", ""),
                       r["ms_per_step"], r["samples_per_s"], s["forward(+weight refresh)"], s["backward"], s["adamw"]))
        for r in recs[:2]:
            out.append("\nBackward by kernel class, %s (ms, launches): %s" % (
                r["workload"].replace("REC_MLM_CPT train step, ", ""),
                ", ".join("%s %.2f (%d)" % (k, v[0], v[1]) for k, v in r["backward_kernels_ms"].items())))
    tl = os.path.join(G, TAG + "_train_launches.csv")
    if os.path.exists(tl):
        out.append("\n### ncu launch list of one training step (B=64, S=120; weight refresh + forward + backward + AdamW)\n")
        out.append(md_launches(launch_table(tl)))
    open(os.path.join(ROOT, "profiles", TAG + "_summary.md"), "w").write("\n".join(out) + "\n")
    print("wrote profiles/%s_summary.md" % TAG)


if __name__ == "__main__":
    main()
