#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/chain_trace.py --stages upd --json gpurun_out/trace_upd.json > /dev/null 2>&1
timeout 120 python tools/chain_trace.py --stages qkvd --json gpurun_out/trace_qkvd.json > /dev/null 2>&1
python - <<'PY'
import json
for name in ("upd","qkvd"):
    d=json.load(open("gpurun_out/trace_%s.json"%name))
    ev=d["ev"]
    for p in (0,37):
        print(name,"pair",p)
        for rec in ev[p]:
            if rec[8]==0 and rec[5]==0: continue
            us=[round(x/1965.0,2) for x in rec[:8]]+[round(rec[9]/1965.0,2)]
            print("  task %4d: dep0 %.2f dep1 %.2f loads_issued %.2f | mma_start %.2f mma_issued %.2f | epi_begin %.2f acc_ready %.2f published %.2f work_done %.2f" % tuple([rec[8]&0xFFFFFF]+us))
PY
