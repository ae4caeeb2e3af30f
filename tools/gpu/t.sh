#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_attn.py 2>&1 | tail -11
python - <<'PY'
import re
s=open('tools/trace_gemm.py').read()
s=s.replace('cases = [("qkv", M, 2304, 768, 0, False, [1256, 2256]), ("ffn_up", M, 3072, 768, 1, False, [2256])]','cases = [("qkv", M, 2304, 768, 0, False, [2256]), ("ffn_up", M, 3072, 768, 1, False, [2256]), ("ffn_down", M, 768, 3072, 0, True, [2192]), ("attn_out", M, 768, 768, 0, True, [1192])]')
open('/tmp/trace_gemm2.py','w').write(s)
PY
cp /tmp/trace_gemm2.py tools/_trace_gemm2.py
timeout 300 python tools/_trace_gemm2.py 2>&1 | tail -5
