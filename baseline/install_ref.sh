#!/bin/bash
# Offline install of the UNMODIFIED reference package into baseline/_ref (git-ignored; travels to the GPU box with
# gpurun).  Run in the build container only: /root/reference does not exist on the GPU box.
#   1. pip install of Oscar/ (setup.py's find_packages takes oscar, oscar.modeling, oscar.datasets, oscar.utils);
#   2. the task scripts oscar/zeroshot and oscar/fewshot have no __init__.py, so the wheel leaves them out: they are
#      placed next to the installed package so that `import oscar.zeroshot.refcoco_cpt` resolves (namespace
#      sub-packages).  tests/test_gpu_dropin.py runs their val()/evaluate() loops unmodified.
# Nothing under baseline/_ref is product source and nothing in cpt_b200/ imports it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=/root/reference/Oscar
[ -d "$SRC" ] || { echo "no $SRC here: keep the prebuilt baseline/_ref"; exit 0; }
TMP=$(mktemp -d)
cp -r "$SRC" "$TMP/oscar_src"          # the source tree is read-only; setup.py writes build/ beside itself
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --upgrade \
    --target "$ROOT/baseline/_ref" "$TMP/oscar_src"
for d in zeroshot fewshot; do
    rm -rf "$ROOT/baseline/_ref/oscar/$d"
    cp -r "$SRC/oscar/$d" "$ROOT/baseline/_ref/oscar/$d"
done
rm -rf "$TMP"
echo "installed: $(ls "$ROOT/baseline/_ref/oscar")"
