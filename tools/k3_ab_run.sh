#!/bin/bash
# GPU side of tools/k3_ab_build.sh: bench every ab/libjxb200_v<i>.so (weak mode, one batch per step) and print the stage times.
# usage: tools/k3_ab_run.sh [tag] [extra bench flags]
set -u
cd "$(dirname "$0")/.."
tag=${1:-ab}; shift || true
mkdir -p gpurun_out
B="--scaling weak --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline $*"
: > gpurun_out/${tag}_results.txt
run() {   # name, lib path ('' = in-tree)
  JXB_LIB_PATH="$2" timeout 150 python bench.py $B > gpurun_out/${tag}_$1.json 2> gpurun_out/${tag}_$1.err
  python - "$1" "gpurun_out/${tag}_$1.json" <<'PY' | tee -a gpurun_out/${tag}_results.txt
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); s = d["stage_ms_last_step_rank0"]
    print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "rotate", s["rotate"], "solve", s["solve"], "MHz", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
for so in $(ls ab/libjxb200_v*.so | sort -V); do
  name=$(basename $so .so); name=${name#libjxb200_}
  run $name "$PWD/$so"
done
cat ab/variants.txt >> gpurun_out/${tag}_results.txt
