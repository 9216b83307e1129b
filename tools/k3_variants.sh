#!/bin/bash
# Rebuild the K3 translation units with different occupancy / tile settings on the GPU box and bench each (weak mode,
# one batch per step).  usage: tools/k3_variants.sh <tag>
set -u
tag=${1:-r2}
mkdir -p gpurun_out
i=0
for v in "-DJXB_K3T_MINB=3" "-DJXB_K3L_TILE=16" "-DJXB_K3T_MINB=3 -DJXB_K3L_TILE=16" ""; do
  JXB_K3_FLAGS="$v" python -m janusx_b200.build --force > gpurun_out/${tag}_build_$i.log 2>&1
  echo "variant $i: '$v'" >> gpurun_out/${tag}_variants.txt
  timeout 400 python bench.py --scaling weak --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_var_$i.json 2> gpurun_out/${tag}_var_$i.err
  python - <<PY >> gpurun_out/${tag}_variants.txt
import json
try:
    d = json.loads(open("gpurun_out/${tag}_var_$i.json").read().strip().splitlines()[-1])
    print("   value", round(d["value"]), "ms_per_step", round(d["ms_per_step"], 1), d["stage_ms_last_step_rank0"])
except Exception as e:
    print("   failed", e)
PY
  i=$((i+1))
done
cat gpurun_out/${tag}_variants.txt
