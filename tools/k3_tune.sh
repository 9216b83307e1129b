#!/bin/bash
# Build K3 with different tuning knobs on the GPU box and time the solve stage (development tool).
# usage: tools/k3_tune.sh "<flags cfg 1>" "<flags cfg 2>" ...
mkdir -p gpurun_out
if [ $# -eq 0 ]; then set -- "-DJXB_K3_BUFS=2 -DJXB_K3_MINB=2" "-DJXB_K3_BUFS=2 -DJXB_K3_MINB=1"; fi
for cfg in "$@"; do
  JXB_K3_FLAGS="$cfg" python janusx_b200/build.py --force > gpurun_out/k3_build.log 2>&1 || { echo "build failed: $cfg"; tail -5 gpurun_out/k3_build.log; continue; }
  echo "== $cfg"
  PROBE_NS=${PROBE_NS:-20000} PROBE_ROWS=${PROBE_ROWS:-16384} PROBE_VARIANTS=${PROBE_VARIANTS:-3} timeout 300 python tools/gpu_probe.py 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    parts=line.split(' ',1)
    if parts[0].isdigit():
        d=json.loads(parts[1]); v=d.get('variant3') or d.get('variant2'); print('n',parts[0],'solve_ms', round(v['stage_ms']['solve'],2), 'rotate_ms', round(v['stage_ms']['rotate'],2), 'snps/s', round(v['snps_per_s']))
"
done 2>&1 | tee gpurun_out/k3_tune.txt
