#!/bin/bash
# Round-2 evidence on one B200: ncu counters of the per-batch kernels of a bench step (only the matching kernels are
# instrumented: the null-model setup launches tens of thousands of library kernels), the full counter set + source of
# the solve kernel, compute-sanitizer on small scans.  Every step is bounded.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --scaling weak --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
K='regex:i8_rotate_kernel|solve_lane_kernel|prefix_|decode_int8_kernel|row_ssq_kernel|compact_kernel'
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k "$K" -c 12 --csv --log-file gpurun_out/r2_ncu_metrics.csv $BENCH > gpurun_out/r2_ncu_metrics.out 2>&1
echo "ncu metrics rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_lane_kernel -c 1 -o gpurun_out/r2_solve_full $BENCH > gpurun_out/r2_solve_full.out 2>&1
echo "ncu full rc=$?"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_scan.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
tail -4 gpurun_out/r2_sanitizer_memcheck.log
