#!/bin/bash
# Round-2 evidence on one B200: ncu launch list of a bench step, full counters of the top kernels, compute-sanitizer
# on the smoke path.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
BENCH="python bench.py --scaling weak --steps 1 --warmup 1 --no-cpu-baseline"
# (1) every launch of the step with its device time (warm-up step skipped by taking the last launches)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $BENCH > gpurun_out/r2_launches.out 2>&1
# (2) targeted counters of the per-batch kernels of the timed step
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"i8_rotate|solve_lane|decode_int8|count_qc|row_ssq" -s 5 -c 5 --csv --log-file gpurun_out/r2_ncu_metrics.csv $BENCH > gpurun_out/r2_ncu_metrics.out 2>&1
# (3) full set + source for the solve kernel (one launch of the timed step)
ncu --set full --clock-control none --import-source on -k regex:solve_lane_kernel -s 1 -c 1 -o gpurun_out/r2_solve_full $BENCH > gpurun_out/r2_solve_full.out 2>&1
# (4) compute-sanitizer on the smoke path (tiny problem through decode -> DMMA rotation -> warp solve) and on a packed int8 scan
compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_scan.py > gpurun_out/r2_sanitizer_scan.log 2>&1
echo "memcheck(scan) rc=$?" >> gpurun_out/r2_sanitizer_scan.log
tail -3 gpurun_out/r2_sanitizer_*.log
