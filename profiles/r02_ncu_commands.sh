set -x
M=gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
# 1. launch list of the bench command (full frontier; lane build), durations only
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_pre_ncu.json 2> gpurun_out/r2_bench_pre_ncu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
# 2. FP64 pipe / DRAM / issue counters of every kernel of one lap: lane build on the full frontier (short lap so ncu can replay)
ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r2_counters_lane_full.csv python bench.py --steps 1 --warmup 0 --lap-scale 0.05 --no-cpu-baseline > gpurun_out/r2_ncu_counters_lane.log 2>&1
# 3. the same for the warp build (8192 cells)
ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r2_counters_warp_8192.csv python bench.py --steps 1 --warmup 0 --cells 8192 --lap-scale 0.05 --no-cpu-baseline > gpurun_out/r2_ncu_counters_warp.log 2>&1
ls -la gpurun_out/r2_launches_full.csv gpurun_out/r2_counters_lane_full.csv gpurun_out/r2_counters_warp_8192.csv
