# ncu captures committed under profiles/ (round 2).  Run on the GPU box: bash tools/capture_r02.sh
cd /root/repo
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
# 1. the headline kernel, one whole-batch launch in the stationary regime
ncu --set full --clock-control none --import-source on -k regex:rollout_persist --launch-skip 10 --launch-count 1 -o gpurun_out/prof_r02c_persist python tools/run_rollout.py 10 5 65536 20 4 8 > gpurun_out/ncu_r02c.log 2>&1
# 2. launch list of the bench's headline loop
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --burnin 160 --min-ms 5 --samples 2 --skip-e2e --skip-secondary --skip-cpu > gpurun_out/ncu_r02_launches.log 2>&1
# 3. generator kernels: instructions per board
ncu --metrics $M --clock-control none -k regex:prw_kernel --csv --log-file gpurun_out/r02_prw_metrics.csv python tools/run_prw.py 1 > gpurun_out/ncu_r02_prw.log 2>&1
# 4. SeedExtension pipeline
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_seedext_metrics.csv python tools/run_seedext_once.py > gpurun_out/ncu_r02_se.log 2>&1
# 5. per-step path: 12 steps
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_perstep_launches.csv python tools/run_perstep.py 12 > gpurun_out/ncu_r02_perstep.log 2>&1
echo captured
