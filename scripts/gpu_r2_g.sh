#!/bin/bash
# round 2, one GPU: full GPU tests, bench (with extras), host pass profile, host stage timing
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|MHz" 
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2g_pytest_gpu.log
tail -4 gpurun_out/r2g_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras "" > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
python scripts/bench_print.py gpurun_out/r2g_bench_n1.json || tail -30 gpurun_out/r2g_bench_n1.err
timeout 300 python scripts/pass_profile.py > gpurun_out/r2g_pass_profile.log 2>&1; head -60 gpurun_out/r2g_pass_profile.log
timeout 300 python scripts/host_stage_bench.py > gpurun_out/r2g_host_stage.log 2>&1; cat gpurun_out/r2g_host_stage.log
timeout 300 python scripts/host_timeline.py > gpurun_out/r2g_host_timeline_n1.log 2>&1; grep -A20 "^world" gpurun_out/r2g_host_timeline_n1.log
FHC_SLOW=1 timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k config3_full_size > gpurun_out/r2g_c3_full_oracle.log 2>&1; tail -6 gpurun_out/r2g_c3_full_oracle.log
