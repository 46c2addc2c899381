#!/bin/bash
# round 2, one GPU: full GPU tests, bench, host stage timing, merge bench (N4), opcode mix of the main kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2i_pytest_gpu.log
tail -4 gpurun_out/r2i_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras "" > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
python scripts/bench_print.py gpurun_out/r2i_bench_n1.json || tail -30 gpurun_out/r2i_bench_n1.err
timeout 300 python scripts/host_stage_bench.py > gpurun_out/r2i_host_stage.log 2>&1; cat gpurun_out/r2i_host_stage.log
(timeout 600 python scripts/merge_bench.py 3000000; timeout 600 python scripts/merge_bench.py 3000000 --host) > gpurun_out/r2i_merge_bench.log 2>&1; tail -15 gpurun_out/r2i_merge_bench.log
