#!/bin/bash
# round 2, one GPU, final tree: full GPU tests, smoke(), bench (default flags, as the driver runs it), CLI wall time
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2n_pytest_gpu.log
tail -4 gpurun_out/r2n_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2n_smoke.log 2>&1; tail -3 gpurun_out/r2n_smoke.log
timeout 900 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
python scripts/bench_print.py gpurun_out/r2n_bench_n1.json || tail -30 gpurun_out/r2n_bench_n1.err
timeout 900 python scripts/cli_wall.py 5000000 10000 > gpurun_out/r2n_cli_wall.log 2>&1; grep -E "input written|CLI wall|wall_s|read_contacts_s|gpu_pass_s|write_s" gpurun_out/r2n_cli_wall.log
