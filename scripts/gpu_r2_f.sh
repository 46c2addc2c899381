#!/bin/bash
# round 2: N GPUs ($1): full GPU tests at N=2, multi check, timeline, bench
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1
  echo "pytest rc $?" >> gpurun_out/r2f_pytest_gpu.log
  tail -4 gpurun_out/r2f_pytest_gpu.log
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 --extras "c3,signal" > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
  python scripts/bench_print.py gpurun_out/r2f_bench_n1.json
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_check.py > gpurun_out/r2f_multi_gpu_check_n$N.log 2>&1
echo "multi_gpu_check rc $?"; tail -2 gpurun_out/r2f_multi_gpu_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 scripts/host_timeline.py > gpurun_out/r2f_host_timeline_n$N.log 2>&1; grep -A20 "^world" gpurun_out/r2f_host_timeline_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
python scripts/bench_print.py gpurun_out/r2f_bench_n$N.json || tail -30 gpurun_out/r2f_bench_n$N.err
