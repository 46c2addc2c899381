#!/bin/bash
# round 2: partial pre-pass check -- pipeline tests, bench (N = $1)
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_edge.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1
  echo "pytest rc $?" >> gpurun_out/r2p_pytest.log; tail -3 gpurun_out/r2p_pytest.log
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 --extras "" > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
  python scripts/bench_print.py gpurun_out/r2p_bench_n1.json || tail -20 gpurun_out/r2p_bench_n1.err
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 2 --extras "" > gpurun_out/r2p_bench_n$N.json 2> gpurun_out/r2p_bench_n$N.err
  python scripts/bench_print.py gpurun_out/r2p_bench_n$N.json || tail -20 gpurun_out/r2p_bench_n$N.err
fi
