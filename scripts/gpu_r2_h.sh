#!/bin/bash
# round 2, N GPUs ($1): parity check under torchrun, host-thread sweep of the pass timeline, bench
N=${1:-8}
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 scripts/multi_gpu_check.py > gpurun_out/r2h_multi_gpu_check_n$N.log 2>&1
echo "multi_gpu_check rc $?"; tail -2 gpurun_out/r2h_multi_gpu_check_n$N.log
port=29540
for T in 8 3; do
  port=$((port+1))
  FHC_HOST_THREADS=$T timeout 600 $TR --master-port $port scripts/host_timeline.py > gpurun_out/r2h_host_timeline_n${N}_t$T.log 2>&1
  grep -A20 "^world" gpurun_out/r2h_host_timeline_n${N}_t$T.log
done
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 3 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err
python scripts/bench_print.py gpurun_out/r2h_bench_n$N.json || tail -30 gpurun_out/r2h_bench_n$N.err
