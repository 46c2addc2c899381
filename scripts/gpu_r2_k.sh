#!/bin/bash
# round 2, N GPUs ($1): bench at N (+ the multi-GPU tests at N = 2, the host fabric microbenchmark at N = 8)
N=${1:-2}
mkdir -p gpurun_out
nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_gpu.log 2>&1
  echo "pytest rc $?" >> gpurun_out/r2k_pytest_gpu.log
  tail -4 gpurun_out/r2k_pytest_gpu.log
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:"pval_|hist_distance|bh_|radix_|fill_f64|lbeta_|outlier|digest|gather_ne|mid_range|scatter" -c 400 --csv \
      --log-file gpurun_out/launches_r02_final.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2k_ncu_list.log 2>&1
  tail -2 gpurun_out/r2k_ncu_list.log
  timeout 900 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
  python scripts/bench_print.py gpurun_out/r2k_bench_n1.json || tail -30 gpurun_out/r2k_bench_n1.err
  timeout 900 python bench.py --impl reference > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; tail -c 600 gpurun_out/r2k_bench_ref.json
  exit 0
fi
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_multi.log 2>&1
  echo "pytest rc $?" >> gpurun_out/r2k_pytest_multi.log
  tail -4 gpurun_out/r2k_pytest_multi.log
  timeout 600 $TR --master-port 29531 scripts/multi_gpu_check.py > gpurun_out/r2k_multi_gpu_check_n$N.log 2>&1
  echo "multi_gpu_check rc $?"; tail -2 gpurun_out/r2k_multi_gpu_check_n$N.log
fi
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 3 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err
python scripts/bench_print.py gpurun_out/r2k_bench_n$N.json || tail -30 gpurun_out/r2k_bench_n$N.err
timeout 600 $TR --master-port 29535 scripts/host_timeline.py > gpurun_out/r2k_host_timeline_n$N.log 2>&1; grep -A20 "^world" gpurun_out/r2k_host_timeline_n$N.log
if [ "$N" = "8" ]; then
  timeout 300 $TR --master-port 29537 scripts/pcie_bench.py 512 > gpurun_out/r2k_pcie_n$N.log 2>&1; grep -A12 "^world" gpurun_out/r2k_pcie_n$N.log
fi
