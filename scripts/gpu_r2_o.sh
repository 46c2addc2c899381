#!/bin/bash
# round 2, N GPUs ($1): does giving every rank its own cores help?  pass timeline without and with FHC_PIN_CORES=1, bench with
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for P in 0 1; do
  FHC_PIN_CORES=$P timeout 600 $TR --master-port $((29540+P)) scripts/host_timeline.py > gpurun_out/r2o_host_timeline_n${N}_pin$P.log 2>&1
  echo "FHC_PIN_CORES=$P"; grep -A9 "^world" gpurun_out/r2o_host_timeline_n${N}_pin$P.log
done
FHC_PIN_CORES=1 timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 2 --extras "" > gpurun_out/r2o_bench_n${N}_pin1.json 2> gpurun_out/r2o_bench_n${N}_pin1.err
python scripts/bench_print.py gpurun_out/r2o_bench_n${N}_pin1.json || tail -30 gpurun_out/r2o_bench_n${N}_pin1.err
