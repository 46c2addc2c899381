#!/bin/bash
# round 2, N GPUs (N = $1, default 2): the multi-GPU tests, the parity check under torchrun, bench at N
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2e_pytest_multi.log 2>&1
  echo "pytest rc $?" >> gpurun_out/r2e_pytest_multi.log
  tail -4 gpurun_out/r2e_pytest_multi.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_check.py > gpurun_out/r2e_multi_gpu_check_n$N.log 2>&1
echo "multi_gpu_check rc $?"; tail -2 gpurun_out/r2e_multi_gpu_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 scripts/host_timeline.py > gpurun_out/r2e_host_timeline_n$N.log 2>&1; grep -A20 "^world" gpurun_out/r2e_host_timeline_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 3 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
python - $N <<'PY'
import json, sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r2e_bench_n%s.json"%n).read().strip().splitlines() if l.startswith("{")][-1])
    print("n%s"%n, "ms/step %.3f"%d["ms_per_step"], "e2e ms %.1f"%d["e2e"]["ms_per_step"], "digest", d.get("digest_line_p_q"), "launches", d["gpu_launches"])
    print("   host", {k:round(v,3) for k,v in d["host_ms_per_pass"].items()})
    print("   kern", {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
    for e in d.get("extra",[]):
        if "failed" in e: print("   extra", e); continue
        print("   extra %s: ms/step %.3f value %.3g q<1 %d k4 %.3f digest %s"%(e["name"],e["ms_per_step"],e["value"],e["lines_with_q_below_1"],e["k4_ms_per_step"],e["digest_line_p_q"]))
except Exception as ex:
    print("failed",ex); print(open("gpurun_out/r2e_bench_n%s.err"%n).read()[-3000:])
PY
