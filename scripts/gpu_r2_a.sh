#!/bin/bash
# round 2, first GPU call: host-stage timing on the box's cores, GPU parity tests, bench with 1 / 4 / 8 host threads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
lscpu > gpurun_out/r2a_lscpu.txt 2>&1
timeout 300 python scripts/host_stage_bench.py > gpurun_out/r2a_host_stage.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2a_pytest_gpu.log
for th in 8 4 1; do
  FHC_HOST_THREADS=$th timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2a_bench_t$th.json 2> gpurun_out/r2a_bench_t$th.err
done
FHC_HOST_STAGE=legacy timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2a_bench_legacy.json 2> gpurun_out/r2a_bench_legacy.err
tail -3 gpurun_out/r2a_pytest_gpu.log
cat gpurun_out/r2a_host_stage.log
python - <<'PY'
import json
for n in ("t8","t4","t1","legacy"):
    try:
        d=json.loads(open("gpurun_out/r2a_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f"%d["ms_per_step"], "e2e ms %.1f"%d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["host_ms_per_pass"].items()})
    except Exception as e:
        print(n,"failed",e)
PY
