#!/bin/bash
# round 2, second GPU call: parity after the chromosome-run / specialised front kernel, bench with and without the runs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2b_pytest_gpu.log
tail -4 gpurun_out/r2b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --extras "" > gpurun_out/r2b_bench_runs.json 2> gpurun_out/r2b_bench_runs.err
FHC_CHR_RUNS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --extras "" > gpurun_out/r2b_bench_chrs.json 2> gpurun_out/r2b_bench_chrs.err
python - <<'PY'
import json
for n in ("runs","chrs"):
    try:
        d=json.loads(open("gpurun_out/r2b_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f"%d["ms_per_step"], "e2e ms %.1f"%d["e2e"]["ms_per_step"], "digest", d.get("digest_line_p_q"))
        print("   host", {k:round(v,3) for k,v in d["host_ms_per_pass"].items()})
        print("   kern", {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n,"failed",e); print(open("gpurun_out/r2b_bench_%s.err"%n).read()[-2000:])
PY
