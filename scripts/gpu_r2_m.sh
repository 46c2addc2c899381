#!/bin/bash
# round 2, one GPU: sort / BH tests after a sort change, signal workload, merge bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_merge.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras signal > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python scripts/bench_print.py gpurun_out/r2m_bench.json || tail -20 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2m_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
for e in d["extra"]:
    print({k:(round(v["ms_per_step"],3), v.get("launches_per_step")) for k,v in e.get("kernels",{}).items()})
PY
timeout 600 python scripts/merge_bench.py 3000000 > gpurun_out/r2m_merge_bench.log 2>&1; tail -14 gpurun_out/r2m_merge_bench.log
