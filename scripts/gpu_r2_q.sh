#!/bin/bash
# round 2: front kernel v2 (loads one group ahead, list positions per warp, rolled group loop) -- K3 tests, variants timing, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_edge.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2q_pytest.log; tail -4 gpurun_out/r2q_pytest.log
timeout 300 python scripts/k3_variants.py > gpurun_out/r2q_variants.log 2>&1; cat gpurun_out/r2q_variants.log
K3V_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pval_front2 -s 2 -c 1 -o gpurun_out/ncu_front2_r02 -f python scripts/k3_variants.py > gpurun_out/r2q_ncu.log 2>&1; tail -2 gpurun_out/r2q_ncu.log
