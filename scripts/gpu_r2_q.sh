#!/bin/bash
# round 2: K3 / K1 experiments -- K3 and kernel tests, variants timing, ncu of the kernels that changed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_edge.py tests/test_gpu_golden.py -k "not sort and not partition and not bh_" -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2q_pytest.log; tail -6 gpurun_out/r2q_pytest.log
timeout 300 python scripts/k3_variants.py > gpurun_out/r2q_variants.log 2>&1; cat gpurun_out/r2q_variants.log
K3V_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pval_iterate|pval_finish" -s 4 -c 2 -o gpurun_out/ncu_k3b_r02 -f python scripts/k3_variants.py > gpurun_out/r2q_ncu.log 2>&1; tail -2 gpurun_out/r2q_ncu.log
