#!/bin/bash
# round 2, 1 GPU: parity after the K1 rewrite, bench, ncu launch list + full capture on the signal workload + FP64 op counts
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2d_pytest_gpu.log
tail -4 gpurun_out/r2d_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --extras signal > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras signal > gpurun_out/r2d_ncu_list.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"pval_front|pval_iterate|pval_finish|hist_distance|radix_onesweep|bh_compact|bh_scatter|bh_cut_hist|bh_tilemax" -s 51 -c 17 \
    -o gpurun_out/ncu_full_r02 python bench.py --signal 0.08 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2d_ncu_full.log 2>&1
timeout 900 ncu --clock-control none --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
    -k regex:"pval_front|pval_iterate|pval_finish" -s 9 -c 3 --csv --log-file gpurun_out/fp64_ops_r02.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2d_ncu_fp64.log 2>&1
ls -la gpurun_out/ | grep r02
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2d_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("n1 ms/step %.3f"%d["ms_per_step"], "e2e ms %.1f"%d["e2e"]["ms_per_step"], "digest", d.get("digest_line_p_q"))
print("   host", {k:round(v,3) for k,v in d["host_ms_per_pass"].items()})
print("   kern", {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
for e in d.get("extra",[]):
    print("   extra", e.get("name"), e.get("ms_per_step"), e.get("k4_ms_per_step"), e.get("failed"))
PY
