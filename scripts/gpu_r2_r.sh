#!/bin/bash
# round 2, one GPU, final tree: full GPU tests, smoke(), bench (default flags, as the driver runs it), ncu (full set of the
# main kernels + FP64 operation counts, launch list)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2r_pytest_gpu.log
tail -4 gpurun_out/r2r_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2r_smoke.log 2>&1; tail -2 gpurun_out/r2r_smoke.log
timeout 600 python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
python scripts/bench_print.py gpurun_out/r2r_bench_n1.json || tail -30 gpurun_out/r2r_bench_n1.err
FP64=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum
FHC_PREPASS=0 timeout 300 ncu --set full --metrics $FP64 --clock-control none --import-source on \
    -k regex:"pval_front|pval_iterate|pval_finish|hist_distance|bh_compact|bh_cut_hist|fill_f64" -s 21 -c 7 \
    -o gpurun_out/ncu_full_r02_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" \
    > gpurun_out/r2r_ncu_full.log 2>&1; tail -2 gpurun_out/r2r_ncu_full.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"pval_|hist_distance|bh_|radix_|fill_f64|lbeta_|outlier|digest|gather_ne|mid_range|scatter|onesweep" -c 400 --csv \
    --log-file gpurun_out/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2r_ncu_list.log 2>&1
tail -2 gpurun_out/r2r_ncu_list.log
