#!/bin/bash
# round 2, one GPU: ncu launch list of the bench command, full capture of the main kernels on the main workload
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2j_ncu_list.log 2>&1
tail -2 gpurun_out/r2j_ncu_list.log
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"pval_front|pval_iterate|pval_finish|hist_distance|bh_compact|bh_cut_hist|fill_f64" -s 21 -c 7 \
    -o gpurun_out/ncu_full_r02_main python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2j_ncu_full.log 2>&1
tail -2 gpurun_out/r2j_ncu_full.log
FHC_PREPASS=1 timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:"pval_front|pval_prepass" -s 6 -c 2 \
    -o gpurun_out/ncu_full_r02_prepass python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras "" > gpurun_out/r2j_ncu_pre.log 2>&1
ls -la gpurun_out/*r02*
