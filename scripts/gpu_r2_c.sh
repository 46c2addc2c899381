#!/bin/bash
# round 2, third GPU call (2 GPUs): parity incl. one-sweep sort and the 2-GPU tests, bench at N=1 (with extras) and N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2c_pytest_gpu.log
tail -4 gpurun_out/r2c_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_check.py > gpurun_out/r2c_multi_gpu_check.log 2>&1
echo "multi_gpu_check rc $?"; grep -c "p err" gpurun_out/r2c_multi_gpu_check.log; tail -3 gpurun_out/r2c_multi_gpu_check.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
python - <<'PY'
import json
for n in ("n1","n2"):
    try:
        d=json.loads([l for l in open("gpurun_out/r2c_bench_%s.json"%n).read().strip().splitlines() if l.startswith("{")][-1])
        print(n, "ms/step %.3f"%d["ms_per_step"], "e2e ms %.1f"%d["e2e"]["ms_per_step"], "digest", d.get("digest_line_p_q"), "launches", d["gpu_launches"])
        print("   host", {k:round(v,3) for k,v in d["host_ms_per_pass"].items()})
        print("   kern", {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
        print("   roofline", json.dumps(d["roofline"])[:600])
        for e in d.get("extra",[]):
            if "failed" in e: print("   extra", e); continue
            print("   extra %s: ms/step %.3f value %.3g q<1 %d sorted %d k4 %.3f digest %s"%(e["name"],e["ms_per_step"],e["value"],e["lines_with_q_below_1"],e["sorted_pairs_rank0"],e["k4_ms_per_step"],e["digest_line_p_q"]))
            print("        kern", {k:round(v["ms_per_step"],3) for k,v in list(e["kernels"].items())[:8]})
    except Exception as ex:
        print(n,"failed",ex); print(open("gpurun_out/r2c_bench_%s.err"%n).read()[-3000:])
PY
