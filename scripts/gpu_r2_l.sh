#!/bin/bash
# round 2, one GPU: compute-sanitizer over the GPU tests (memcheck with every tensor its own allocation, racecheck)
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_r02.txt
echo "compute-sanitizer on B200 (round 2)" > $OUT
echo >> $OUT
echo "memcheck, PYTORCH_NO_CUDA_MEMORY_CACHING=1 (every tensor its own cudaMalloc, so an out-of-bounds access cannot hide in the caching allocator):" >> $OUT
echo "  python -m pytest tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_pipeline.py tests/test_hickry.py tests/test_gpu_kernels.py tests/test_merge.py -m gpu -x -q" >> $OUT
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 \
    python -m pytest tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_pipeline.py tests/test_hickry.py tests/test_gpu_kernels.py tests/test_merge.py -m gpu -x -q > gpurun_out/r2l_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|COMPUTE-SANITIZER$" gpurun_out/r2l_memcheck.log | tail -4 >> $OUT
grep -B2 -A12 "Invalid\|out of bounds\|misaligned" gpurun_out/r2l_memcheck.log | head -60 >> $OUT
echo >> $OUT
echo "racecheck (shared-memory hazards: front / iterate / one-sweep sort / cut kernels / KR / merge kernels all stage through shared memory):" >> $OUT
echo "  python -m pytest tests/test_gpu_golden.py tests/test_gpu_edge.py tests/test_gpu_kernels.py tests/test_hickry.py -m gpu -x -q" >> $OUT
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 \
    python -m pytest tests/test_gpu_golden.py tests/test_gpu_edge.py tests/test_gpu_kernels.py tests/test_hickry.py -m gpu -x -q > gpurun_out/r2l_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|COMPUTE-SANITIZER$" gpurun_out/r2l_racecheck.log | tail -4 >> $OUT
grep -B2 -A10 "hazard detected\|Race reported" gpurun_out/r2l_racecheck.log | head -60 >> $OUT
cat $OUT
