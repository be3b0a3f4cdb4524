#!/bin/bash
# compute-sanitizer memcheck over a small subset of the kernel tests. Bounded: the tool slows kernels
# down by 10-100x. Output: gpurun_out/sanitizer_memcheck.log
mkdir -p gpurun_out
timeout -k 10 170 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv.py tests/test_gpu_attention.py tests/test_gpu_fp8.py -m gpu -q -x \
  -k "gemm_epilogues or (norm_fwd_bwd and 768) or cross_entropy or (conv_fprop_implicit and shape0) or (conv_wgrad_implicit and shape0-0) or (attention_bwd and case0) or (gemm_fp8 and mnk0)" \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer_memcheck.log | tail -8
timeout -k 10 200 python -m pytest tests -m gpu -q -x -k "gpt2 or lagom or trainstep or residual" 2>&1 | tail -3
timeout -k 10 120 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_last.log 2>&1
grep '^{' gpurun_out/bench_last.log | tail -1 | cut -c1-240
