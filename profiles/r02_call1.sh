#!/bin/bash
# Round 2, GPU call 1: new parity tests (reference GPU forward, NaN poison, TP over gloo on one GPU), then the staged
# tcgen05 linear / fused LM head / PDL flavour measured for the first time.  Every leg has its own timeout.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c1_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> $O/c1_gpu_tests.log
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "tiny-g1 or tiny-g8" > $O/c1_model_g1_g8.log 2>&1
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q > $O/c1_linear_tests.log 2>&1
rc=$?; echo "linear tests rc=$rc" >> $O/c1_linear_tests.log
timeout 600 python bench.py --steps 2 --warmup 2 > $O/c1_bench_default.json 2> $O/c1_bench_default.err
if [ $rc -eq 0 ]; then
  timeout 900 python profiles/linear_microbench.py > $O/c1_linear_microbench.json 2> $O/c1_linear_microbench.err
  B200_LINEAR=tc timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity > $O/c1_bench_tc.json 2> $O/c1_bench_tc.err
  B200_LM_HEAD=fused timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/c1_bench_fused_head.json 2> $O/c1_bench_fused_head.err
fi
PDL_LIB=$PWD/nano-vllm_b200/lib/libb200attn_pdl.so
B200ATTN_LIB=$PDL_LIB timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q > $O/c1_gpu_tests_pdl.log 2>&1; echo "pdl tests rc=$?" >> $O/c1_gpu_tests_pdl.log
B200ATTN_LIB=$PDL_LIB timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity > $O/c1_bench_pdl.json 2> $O/c1_bench_pdl.err
if [ $rc -eq 0 ]; then
  B200ATTN_LIB=$PDL_LIB B200_LINEAR=tc B200_LINEAR_CFG=32,32,64,8,64,8,1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/c1_bench_pdl_tc.json 2> $O/c1_bench_pdl_tc.err
fi
tail -3 $O/c1_gpu_tests.log $O/c1_linear_tests.log $O/c1_gpu_tests_pdl.log
