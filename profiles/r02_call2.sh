#!/bin/bash
# Round 2, GPU call 2: the whole GPU suite once more without -x, the fixed linear micro-benchmark, per-batch step times of
# the library flavours, BASELINE config 3 end to end.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/c2_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> $O/c2_gpu_tests.log
timeout 900 python profiles/linear_microbench.py > $O/c2_linear_microbench.json 2> $O/c2_linear_microbench.err
timeout 300 python profiles/step_time.py > $O/c2_step_default.json 2> $O/c2_step_default.err
B200_LINEAR=tc timeout 300 python profiles/step_time.py > $O/c2_step_tc.json 2> $O/c2_step_tc.err
PDL_LIB=$PWD/nano-vllm_b200/lib/libb200attn_pdl.so
B200ATTN_LIB=$PDL_LIB B200_LINEAR=tc B200_LINEAR_CFG=32,32,64,8,64,8,1 timeout 300 python profiles/step_time.py > $O/c2_step_pdl_tc.json 2> $O/c2_step_pdl_tc.err
B200_LM_HEAD=fused timeout 300 python profiles/step_time.py > $O/c2_step_fused_head.json 2> $O/c2_step_fused_head.err
timeout 900 python profiles/run_config.py 3 $O/c2_config3.json > $O/c2_config3.out 2> $O/c2_config3.err
tail -3 $O/c2_gpu_tests.log; cat $O/c2_step_*.json; tail -c 600 $O/c2_config3.out
