#!/bin/bash
# Round 2, GPU call 3: per-batch step times of the linear / PDL variants (which one becomes the default)
mkdir -p gpurun_out
O=gpurun_out
PDL_LIB=$PWD/nano-vllm_b200/lib/libb200attn_pdl.so
timeout 300 python profiles/step_time.py > $O/c3_step_default.json 2> $O/c3_step_default.err
B200_LINEAR=rows timeout 300 python profiles/step_time.py > $O/c3_step_rows.json 2> $O/c3_step_rows.err
B200ATTN_LIB=$PDL_LIB B200_LINEAR=rows B200_LINEAR_CFG=64,64,64,8,64,8,1 timeout 300 python profiles/step_time.py > $O/c3_step_rows_pdl.json 2> $O/c3_step_rows_pdl.err
B200ATTN_LIB=$PDL_LIB B200_LINEAR=tc B200_LINEAR_CFG=64,64,64,8,64,8,1 timeout 300 python profiles/step_time.py > $O/c3_step_tc_pdl.json 2> $O/c3_step_tc_pdl.err
B200ATTN_LIB=$PDL_LIB timeout 300 python profiles/step_time.py > $O/c3_step_pdl.json 2> $O/c3_step_pdl.err
B200_LM_HEAD=fused timeout 300 python profiles/step_time.py > $O/c3_step_fused_head.json 2> $O/c3_step_fused_head.err
B200_FUSED_DECODE_MAX=256 timeout 300 python profiles/step_time.py > $O/c3_step_fuseddec256.json 2> $O/c3_step_fuseddec256.err
cat $O/c3_step_*.json
# (appended while queueing) the layer-tail kernel's first run
bash profiles/r02_call4.sh
