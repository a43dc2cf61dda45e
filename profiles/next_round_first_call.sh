#!/bin/bash
# First GPU call of the next round: validate what could not be run this round, in order of risk.
#   gpurun --timeout 2400 -- 'bash profiles/next_round_first_call.sh'
# Every leg has its own timeout so that a hanging staged kernel cannot take the box with it.
mkdir -p gpurun_out
# 1. changes made after the last validated GPU run: golden-logit comparison, peer-exchange timeout / self-test plumbing
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nr_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/nr_gpu_tests.log
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "tiny-g1 or tiny-g8" > gpurun_out/nr_model_g1_g8.log 2>&1
# 2. the staged tcgen05 linear layer: parity first, then the sweep, then the whole model through it
B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q > gpurun_out/nr_linear_tests.log 2>&1
rc=$?; echo "linear tests rc=$rc" >> gpurun_out/nr_linear_tests.log
if [ $rc -eq 0 ]; then
  timeout 900 python profiles/linear_microbench.py > gpurun_out/nr_linear_microbench.json 2> gpurun_out/nr_linear_microbench.err
  B200_LINEAR=tc timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nr_bench_tc.json 2> gpurun_out/nr_bench_tc.err
  B200_LM_HEAD=fused timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nr_bench_fused_head.json 2> gpurun_out/nr_bench_fused_head.err
fi
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nr_bench_default.json 2> gpurun_out/nr_bench_default.err
# 3. the PDL flavour of the library (every kernel: launch_dependents + wait, every launch with the PDL attribute)
PDL_LIB=$PWD/nano-vllm_b200/lib/libb200attn_pdl.so
B200ATTN_LIB=$PDL_LIB timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/nr_gpu_tests_pdl.log 2>&1; echo "pdl tests rc=$?" >> gpurun_out/nr_gpu_tests_pdl.log
B200ATTN_LIB=$PDL_LIB timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nr_bench_pdl.json 2> gpurun_out/nr_bench_pdl.err
B200ATTN_LIB=$PDL_LIB B200_LINEAR=tc B200_LINEAR_CFG=32,32,64,8,64,8,1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nr_bench_pdl_tc.json 2> gpurun_out/nr_bench_pdl_tc.err
# Multi-GPU legs (separate gpurun --gpus N calls; fused exchange at N=8 was never measured, NVLS variant never run):
#   gpurun --gpus 8 --timeout 1500 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/nr_tp8_peer.json; B200_TP_ALLREDUCE=nvls python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/nr_tp8_nvls.json'
