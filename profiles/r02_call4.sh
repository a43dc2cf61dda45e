#!/bin/bash
# Round 2, GPU call 4: first run of the persistent layer-tail kernel (parity, micro-benchmark, step times, bench)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_layer_tail.py -m gpu -x -q > $O/c4_tail_tests.log 2>&1; rc=$?; echo "tail tests rc=$rc" >> $O/c4_tail_tests.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "qknorm" > $O/c4_qknorm_tests.log 2>&1
if [ $rc -eq 0 ]; then
  timeout 300 python profiles/tail_microbench.py 8,8 > $O/c4_tail_microbench.json 2> $O/c4_tail_microbench.err
  timeout 300 python profiles/tail_microbench.py 4,4 > $O/c4_tail_microbench_s4.json 2> $O/c4_tail_microbench_s4.err
  B200_TAIL=mega timeout 300 python profiles/step_time.py > $O/c4_step_mega.json 2> $O/c4_step_mega.err
  B200_TAIL=mega B200_FUSED_DECODE_MAX=256 timeout 300 python profiles/step_time.py > $O/c4_step_mega_fd256.json 2> $O/c4_step_mega_fd256.err
  B200_TAIL=mega timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/c4_bench_mega.json 2> $O/c4_bench_mega.err
fi
timeout 300 python profiles/step_time.py > $O/c4_step_default.json 2> $O/c4_step_default.err
tail -4 $O/c4_tail_tests.log; cat $O/c4_step_*.json; tail -2 $O/c4_tail_microbench.err
