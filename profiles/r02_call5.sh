#!/bin/bash
# Round 2, GPU call 5: first run of the warp-specialised prefill kernel (parity vs oracle and vs flash-attn, micro-benchmark
# against the first-generation kernel), then the new default build (PDL flavour, auto linear policy): suite + bench.
mkdir -p gpurun_out
O=gpurun_out
B200_PREFILL=ws timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zz_flash_attn_parity.py -m gpu -q -k "prefill or attention_module" > $O/c5_ws_tests.log 2>&1; rc=$?; echo "ws tests rc=$rc" >> $O/c5_ws_tests.log
for impl in tc ws; do
  B200_PREFILL=$impl timeout 120 python profiles/prefill_microbench.py > $O/c5_prefill_${impl}_packed.json 2> $O/c5_prefill_${impl}_packed.err
  B200_PREFILL=$impl timeout 120 python profiles/prefill_microbench.py --paged > $O/c5_prefill_${impl}_paged.json 2> $O/c5_prefill_${impl}_paged.err
done
cat $O/c5_prefill_*.json
timeout 1500 python -m pytest tests -m gpu -q > $O/c5_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> $O/c5_gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/c5_bench_default.json 2> $O/c5_bench_default.err
B200_PREFILL=ws timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity > $O/c5_bench_ws.json 2> $O/c5_bench_ws.err
timeout 300 python profiles/step_time.py > $O/c5_step_default.json 2> $O/c5_step_default.err
timeout 900 python profiles/run_config.py 3 $O/c5_config3.json > $O/c5_config3.out 2> $O/c5_config3.err
tail -3 $O/c5_ws_tests.log $O/c5_gpu_tests.log; cat $O/c5_step_default.json
