#!/bin/bash
# Round 2, GPU call 6 (first call after the container was re-created): validate HEAD end to end on one B200, take the
# round's ncu evidence, and give the warp-specialised prefill kernel its first run.
#   gpurun --timeout 2100 -- 'bash profiles/r02_call6.sh'
mkdir -p gpurun_out
O=gpurun_out
# 1. the whole GPU suite on the default build (PDL flavour, auto linear policy)
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > $O/c6_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> $O/c6_gpu_tests.log
# 2. the headline line
timeout 600 python bench.py --steps 3 --warmup 3 > $O/c6_bench_default.json 2> $O/c6_bench_default.err
# 3. captured decode step per batch size
timeout 300 python profiles/step_time.py > $O/c6_step_default.json 2> $O/c6_step_default.err
# 4. ncu: launch lists of one decode step (batch 256) and one prefill step, eager, then the dominant kernel with --set full
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $O/c6_launches_decode.csv python profiles/step_trace.py decode 1 > $O/c6_trace_decode.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $O/c6_launches_prefill.csv python profiles/step_trace.py prefill 1 > $O/c6_trace_prefill.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:paged_decode -c 1 \
  -o $O/c6_decode_full -f python profiles/decode_microbench.py --ncu --steps 0 > $O/c6_decode_full.log 2>&1
# 5. warp-specialised prefill kernel: parity (oracle + flash-attn), micro-benchmark against the first-generation kernel
B200_PREFILL=ws timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zz_flash_attn_parity.py -m gpu -q -k "prefill or attention_module" > $O/c6_ws_tests.log 2>&1; rc=$?; echo "ws tests rc=$rc" >> $O/c6_ws_tests.log
for impl in tc ws; do
  B200_PREFILL=$impl timeout 120 python profiles/prefill_microbench.py > $O/c6_prefill_${impl}_packed.json 2> $O/c6_prefill_${impl}_packed.err
  B200_PREFILL=$impl timeout 120 python profiles/prefill_microbench.py --paged > $O/c6_prefill_${impl}_paged.json 2> $O/c6_prefill_${impl}_paged.err
done
# 6. two-stream decode step (B200_DUAL): parity first, then does the chain hide under the other half's attention, then the step
grep -q "two_stream.*FAILED\|FAILED.*two_stream" $O/c6_gpu_tests.log; drc=$((1 - $?))            # 0 when the two-stream tests did not fail
for cfg in "128 3 64,64,64,8,64,8" "128 3 32,64,64,8,64,8" "128 2 64,64,64,8,64,8"; do
  timeout 120 python profiles/dual_microbench.py $cfg >> $O/c6_dual_microbench.jsonl 2>> $O/c6_dual_microbench.err
done
if [ $drc -eq 0 ]; then
  B200_DUAL=1 timeout 300 python profiles/step_time.py 256,224,192,160,128 > $O/c6_step_dual.json 2> $O/c6_step_dual.err
  B200_DUAL=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/c6_bench_dual.json 2> $O/c6_bench_dual.err
fi
tail -3 $O/c6_gpu_tests.log $O/c6_ws_tests.log; grep -E "FAILED|ERROR|Timeout" $O/c6_gpu_tests.log | head -20; cat $O/c6_prefill_*.json; cat $O/c6_step_default.json | head -c 1500
cat $O/c6_dual_microbench.jsonl; head -c 900 $O/c6_step_dual.json; tail -2 $O/c6_dual_microbench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/c6_bench_dual.json').read().strip().splitlines()[-1])
    print('bench dual', round(d['value']), round(d['e2e']['value']), d.get('parity', {}).get('ok'))
except Exception as e:
    print('bench dual: no json', e)
PY
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/c6_bench_default.json').read().strip().splitlines()[-1])
    print('bench', round(d['value']), round(d['e2e']['value']), d.get('parity', {}).get('ok'), d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('bench: no json', e)
PY
