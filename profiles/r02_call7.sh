#!/bin/bash
# Round 2, GPU call 7 (one B200): the suite on the round's final tree with the opt-in presets enabled, smoke(), the headline line,
# and the last policy A/B: o_proj / down_proj on tcgen05 split-K 4 also above 128 rows (B200_LINEAR_MAX_ROWS=256).
#   gpurun --timeout 720 -- 'bash profiles/r02_call7.sh'
mkdir -p gpurun_out
O=gpurun_out
B200_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -rs > $O/c7_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> $O/c7_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > $O/c7_smoke.log 2>&1; echo "smoke rc=$?" >> $O/c7_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/c7_bench_default.json 2> $O/c7_bench_default.err
timeout 200 python profiles/step_time.py 256,192 16 > $O/c7_step_default.json 2> $O/c7_step_default.err
B200_LINEAR_MAX_ROWS=256 timeout 200 python profiles/step_time.py 256,192 16 > $O/c7_step_rows256.json 2> $O/c7_step_rows256.err
tail -4 $O/c7_gpu_tests.log; grep -E "FAILED|ERROR|Timeout" $O/c7_gpu_tests.log | head; grep SKIPPED $O/c7_gpu_tests.log | cut -c1-160 | head -12; tail -2 $O/c7_smoke.log
cat $O/c7_step_default.json $O/c7_step_rows256.json | cut -c1-400
python - <<'PY'
import json
for n in ('default',):
    try:
        d = json.loads([l for l in open(f'gpurun_out/c7_bench_{n}.json').read().splitlines() if l.startswith('{')][-1])
        print(n, round(d['value'], 1), round(d['e2e']['value'], 1), d.get('parity', {}).get('ok'), (d.get('roofline') or {}).get('frac'), d.get('clocks'))
    except Exception as e:
        print(n, 'no json', e)
PY
