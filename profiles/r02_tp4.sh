#!/bin/bash
# Round 2, 4-GPU call: bench at N=4 (NVLS exchange is the default from 4 ranks; then the peer-memory kernel for comparison if time
# allows) and BASELINE config 4 (Qwen3-8B dims, TP4, 1 024 requests in/out U[100,1024], at most 512 running).
#   gpurun --gpus 4 --timeout 900 -- 'bash profiles/r02_tp4.sh'
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29621 profiles/run_config.py 4 $O/tp4_config4.json > $O/tp4_config4.out 2> $O/tp4_config4.err
timeout 300 $TR --master-port 29622 bench.py --gpus 4 --steps 2 --warmup 2 > $O/tp4_bench_default.json 2> $O/tp4_bench_default.err
python - <<'PY'
import json
for f in ('gpurun_out/tp4_bench_default.json',):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f, round(d['value']), round(d['e2e']['value']), d.get('parity'), d['notes'].get('host_loop'))
    except Exception as e:
        print(f, 'no json', e)
PY
tail -c 1800 $O/tp4_config4.out; tail -4 $O/tp4_config4.err; tail -3 $O/tp4_bench_default.err
