#!/bin/bash
# Round 2, last GPU call: bench at N=2 on the final tree (peer-memory exchange kernel with the all-at-once announcement).
#   gpurun --gpus 2 --timeout 240 -- 'bash profiles/r02_tp2b.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29931 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/tp2b_default.json 2> gpurun_out/tp2b_default.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/tp2b_default.json').read().splitlines() if l.startswith('{')][-1])
    print('tp2', round(d['value']), round(d['e2e']['value']), d.get('parity'), round(d['notes']['host_loop']['ms_per_step'], 3))
except Exception as e:
    print('no json', e)
PY
tail -3 gpurun_out/tp2b_default.err
