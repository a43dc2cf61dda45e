#!/bin/bash
# Round 2, second 8-GPU call: bench at N=8 after the exchange kernel's all-at-once announcement (default = in-switch reduction),
# and B200_LINEAR=gu (gate_up + SiluAndMul as one tcgen05 launch) as A/B.
#   gpurun --gpus 8 --timeout 600 -- 'bash profiles/r02_tp8b.sh'
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29811 bench.py --gpus 8 --steps 2 --warmup 2 > $O/tp8b_default.json 2> $O/tp8b_default.err
B200_LINEAR=gu timeout 240 $TR --master-port 29812 bench.py --gpus 8 --steps 2 --warmup 2 > $O/tp8b_gu.json 2> $O/tp8b_gu.err
python - <<'PY'
import json
for n in ('default', 'gu'):
    f = f'gpurun_out/tp8b_{n}.json'
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(n, round(d['value']), round(d['e2e']['value']), d.get('parity', {}).get('ok'), d.get('parity', {}).get('tp_exchange'), round(d['notes']['host_loop']['ms_per_step'], 3))
    except Exception as e:
        print(n, 'no json', e)
PY
tail -2 $O/tp8b_default.err
