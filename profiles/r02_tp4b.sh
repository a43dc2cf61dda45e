#!/bin/bash
# Round 2, second 4-GPU call: the exchange kernel now announces to all peers at once (one thread per peer); A/B of two per-layer
# launch savings under tensor parallelism: fused decode front end at every batch size, gate_up + SiluAndMul on tcgen05.
#   gpurun --gpus 4 --timeout 900 -- 'bash profiles/r02_tp4b.sh'
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 240 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 4 --steps 2 --warmup 2 > $O/tp4b_$name.json 2> $O/tp4b_$name.err
}
run default B200_NOOP=1
run fused512 B200_FUSED_DECODE_MAX=512
run gu B200_LINEAR=gu
run gu_fused512 B200_LINEAR=gu B200_FUSED_DECODE_MAX=512
python - <<'PY'
import json
for n in ('default', 'fused512', 'gu', 'gu_fused512'):
    f = f'gpurun_out/tp4b_{n}.json'
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(n, round(d['value']), round(d['e2e']['value']), d.get('parity', {}).get('ok'), d.get('parity', {}).get('tp_exchange'), round(d['notes']['host_loop']['ms_per_step'], 3))
    except Exception as e:
        print(n, 'no json', e)
PY
tail -2 $O/tp4b_default.err
