#!/bin/bash
# Round 2, 8-GPU call: bench at N=8 with the peer-memory exchange and with the in-switch (NVLS) reduction, then BASELINE
# config 5 (Qwen3-32B dims, TP8, 128 x 8192-token prompts, 1024 output tokens).
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 bench.py --gpus 8 --steps 2 --warmup 2 > $O/tp8_bench_peer.json 2> $O/tp8_bench_peer.err
B200_TP_ALLREDUCE=nvls timeout 600 $TR --master-port 29612 bench.py --gpus 8 --steps 2 --warmup 2 > $O/tp8_bench_nvls.json 2> $O/tp8_bench_nvls.err
timeout 1200 $TR --master-port 29613 profiles/run_config.py 5 $O/tp8_config5.json > $O/tp8_config5.out 2> $O/tp8_config5.err
for f in peer nvls; do python - $O/tp8_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['e2e']['value']), d.get('parity'), d['notes'].get('host_loop'))
except Exception as e: print(sys.argv[1], 'no json', e)
PY
done
tail -c 1500 $O/tp8_config5.out; tail -5 $O/tp8_config5.err
