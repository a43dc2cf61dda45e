#!/bin/bash
# Round 2, 2-GPU call: TP tests on real NVLink (NCCL + the peer-memory exchange kernel), bench at N=2 for the exchange variants.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q > $O/tp2_tests.log 2>&1; echo "tp tests rc=$?" >> $O/tp2_tests.log
timeout 600 $TR --master-port 29601 bench.py --gpus 2 --steps 2 --warmup 2 > $O/tp2_bench_peer.json 2> $O/tp2_bench_peer.err
B200_TP_ALLREDUCE=nvls timeout 600 $TR --master-port 29602 bench.py --gpus 2 --steps 2 --warmup 2 > $O/tp2_bench_nvls.json 2> $O/tp2_bench_nvls.err
B200_TP_ALLREDUCE=nccl timeout 600 $TR --master-port 29603 bench.py --gpus 2 --steps 2 --warmup 2 --no-parity > $O/tp2_bench_nccl.json 2> $O/tp2_bench_nccl.err
B200_LINEAR=rows timeout 600 $TR --master-port 29604 bench.py --gpus 2 --steps 2 --warmup 2 > $O/tp2_bench_rows.json 2> $O/tp2_bench_rows.err
tail -3 $O/tp2_tests.log; for f in peer nvls nccl rows; do python - $O/tp2_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['e2e']['value']), d.get('parity'), d['notes'].get('host_loop'))
except Exception as e: print(sys.argv[1], 'no json', e)
PY
done
