#!/bin/bash
# round 2, 8 GPUs: bench --gpus 8 with the ens block (K = 8 members, the four architectures cycled)
mkdir -p gpurun_out
echo "== bench 4 gpus"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
ln=[l for l in open('gpurun_out/bench_n8.log').read().splitlines() if l.startswith('{')]
if ln:
    d=json.loads(ln[-1]); print({k:d[k] for k in ('value','n_gpus','gpu_launches')}, d['roofline']['frac'], d['parity']['bit_identical']); print(json.dumps(d.get('ens'))[:2500])
else:
    print(open('gpurun_out/bench_n8.log').read()[-3000:])
PY
