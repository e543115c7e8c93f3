#!/bin/bash
# round 2, 2 GPUs: bench --gpus 2 after moving the clock sampler's start-up out of the timed region (rank 0 only)
mkdir -p gpurun_out
for i in 1 2; do
echo "== bench 2 gpus, run $i"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n2_$i.log 2>&1; echo "rc=$?"; python - <<PY
import json
ln=[l for l in open('gpurun_out/bench_n2_$i.log').read().splitlines() if l.startswith('{')]
if ln:
    d=json.loads(ln[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['parity']['bit_identical'])
else:
    print(open('gpurun_out/bench_n2_$i.log').read()[-2000:])
PY
done
