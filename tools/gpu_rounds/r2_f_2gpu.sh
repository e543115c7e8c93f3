#!/bin/bash
# round 2, 2 GPUs: multi-GPU tests (NCCL sharding, NCCL ensemble incl. random start, fused P2P ensemble), bench --gpus 2 with the ens block
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_multigpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_multigpu.log
echo "== bench 2 gpus"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
ln=[l for l in open('gpurun_out/bench_n2.log').read().splitlines() if l.startswith('{')]
if ln:
    d=json.loads(ln[-1]); print({k:d[k] for k in ('value','n_gpus','gpu_launches')}, d['roofline']['frac'], d['parity']['bit_identical']); print(json.dumps(d.get('ens'))[:1500])
else:
    print(open('gpurun_out/bench_n2.log').read()[-2000:])
PY
echo "== bench reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-200
