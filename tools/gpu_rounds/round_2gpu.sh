#!/bin/bash
mkdir -p gpurun_out
echo "== bench ens K=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/bench_ens.py --batch 64 --steps 5 > gpurun_out/bench_ens2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ens2.log | cut -c1-1200
echo "== ncu full stream kernel"; CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"FusedStreamOp" -s 2 -c 2 -o gpurun_out/prof_fused_stream_r1 python tools/prof_fused.py fused > gpurun_out/ncu_fused_stream.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_fused_stream.log
echo "== other configs (1 GPU)"
for cfg in "ditimi resnet50 32 10" "sim resnet50 16 10" "vmifgsm vit_b_16 32 2" "emifgsm resnet50 16 10" "tim resnet50 64 10"; do set -- $cfg
  CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --attack $1 --arch $2 --batch $3 --epoch $4 --steps 3 --warmup 2 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_cfg_$1.log 2>&1; echo "$1 rc=$?"; tail -1 gpurun_out/bench_cfg_$1.log | cut -c1-220
done
