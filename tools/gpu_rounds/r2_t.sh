#!/bin/bash
# round 2, call T: the whole GPU suite + smoke on the final tree
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=10 > gpurun_out/pytest_gpu_r2.txt 2>&1; echo "rc=$?"; tail -16 gpurun_out/pytest_gpu_r2.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
