#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/../.."
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02za.log 2>&1; tail -3 $O/pytest_gpu_r02za.log
timeout 900 bash scratch/dropin_mt.sh > $O/dropin_mt_r02za.txt 2>&1; cat $O/dropin_mt_r02za.txt
for T in 8; do echo -n "b200: "; timeout 90 scratch/dropin_mt_b200 2048 2048 60 0 $T; echo -n "ref : "; timeout 90 scratch/dropin_mt_ref 2048 2048 60 0 $T; done | tee -a $O/dropin_mt_r02za.txt
