#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02h.log 2>&1; tail -3 $O/pytest_gpu_r02h.log
bash scratch/dropin_mt.sh > $O/dropin_mt_r02h.txt 2>&1; cat $O/dropin_mt_r02h.txt
{
for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "300 10 mix 2049 1003" "1000 10 micro 2048 2048" "300 10 low 4096 4096" "400 10 micro 333 777"; do echo "--- $cfg"; python scratch/enc_only.py $cfg 2>&1 | tail -2; done
} > $O/ab_r02h.log 2>&1
cat $O/ab_r02h.log
python bench.py > $O/bench_r02h.json 2> $O/bench_r02h.err; head -c 400 $O/bench_r02h.json
