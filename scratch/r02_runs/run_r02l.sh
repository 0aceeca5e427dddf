#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02l.log 2>&1; tail -3 $O/pytest_gpu_r02l.log
timeout 600 bash scratch/dropin_mt.sh > $O/dropin_mt_r02l.txt 2>&1; cat $O/dropin_mt_r02l.txt
timeout 300 python bench.py > $O/bench_r02l.json 2> $O/bench_r02l.err; head -c 300 $O/bench_r02l.json
