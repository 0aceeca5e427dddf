#!/bin/bash
# one GPU call that refreshes everything under profiles/ for the current code: scratch/profile_round.sh TAG
TAG=${1:-r02n}
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
timeout 300 python bench.py --impl reference > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
timeout 300 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_${TAG}_micro2048 python scratch/enc_only.py 1000 1 micro 2048 2048 > $O/ncu_full_$TAG.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_${TAG}_low4096 python scratch/enc_only.py 300 1 low 4096 4096 >> $O/ncu_full_$TAG.log 2>&1
head -c 600 $O/bench_$TAG.json
