#!/bin/bash
# one GPU call that refreshes everything under profiles/ for the current code: scratch/profile_round.sh TAG
TAG=${1:-r01d}
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
python bench.py --impl reference > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
python bench.py --kind mix --width 1001 --height 1003 --no-e2e --no-cpu-baseline > $O/bench_mix1001_$TAG.json 2>> $O/bench_$TAG.err
python bench.py --kind low --width 4096 --height 4096 --frames 300 --no-e2e --no-cpu-baseline > $O/bench_low4096_$TAG.json 2>> $O/bench_$TAG.err
python bench.py --kind noise --frames 500 --no-e2e --no-cpu-baseline > $O/bench_noise2048_$TAG.json 2>> $O/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_${TAG}_micro2048 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_${TAG}_mix1001 python bench.py --kind mix --width 1001 --height 1003 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline >> $O/ncu_full_$TAG.log 2>&1
cat $O/bench_$TAG.json | head -c 600
