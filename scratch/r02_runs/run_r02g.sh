#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02g.log 2>&1; tail -3 $O/pytest_gpu_r02g.log
{
for v in "dbce-video-cpp_b200/libdbde_b200.so staged" "scratch/libdbde_edgeold.so staged"; do
  set -- $v
  echo "=== $1 ($2)"
  for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "1000 10 noise 1001 1003" "1000 10 low 1001 1003"; do echo "--- $cfg"; DBDE_B200_ODD_DECODE=$2 DBDE_B200_LIB=$1 python scratch/enc_only.py $cfg 2>&1 | tail -2; done
done
for lib in scratch/libdbde_head.so dbce-video-cpp_b200/libdbde_b200.so; do
echo "=== $lib wide odd"
for cfg in "300 10 mix 2049 1003" "300 10 micro 4100 1003"; do echo "--- $cfg"; DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -2; done
done
} > $O/ab_r02g.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_r02g_mix1001 python scratch/enc_only.py 1000 1 mix 1001 1003 > $O/ncu_full_r02g.log 2>&1
cat $O/ab_r02g.log
