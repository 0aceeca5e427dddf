#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02f.log 2>&1; tail -3 $O/pytest_gpu_r02f.log
{
for v in "dbce-video-cpp_b200/libdbde_b200.so staged" "scratch/libdbde_edgeold.so staged" "scratch/libdbde_dynold.so staged" "dbce-video-cpp_b200/libdbde_b200.so direct" "scratch/libdbde_dirpass.so direct"; do
  set -- $v
  echo "=== $1 ($2)"
  for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "1000 10 noise 1001 1003" "1000 10 low 1001 1003"; do echo "--- $cfg"; DBDE_B200_ODD_DECODE=$2 DBDE_B200_LIB=$1 python scratch/enc_only.py $cfg 2>&1 | tail -1; done
done
echo "=== default lib, aligned"
for cfg in "300 10 low 4096 4096" "1000 10 micro 2048 2048"; do echo "--- $cfg"; python scratch/enc_only.py $cfg 2>&1 | tail -2; done
} > $O/ab_r02f.log 2>&1
M=smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
for path in staged direct; do
DBDE_B200_ODD_DECODE=$path ncu --metrics $M --clock-control none -k regex:dbde_decode_ -c 4 --csv --log-file $O/ncu_counts_r02f_$path.csv python scratch/enc_only.py 1000 1 mix 1001 1003 > /dev/null 2>&1
done
cat $O/ab_r02f.log; for path in staged direct; do grep -v "^==" $O/ncu_counts_r02f_$path.csv | cut -d, -f5,13- | tail -7; done
