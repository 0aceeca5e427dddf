#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02e.log 2>&1; tail -3 $O/pytest_gpu_r02e.log
{
for lib in scratch/libdbde_head.so dbce-video-cpp_b200/libdbde_b200.so scratch/libdbde_spreadold.so; do
  echo "=== $lib"
  for cfg in "300 10 low 4096 4096" "1000 10 micro 2048 2048" "500 10 mix 2048 2048"; do echo "--- $cfg"; DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -1; done
done
for lib in dbce-video-cpp_b200/libdbde_b200.so scratch/libdbde_stgdyn.so scratch/libdbde_spreadold.so; do
  echo "=== $lib (staged)"
  for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "1000 10 noise 1001 1003" "1000 10 low 1001 1003"; do echo "--- $cfg"; DBDE_B200_ODD_DECODE=staged DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -1; done
done
} > $O/ab_r02e.log 2>&1
M=smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
DBDE_B200_ODD_DECODE=staged ncu --metrics $M --clock-control none -k regex:dbde_decode -c 4 --csv --log-file $O/ncu_counts_r02e_staged.csv python scratch/enc_only.py 1000 1 mix 1001 1003 > /dev/null 2>&1
cat $O/ab_r02e.log; grep -v "^==" $O/ncu_counts_r02e_staged.csv | cut -d, -f5,13- | tail -16
