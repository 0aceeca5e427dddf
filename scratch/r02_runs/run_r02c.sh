#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02c.log 2>&1; tail -5 $O/pytest_gpu_r02c.log
scratch/ab_r02c.sh dbce-video-cpp_b200/libdbde_b200.so > $O/ab_r02c.log 2>&1
{
for lib in scratch/libdbde_minb3.so scratch/libdbde_cs.so; do
  echo "=== $lib"
  for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003"; do echo "--- direct $cfg"; DBDE_B200_ODD_DECODE=direct DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -1; done
done
echo "=== scratch/libdbde_copy8.so"
for cfg in "1000 10 micro 2048 2048" "1000 10 mix 1001 1003" "500 10 noise 2048 2048"; do echo "--- $cfg"; CHECK=0 DBDE_B200_LIB=scratch/libdbde_copy8.so python scratch/enc_only.py $cfg 2>&1 | tail -1; done
} >> $O/ab_r02c.log 2>&1
M=smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for path in direct staged; do
  DBDE_B200_ODD_DECODE=$path ncu --metrics $M --clock-control none -k regex:dbde_ --csv --log-file $O/ncu_counts_mix1001_$path.csv python scratch/enc_only.py 1000 1 mix 1001 1003 > $O/ncu_counts_$path.log 2>&1
done
ncu --metrics $M --clock-control none -k regex:dbde_ --csv --log-file $O/ncu_counts_micro2048.csv python scratch/enc_only.py 1000 1 micro 2048 2048 > $O/ncu_counts_micro.log 2>&1
cat $O/ab_r02c.log
