#!/bin/bash
for st in 1 2 3; do for ctas in 2 3 4; do
  echo "=== stages $st ctas $ctas"
  for cfg in "1000 5 micro 2048 2048" "300 5 low 4096 4096" "300 5 micro 4096 4096" "500 5 noise 2048 2048"; do
    DBDE_B200_DEC_CTAS=$ctas DBDE_B200_LIB=/root/repo/scratch/libdbde_ds$st.so python scratch/enc_only.py $cfg 2>&1 | grep decode | sed "s/^/$cfg : /"
  done
done; done
