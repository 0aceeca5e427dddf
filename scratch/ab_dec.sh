#!/bin/bash
# A/B of library variants on the odd-size / depth-mix workloads:  scratch/ab_dec.sh libA.so libB.so ...
for lib in "$@"; do
  echo "=== $lib"
  for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "1000 10 noise 1001 1003" "500 10 mix 2048 2048" "1000 10 micro 2048 2048" "300 10 low 4096 4096"; do
    echo "--- $cfg"
    DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -3
  done
done
