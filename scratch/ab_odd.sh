#!/bin/bash
# odd-size workloads only: scratch/ab_odd.sh libA.so libB.so ...
for lib in "$@"; do
  echo "=== $lib"
  for cfg in "1000 5 mix 1001 1003" "1000 5 micro 1001 1003" "1000 5 noise 1001 1003" "1000 5 low 1001 1003"; do
    echo "--- $cfg"
    DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -3
  done
done
