#!/bin/bash
# Splits the cost of the odd-size path from the cost of depth divergence: scratch/ab_generic.sh libA.so ...
for lib in "$@"; do
  echo "=== $lib"
  for cfg in "1000 5 mix 1001 1003" "1000 5 micro 1001 1003" "1000 5 mix 1024 1024" "1000 5 micro 1024 1024" "500 5 mix 2048 2048" "1000 5 noise 1001 1003" "1000 5 low 1001 1003"; do
    echo "--- $cfg"
    DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -3
  done
done
