#!/bin/bash
# quick A/B on the headline config + two more: scratch/ab2.sh libA.so libB.so ...
for lib in "$@"; do
  echo "=== $lib"
  for cfg in "1000 5 micro 2048 2048" "300 5 low 4096 4096" "300 5 micro 4096 4096" "500 5 noise 2048 2048"; do
    echo "--- $cfg"
    DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -2
  done
done
