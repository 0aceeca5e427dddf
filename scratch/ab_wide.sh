#!/bin/bash
for lib in "$@"; do
  echo "=== $lib"
  for cfg in "300 5 micro 4096 4096" "300 5 low 4096 4096" "100 5 noise 4096 4096" "1000 5 micro 2048 2048"; do
    echo "--- $cfg"
    DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -3
  done
done
