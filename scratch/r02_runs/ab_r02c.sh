#!/bin/bash
# round-2 A/B: odd-size decode paths (direct vs staged), encoder variants.  scratch/ab_r02c.sh lib...
for lib in "$@"; do
  echo "=== $lib"
  for path in direct staged; do
    for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "1000 10 noise 1001 1003"; do
      echo "--- $path $cfg"
      DBDE_B200_ODD_DECODE=$path DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -3
    done
  done
  for cfg in "1000 10 micro 2048 2048" "500 10 mix 2048 2048" "300 10 low 4096 4096" "500 10 noise 2048 2048" "300 10 mix 2049 1003"; do
    echo "--- $cfg"
    DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -3
  done
done
