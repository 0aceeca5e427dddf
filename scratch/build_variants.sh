#!/bin/bash
# builds A/B variants of the library into scratch/:  scratch/build_variants.sh name1 "-DFLAG=1 ..." name2 "..." ...
# (each lands in scratch/libdbde_<name>.so; run them with DBDE_B200_LIB=... python scratch/enc_only.py ...)
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -B -C dbce-video-cpp_b200/csrc OUT=$PWD/scratch/libdbde_$name.so VARIANT="$flags" 2>&1 | grep -E "error|nvcc failed" 
  ls -la scratch/libdbde_$name.so | awk '{print $5, $9}'
done
