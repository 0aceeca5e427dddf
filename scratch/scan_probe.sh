#!/bin/bash
for lib in "$@"; do echo "=== $lib"; DBDE_B200_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dbde_decode_scan -c 4 python scratch/enc_only.py 1000 1 micro 2048 2048 2>&1 | grep -E "gpu__time_duration" ; done
