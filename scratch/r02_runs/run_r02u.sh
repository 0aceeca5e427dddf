#!/bin/bash
O=gpurun_out
timeout 120 python scratch/soak.py 60 > $O/soak_r02u.log 2>&1; tail -2 $O/soak_r02u.log
