#!/bin/bash
O=gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests -m gpu -q -k "sharded or visible_devices" > $O/pytest_gpu2_r02o.log 2>&1; tail -3 $O/pytest_gpu2_r02o.log
timeout 120 python scratch/soak.py 45 > $O/soak_r02o.log 2>&1; tail -3 $O/soak_r02o.log
DBDE_B200_ODD_DECODE=direct timeout 60 python scratch/soak.py 20 > $O/soak_direct_r02o.log 2>&1; tail -2 $O/soak_direct_r02o.log
timeout 300 python scratch/stream_bench.py 2048 256 micro 4096 4096 > $O/stream_n2_r02o.log 2>&1; tail -3 $O/stream_n2_r02o.log
