#!/bin/bash
O=gpurun_out
{
for cfg in "1000 10 micro 1920 1080" "400 10 micro 2560 2160" "1500 10 micro 1280 1024" "400 10 micro 2304 2304" "1500 10 micro 1392 1040" "4000 10 micro 658 494" "4000 10 micro 640 480" "200 10 micro 4096 3000" "1000 10 micro 1936 1216" "300 10 micro 3840 2160"; do echo "--- $cfg"; timeout 120 python scratch/enc_only.py $cfg 2>&1 | tail -3; done
} > $O/camera_sizes_r02r.log 2>&1
cat $O/camera_sizes_r02r.log
