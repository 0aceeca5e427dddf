#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_r02zb_micro2304 python scratch/enc_only.py 400 1 micro 2304 2304 > $O/ncu_full_r02zb.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_r02zb_micro4100 python scratch/enc_only.py 300 1 micro 4100 1003 >> $O/ncu_full_r02zb.log 2>&1
tail -3 $O/ncu_full_r02zb.log
