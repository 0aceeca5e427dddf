#!/bin/bash
# builds scratch/dropin_latency.cpp against the B200 library (here or on the GPU box) and runs it;
# the reference build (needs oracle/_ref/dbde_util.o, made where /root/reference exists) runs beside it
cd "$(dirname "$0")/.."
g++ -O2 -std=c++14 -DWITH_B200 -Iinclude scratch/dropin_latency.cpp -Ldbce-video-cpp_b200 -ldbde_b200 -Wl,-rpath,$PWD/dbce-video-cpp_b200 -o scratch/dropin_latency_b200 || exit 1
[ -f oracle/_ref/dbde_util.o ] && g++ -O2 -std=c++14 -Iinclude scratch/dropin_latency.cpp oracle/_ref/dbde_util.o -o scratch/dropin_latency_ref
g++ -O2 -std=c++14 -Iinclude scratch/walker_latency.cpp -Ldbce-video-cpp_b200 -ldbde_b200 -Wl,-rpath,$PWD/dbce-video-cpp_b200 -o scratch/walker_latency_b200 || exit 1
[ -f oracle/_ref/dbde_util.o ] && g++ -O2 -std=c++14 -Iinclude scratch/walker_latency.cpp oracle/_ref/dbde_util.o -o scratch/walker_latency_ref
for cfg in "2048 2048 1024 16" "2048 2048 1024 2" "1001 1003 4096 16"; do
  echo "--- walker $cfg"; echo -n "b200: "; scratch/walker_latency_b200 $cfg
  [ -x scratch/walker_latency_ref ] && { echo -n "ref : "; scratch/walker_latency_ref $cfg; }
done
for cfg in "2048 2048 50 0" "2048 2048 50 1" "1001 1003 100 0" "4096 4096 20 0" "512 512 200 0"; do
  echo "--- $cfg"; echo -n "b200: "; scratch/dropin_latency_b200 $cfg
  echo -n "b200 (buffers registered): "; REGISTER=1 scratch/dropin_latency_b200 $cfg
  [ -x scratch/dropin_latency_ref ] && { echo -n "ref : "; scratch/dropin_latency_ref $cfg; }
done
