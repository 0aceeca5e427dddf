#!/bin/bash
# scratch/dropin_mt.sh [lib]: threaded callers of dbde_pack_frame / dbde_unpack_frame on malloc'd buffers, this library
# beside the reference object on the same box (the reference build needs oracle/_ref/dbde_util.o, made where /root/reference exists)
cd "$(dirname "$0")/.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
[ -f oracle/_ref/dbde_util.o ] && g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp oracle/_ref/dbde_util.o -o scratch/dropin_mt_ref
for mode in 0 1; do
  for T in 1 4 16; do
    echo -n "b200: "; timeout 90 scratch/dropin_mt_b200 2048 2048 $((T > 4 ? 60 : 100)) $mode $T
    [ -x scratch/dropin_mt_ref ] && { echo -n "ref : "; timeout 90 scratch/dropin_mt_ref 2048 2048 $((T > 4 ? 60 : 100)) $mode $T; }
  done
done
for T in 1 16; do
  echo -n "b200: "; timeout 90 scratch/dropin_mt_b200 1001 1003 200 0 $T
  [ -x scratch/dropin_mt_ref ] && { echo -n "ref : "; timeout 90 scratch/dropin_mt_ref 1001 1003 200 0 $T; }
done
