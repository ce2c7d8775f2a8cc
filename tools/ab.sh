#!/bin/bash
# A/B two builds of libcplxk.so on the same box: tools/ab.sh <libA.so> <libB.so> [rounds]
# (box-to-box variance of the pool is ~5 %, larger than most single optimisations)
A=$1; B=$2; R=${3:-2}
export DBG_VARIANTS=${DBG_VARIANTS:-default} DBG_NOISE=${DBG_NOISE:-torch} DBG_ROUNDS=${DBG_ROUNDS:-3}
for i in $(seq 1 $R); do
  for L in "$A" "$B"; do
    echo "== $L"
    CPLXK_LIB=$L python tools/dbg_bench.py 2>&1 | grep ms_med
  done
done
