#!/bin/bash
# E-step timing on short reads, bundled kernel on/off: bash profiles/run_estep_short.sh
for sh in 0 1; do
  for cfg in "2 4000000 100" "2 2352941 170" "2 1600000 250" "3 2352941 170" "3 1600000 250"; do
    echo -n "SMF_SHORT=$sh "; SMF_SHORT=$sh timeout 120 python profiles/run_estep.py $cfg 2>&1 | tail -1
  done
done
