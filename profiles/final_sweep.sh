#!/bin/bash
# Final-state sweep of the posterior and E-step kernels (default planner policy): bash profiles/final_sweep.sh
echo "# posterior decode (all-state gamma + states, f32 obs), ~4e8 read-positions per launch"
timeout 600 python profiles/sweep.py 2,3,4 34,100,170,250,340,450,544,700,1000,1500,2000,3000,5000,8000,10000,16000 2>&1 | grep "K="
echo "# Baum-Welch E-step, ~4e8 read-positions per launch"
for K in 2 3 4; do
  for L in 100 170 400 1000 5000 10000; do
    N=$((400000000 / L)); timeout 120 python profiles/run_estep.py $K $N $L 2>&1 | tail -1
  done
done
