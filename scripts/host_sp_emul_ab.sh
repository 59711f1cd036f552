#!/bin/bash
# Host-only emulation of a 4-GPU host with 4 cores per replica on a 16-core box: 4 pinned replicas of the whole self-play
# driver with an emulated GPU each (scripts/host_sp_emul.py), under different generator wake-up policies.
mkdir -p gpurun_out
out=gpurun_out/host_sp_emul_ab.txt
: > $out
run() {  # label, replicas, env...
  label=$1; reps=$2; shift 2
  rm -f /tmp/emul_*.log
  for ((i = 0; i < reps; i++)); do env "$@" SEED=$((i + 1)) taskset -c $((4 * i))-$((4 * i + 3)) python scripts/host_sp_emul.py 4 3 0 5 > /tmp/emul_$i.log 2>&1 & done
  wait
  cat /tmp/emul_*.log | awk -v l="$label" -v r=$reps '{n += $2; e += $4; b += $6; c++} END {printf "%-28s replicas %d  nodes/s per replica %.0f  nn/s per replica %.0f  mean batch %.0f\n", l, r, n / c, e / c, b / c}' | tee -a $out
}
run "alone" 1 X=1
run "alone spin100" 1 KZB_SP_SPIN_US=100
run "4 replicas" 4 X=1
run "4 replicas spin50" 4 KZB_SP_SPIN_US=50
run "4 replicas spin200" 4 KZB_SP_SPIN_US=200
run "4 replicas pinned gens" 4 KZB_SP_PIN_GENERATORS=1
run "4 replicas pinned+spin100" 4 KZB_SP_PIN_GENERATORS=1 KZB_SP_SPIN_US=100
run "4 replicas latency 400" 4 KZB_SP_DUMMY_LATENCY_US=400
