#!/bin/bash
# A/B of the search tree layouts under memory contention, host only: P pinned copies of the MCTS microbenchmark
# (scripts/micro/mcts_host_bench.cpp built from four revisions into scripts/micro/ab/ by scripts/build_layout_ab.sh), 64 trees each,
# aggregate nodes/s.
mkdir -p gpurun_out
out=gpurun_out/host_layout_ab.txt
lscpu | egrep "Model name|^CPU\(s\)|L2|L3" > $out
P=$(nproc)
for variant in soa stat lists blocks; do
  for procs in 1 $P; do
    rm -f /tmp/ab_*.log
    for ((i = 0; i < procs; i++)); do taskset -c $i scripts/micro/ab/mcts_$variant 64 800 4 > /tmp/ab_$i.log & done
    wait
    cat /tmp/ab_*.log | awk -v v=$variant -v p=$procs '{n += $3; g += $8; a += $11; c++} END {printf "%-7s procs %2d  total %.0f nodes/s  per proc %.0f  gather %.0f  apply %.0f cycles/node\n", v, p, n, n / c, g / c, a / c}' | tee -a $out
  done
done
