#!/bin/bash
# Builds the four binaries scripts/host_layout_ab.sh compares: the MCTS microbenchmark against four revisions of
# kzero_b200/csrc/selfplay/mcts.hpp (the synthetic game is the current one for all of them, so only the tree differs).
#   soa     388292e  structure of arrays over every created child (start of the host rework)
#   stat    e4f03cb  visited pool + per-child index slice
#   lists   81ef915  visited pool + per-node visited lists
#   blocks  66a24d8  statistics rows in the parent's block (shipped)
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/micro/ab
build() {  # name, mcts.hpp revision, microbenchmark revision
  d=$(mktemp -d)
  mkdir -p $d/kzero_b200/csrc/selfplay $d/scripts/micro
  git show $2:kzero_b200/csrc/selfplay/mcts.hpp > $d/kzero_b200/csrc/selfplay/mcts.hpp
  cp kzero_b200/csrc/selfplay/games.hpp $d/kzero_b200/csrc/selfplay/games.hpp
  git show $3:scripts/micro/mcts_host_bench.cpp > $d/scripts/micro/mcts_host_bench.cpp
  if [ $1 = soa ]; then  # the first microbenchmark revision, adapted to the old tree interface
    sed -i 's/tree->reserve(800 \* 48 + 64, 800 \* 2 + 64);/tree->reserve(800 * 48 + 64);/; s/tree.pool\[size_t(req.node)\].child_count/tree.child_count[size_t(req.node)]/; s/tree.root().child_start/tree.child_start[0]/' $d/scripts/micro/mcts_host_bench.cpp
  fi
  g++ -O3 -std=c++17 -o scripts/micro/ab/mcts_$1 $d/scripts/micro/mcts_host_bench.cpp
  rm -rf $d
}
build soa 388292e e0fbe26
build stat e4f03cb e4f03cb
build lists 81ef915 81ef915
build blocks 66a24d8 66a24d8
ls -la scripts/micro/ab
