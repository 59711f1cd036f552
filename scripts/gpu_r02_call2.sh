#!/bin/bash
# round 2, second GPU visit: the CTA-pair kernel with direct peer-to-leader TMA signalling; ncu on both go kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pair kernel parity (bounded)"
KZB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "conv_pair" 2>&1 | tail -4
echo "== A/B go-9 20x256"
for b in 4096 1024 256; do
  for v in "0 1" "1 1"; do
    set -- $v
    echo -n "b=$b pair=$1 pdl=$2 "
    KZB_CONV_PAIR=$1 KZB_PDL=$2 timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 | cut -c1-260
  done
done | tee gpurun_out/r02_go9_pair2_ab.txt
echo -n "go19 b2048 pair=1 pdl=1 "
KZB_CONV_PAIR=1 KZB_PDL=1 timeout 200 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 2048 --iters 3 2>&1 | tail -1 | cut -c1-260 | tee gpurun_out/r02_go19_pair2.txt
echo "== ncu full: one go-9 layer, both kernels"
for pair in 0 1; do
  KZB_CONV_PAIR=$pair timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tch -s 12 -c 2 -f -o gpurun_out/r02_go9_tch_pair$pair \
     python scripts/quick_profile.py --game go-9 --depth 3 --channels 256 --batch 2048 --iters 1 > gpurun_out/r02_ncu_pair$pair.log 2>&1
  echo "ncu pair=$pair rc=$?"; tail -2 gpurun_out/r02_ncu_pair$pair.log | cut -c1-200
done
echo "== parity at size"
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -q --timeout 600 > gpurun_out/r02_parity_at_size.txt 2>&1; grep -E "^\[|passed|failed|^E " gpurun_out/r02_parity_at_size.txt | cut -c1-400
