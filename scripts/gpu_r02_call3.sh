#!/bin/bash
# round 2, third GPU visit: conv_i2c.cu (dense rows, TMA im2col, CTA-pair MMA, TMA-staged epilogue) for the first time.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== i2c parity (bounded)"
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "test_bf16_packed_vs_oracle and (linear or pdl or no_conv_split or default)" 2>&1 | tail -6
echo "== A/B go-9 20x256"
for b in 4096 1024 256 64; do
  for v in "1 1 1" "0 1 0" "0 1 1"; do
    set -- $v
    echo -n "b=$b no_i2c=$1 pair=$2 pdl=$3 "
    KZB_NO_I2C=$1 KZB_CONV_PAIR=$2 KZB_PDL=$3 timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 10 2>&1 | tail -1 | cut -c1-260
  done
done | tee gpurun_out/r02_go9_i2c_ab.txt
for v in "0 0" "0 1"; do
  set -- $v
  echo -n "go19 b2048 no_i2c=$1 pdl=$2 "
  KZB_NO_I2C=$1 KZB_PDL=$2 timeout 200 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch 2048 --iters 3 2>&1 | tail -1 | cut -c1-260
done | tee gpurun_out/r02_go19_i2c.txt
echo "== ncu full: go-9 layers on conv_i2c"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_i2c -s 12 -c 2 -f -o gpurun_out/r02_go9_i2c \
   python scripts/quick_profile.py --game go-9 --depth 3 --channels 256 --batch 2048 --iters 1 > gpurun_out/r02_ncu_i2c.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_ncu_i2c.log | cut -c1-200
echo "== parity at size"
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -q --timeout 600 > gpurun_out/r02_parity_at_size.txt 2>&1; grep -E "^\[|passed|failed|^E " gpurun_out/r02_parity_at_size.txt | cut -c1-400
