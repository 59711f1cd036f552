#!/bin/bash
# round 2, seventh 1-GPU visit: conv_i2c's persistent mode (whole go tower in one cooperative launch) for the first time.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== parity (bounded: a grid barrier that never opens must not hang the box)"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 100 -k "test_bf16_packed_vs_oracle and (linear or no_persist or default) and (go or 256)" 2>&1 | tail -5
rc=$?
echo "parity rc=$rc"
echo "== go-9 20x256: persistent vs per-layer, whole evaluation ms (median of 20, L2 flushed)"
for b in 16 64 256 1024 2048 4096; do
  for persist in 0 100000; do
    echo -n "b=$b KZB_I2C_PERSIST=$persist "
    KZB_I2C_PERSIST=$persist timeout 120 python scripts/quick_profile.py --game go-9 --depth 20 --channels 256 --batch $b --iters 20 2>&1 | tail -1 | cut -c1-250
  done
done | tee gpurun_out/r02_go9_persist_ab.txt
echo "== go-19 40x256 b64 / b512"
for b in 64 512; do
  for persist in 0 100000; do
    echo -n "b=$b KZB_I2C_PERSIST=$persist "
    KZB_I2C_PERSIST=$persist timeout 200 python scripts/quick_profile.py --game go-19 --depth 40 --channels 256 --batch $b --iters 5 2>&1 | tail -1 | cut -c1-250
  done
done | tee gpurun_out/r02_go19_persist_ab.txt
echo "== parity at size + graphs + whole suite"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
