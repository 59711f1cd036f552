#!/bin/bash
# smoke + bench (N=1) + ncu launch list + one full ncu capture of the tower kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ -n "$NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  tail -2 gpurun_out/ncu_list.log | cut -c1-300
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tower8 -s 4 -c 1 -f -o gpurun_out/tower8 \
      python scripts/quick_profile.py --iters 1 > gpurun_out/ncu_tower8.log 2>&1
  tail -2 gpurun_out/ncu_tower8.log
fi
