#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/ncu_digest.py launches gpurun_out/launches.csv profiles/r01_chess_b1024_launches.md
  python scripts/ncu_digest.py kernel   gpurun_out/tower8.ncu-rep profiles/r01_tower8_ncu.md [--traffic-json profiles/dram_traffic.json --workload chess]
"""
import collections
import csv
import json
import subprocess
import sys

KERNEL_METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__cycles_elapsed.avg.per_second",
    "sm__cycles_active.avg",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.per_second",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__inst_executed.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    kn, mv, grid, block = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn].replace("unnamed>::", "").replace("void ", "")
        name = name.split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[grid], r[block]])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list digest of `{src}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised:\n"
                "compare SHARES with bench.py's step_breakdown_ms, not absolutes)\n\n")
        f.write("| kernel | launches | grid | block | total us | avg us | share |\n|---|---|---|---|---|---|---|\n")
        for k, (n, t, g, b) in agg.items():
            f.write(f"| `{k}` | {n} | {g} | {b} | {t / 1e3:.1f} | {t / n / 1e3:.1f} | {t / total:.3f} |\n")
    print(open(dst).read())


def kernel(src, dst, traffic_json=None, workload=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"ncu --set full digest of `{src}` (--clock-control none; values are per launch, under the profiler)\n\n")
        for r in data:
            f.write(f"kernel: `{r[name_col]}`\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---|" * len(data) + "\n")
        for m in KERNEL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
    print(open(dst).read())
    if traffic_json:
        # per kernel and workload (key `<kernel>/<bench.py config>`, launches of the same kernel averaged): dram__bytes_read.sum + dram__bytes_write.sum per launch;
        # merged into the file, which bench.py reads for `roofline*.traffic`
        import os
        import re

        def val(r, m):
            i = hdr.index(m)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        table = json.load(open(traffic_json)) if os.path.exists(traffic_json) else {}
        table = {k: v for k, v in table.items() if isinstance(v, dict)}
        groups = collections.OrderedDict()
        for r in data:
            short = re.search(r"(\w+_kernel)", r[name_col])
            groups.setdefault(short.group(1) if short else r[name_col], []).append(r)
        for short, rs in groups.items():
            rd = sum(val(r, "dram__bytes_read.sum") for r in rs) / len(rs)
            wr = sum(val(r, "dram__bytes_write.sum") for r in rs) / len(rs)
            table[f"{short}/{workload}"] = {"workload": workload, "source": src, "launches_averaged": len(rs), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr}
        json.dump(table, open(traffic_json, "w"), indent=1)
        print(open(traffic_json).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst)
    else:
        tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
        wl = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else None  # bench.py config name the capture was taken on
        kernel(src, dst, tj, wl)
