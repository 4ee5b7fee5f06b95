#!/usr/bin/env python
"""Summarise ncu output into small, tracked text files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_TAG.csv profiles/TAG_launches.md
    python scripts/ncu_summary.py full gpurun_out/prof_KERNEL_TAG.ncu-rep profiles/TAG_KERNEL_full.md

`launches`: per-kernel totals and shares of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised: shares
are meaningful, absolutes are not).  `full`: the metrics of a `--set full` capture that the roofline discussion uses
(DRAM bytes, L2/L1 hit rates, warp execution efficiency, occupancy, registers, stall reasons).
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = min(a[2], v)
        a[3] = max(a[3], v)
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised;\n")
        f.write("use the SHARE column, not the absolutes.\n\n")
        f.write("| kernel | launches | total ms | share | min us | max us |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot:.3f} | {v[2] / 1e3:.1f} | {v[3] / 1e3:.1f} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        f.write("One column per captured launch.  Taken under the profiler: times are for shares and ratios only.\n\n")
        names = [r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"{n} #{i}" for i, n in enumerate(names)) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for m in FULL_METRICS:
            if m not in hdr:
                continue
            i = hdr.index(m)
            vals = []
            for r in rows[2:]:
                try:
                    vals.append(f"{float(r[i].replace(',', '')):.4g}")
                except ValueError:
                    vals.append(r[i])
            f.write(f"| `{m}` | {units[i]} | " + " | ".join(vals) + " |\n")
    print(open(dst).read())


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(src, iters_log, skip, stage, dst):
    """DRAM traffic of the captured launches (dram__bytes_read.sum + dram__bytes_write.sum) next to the number of rays / states
    each launch processed (BTRC_ITER_LOG lines of the same run; launch k of a stage kernel = iteration k).  Appends to a JSON
    file that bench.py reads for `roofline.traffic`."""
    import json
    import os
    import re

    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    iters = {}
    for line in open(iters_log):
        m = re.search(r"\[btrc iter\] (\d+) n_active=(\d+) n_new=(\d+) n_next=(\d+) n_shadow=(\d+) classes=([\d,]+)", line)
        if m:
            cls = [int(x) for x in m.group(6).split(",")]
            # class 3 = 1 + MAT_DISNEY; SHADE_CLASS selects another material's list (e.g. 1 = diffuse for C1 / C3 / C4)
            iters[int(m.group(1))] = {"trace": int(m.group(2)), "shadow": int(m.group(5)), "shade": cls[int(os.environ.get("SHADE_CLASS", "3"))],
                                      "medium": int(m.group(2)), "generate": int(m.group(3))}
    launches_ = []
    for k, r in enumerate(rows[2:]):
        idx = int(skip) + k
        units_of = iters.get(idx, {}).get(stage)
        rd = float(r[ir].replace(",", "")) * UNIT[units[ir]]
        wr = float(r[iw].replace(",", "")) * UNIT[units[iw]]
        launches_.append({"launch": idx, "units": units_of, "dram_read_bytes": rd, "dram_write_bytes": wr,
                          "time_us_under_ncu": float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[it], 1.0)})
    good = [l for l in launches_ if l["units"]]
    entry = {"source": os.path.basename(src), "launches": launches_,
             "dram_bytes_per_unit": (sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in good) / sum(l["units"] for l in good)) if good else None}
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    data[stage] = entry
    json.dump(data, open(dst, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:7])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
