#!/usr/bin/env python
"""Turn the raw ncu outputs a GPU session left in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarise_profiles.py launches gpurun_out/launches_<tag>.csv profiles/<name>.md "<title>"
  python tools/summarise_profiles.py full gpurun_out/prof_<tag>.ncu-rep profiles/<name>.md "<title>" [records]

`launches`: per-kernel share of the LAST resident step of the bench command (a step starts at k1_classify); the
per-launch times are cold-cache and serialised, so only shares are meaningful.
`full`: the metrics the roofline is read from (DRAM bytes, duration, throughput percentages) for every kernel in
the report; also writes profiles/<name>.json with per-launch DRAM traffic so bench.py can quote `roofline.traffic`.
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

OURS_SKIP = ("native::", "at::", "vectorized_elementwise", "elementwise_kernel", "cub::", "thrust::", "nccl")


def short(name):
    import re
    m = re.match(r"^(?:void )?([\w:]+)", name.strip())          # templated kernels: k1_classify<(bool)1, (bool)1>(...) -> k1_classify
    return m.group(1) if m else name.split("(")[0].replace("void ", "").strip()


def launches(src, dst, title):
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
    ours = [r for r in rows if not any(s in r[0] for s in OURS_SKIP)]
    starts = [i for i, r in enumerate(ours) if r[0].startswith("k1_classify")]
    # the bench runs warm-up steps, the timed resident steps, then the e2e steps; the resident steps are identical,
    # take the last one that is followed by another k1_classify (= a complete step)
    assert len(starts) >= 2, "no complete step in the launch list"
    # the e2e steps at the end push host batches (widen_* kernels): take the last complete RESIDENT step
    step = None
    for a, b in reversed(list(zip(starts[:-1], starts[1:]))):
        if not any(r[0].startswith("widen_") for r in ours[a:b]):
            step = ours[a:b]
            break
    assert step is not None, "no resident step in the launch list"
    tot = sum(r[3] for r in step)
    agg = OrderedDict()
    for r in step:
        k = agg.setdefault(r[0], [0, 0.0])
        k[0] += 1; k[1] += r[3]
    with open(dst, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n")
        f.write("One complete step = %d launches, %.1f us summed kernel time.\n\n" % (len(step), tot / 1e3))
        f.write("| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, n, t / 1e3, 100 * t / tot))
        f.write("\nRaw rows of that step (kernel, grid, block, ns):\n\n```\n")
        for r in step:
            f.write("%s,%s,%s,%d\n" % (r[0], r[1].replace(" ", ""), r[2].replace(" ", ""), r[3]))
        f.write("```\n")
    print("wrote", dst, "step launches", len(step), "us", tot / 1e3)


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
           "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}


def full(src, dst, title, records=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = csv.reader(io.StringIO(out))
    hdr = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rd:
        if not r:
            continue
        k = OrderedDict(name=short(r[col["Kernel Name"]]))
        for m in METRICS:
            if m in col:
                try:
                    v = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                k[m] = (v, units[col[m]])
        kernels.append(k)
    js = {}
    with open(dst, "w") as f:
        f.write("# %s\n\n`ncu --set full --clock-control none --import-source on`, one launch per kernel (first warm-up step of the bench command).\n\n" % title)
        for k in kernels:
            f.write("## %s\n\n| metric | unit | value |\n|---|---|---|\n" % k["name"])
            for m in METRICS:
                if m in k:
                    f.write("| %s | %s | %f |\n" % (m, k[m][1], k[m][0]))
            rd_b = k.get("dram__bytes_read.sum"); wr_b = k.get("dram__bytes_write.sum"); du = k.get("gpu__time_duration.sum")
            if rd_b and wr_b and du:
                tb = rd_b[0] * _UNIT.get(rd_b[1], 1.0) + wr_b[0] * _UNIT.get(wr_b[1], 1.0)
                sec = du[0] * _UNIT.get(du[1], 1.0)
                f.write("\nDRAM traffic per launch: %.4f GB in %.1f us -> %.1f GB/s\n\n" % (tb / 1e9, sec * 1e6, tb / sec / 1e9))
                if k["name"] not in js:
                    js[k["name"]] = {"dram_bytes_per_launch": tb, "duration_us": sec * 1e6}
                    if records:
                        js[k["name"]]["dram_bytes_per_record"] = tb / float(records)
    with open(dst.replace(".md", ".json"), "w") as f:
        json.dump({"source": src, "records": records, "kernels": js}, f, indent=1)
    print("wrote", dst, [k["name"] for k in kernels])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else None)
