#!/usr/bin/env python
"""Turns the raw ncu output a gpurun call brought back (gpurun_out/) into the small text summaries kept under profiles/.

    python tools/summarize_profiles.py r01
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def launches(name, tag):
    path = os.path.join(OUT, name)
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        k = "%s grid=%s block=%s" % (row["Kernel Name"][:70], row["Grid Size"], row["Block Size"])
        agg[k].append(v)
        order.append((k, v))
    tot = sum(sum(v) for v in agg.values())
    base = name.replace(".csv", "")
    base = base[len(tag) + 1:] if base.startswith(tag + "_") else base
    with open(os.path.join(PROF, "%s_%s.txt" % (tag, base)), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache: compare SHARES)\n")
        f.write("# %d launches, %.1f us total\n" % (len(order), tot))
        f.write("%-110s %6s %10s %8s\n" % ("kernel", "n", "avg_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-110s %6d %10.2f %8.4f\n" % (k, len(v), sum(v) / len(v), sum(v) / tot))


def report(rep, tag):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    base = rep.replace(".ncu-rep", "")
    base = base[len(tag) + 1:] if base.startswith(tag + "_") else base
    with open(os.path.join(PROF, "%s_%s.txt" % (tag, base)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; %s\n" % rep)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\n== %s  grid %s block %s\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
            for k in KEYS:
                if k in d:
                    f.write("  %-72s %14s %s\n" % (k, d[k], units[hdr.index(k)]))
        # stall-reason totals per kernel from the source page
        blocks = src.split('"Kernel Name"')
        for b in blocks[1:]:
            rr = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
            if len(rr) < 3:
                continue
            h = rr[1]
            idx = {x: i for i, x in enumerate(h)}
            stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            tot = collections.Counter()
            n = 0
            for r in rr[2:]:
                if len(r) < len(h) or not r[idx["# Samples"]].isdigit():
                    continue
                n += int(r[idx["# Samples"]])
                for s in stalls:
                    tot[s] += int(r[idx[s]])
            f.write("\n-- warp-state samples of %s (%d samples)\n" % (rr[0][1][:80], n))
            for s, v in tot.most_common(8):
                f.write("  %-28s %8d %6.3f\n" % (s, v, v / max(n, 1)))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    for n in sorted(os.listdir(OUT)):
        if "launches" in n and n.endswith(".csv"):
            launches(n, tag)
    for r in sorted(os.listdir(OUT)):
        if r.endswith(".ncu-rep") and (r.startswith(tag + "_") or tag == "r01"):
            report(r, tag)
    for n in ("ubench.txt", "lstm_bench.txt", "bench.json", "bench_pdl.json", "gpu.txt", "host.txt", "prepare_bench.txt", "variants_bench.txt",
              "train_bench.json", "train_bench_one_row.json", "train_bench_ffma2_0.json", "train_bench_rpi.json", "train_ddp_n2.json",
              "smoke.log"):
        p = os.path.join(OUT, n)
        if os.path.exists(p):
            open(os.path.join(PROF, "%s_%s" % (tag, n)), "w").write(open(p).read())
    print(sorted(os.listdir(PROF)))
