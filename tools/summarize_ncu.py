#!/usr/bin/env python
"""Turn the ncu captures of tools/collect_ncu.sh (gpurun_out/r02_*.csv, *.ncu-rep) into the committed
summaries under profiles/ (run in the build container: ncu reads .ncu-rep files without a GPU).

  profiles/r02_ncu_launch_summary.txt   kernels of ONE training step: launches, total device time, share
  profiles/r02_ncu_full_<name>.txt      per captured launch: duration, DRAM read/write bytes and GB/s,
                                        tensor-pipe %, registers, shared memory, achieved occupancy
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
PEAK_TF, PEAK_GBS = 1393.7, 6489.9   # MEASURED_PEAKS.json: sustained bf16 TFLOP/s, HBM GB/s


def short(name):
    name = name.replace("void ", "").replace("b200::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([A-Za-z_0-9:]+)(<[^()]*>)?\(", name + "(")
    if not m:
        return name[:60]
    base = m.group(1).split("::")[-1]
    targs = m.group(2) or ""
    return base + (targs if len(targs) <= 24 else "<..>")


def launch_summary():
    path = os.path.join(OUT, "r02_launches.csv")
    if not os.path.exists(path):
        return
    rows = []
    txt = open(path).read()
    start = txt.find('"ID"')
    rd = csv.DictReader(io.StringIO(txt[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((short(r["Kernel Name"]), us))
    # one step = from one chain-kernel quadruple to the next: find the launches between the 1st and the 5th rdb_chain_kernel
    idx = [i for i, (n, _) in enumerate(rows) if n == "rdb_chain_kernel"]
    note = ""
    if len(idx) >= 5:
        # a step holds 4 chain launches (2 fwd groups, 2 bwd groups); step boundaries are invisible in the list, so take
        # the window between equal phases of two consecutive steps
        rows_step = rows[idx[0]:idx[4]]
        note = "window: launch %d .. %d of the capture (between the first rdb_chain_kernel of two consecutive steps)" % (idx[0], idx[4])
    else:
        rows_step = rows
        note = "whole capture (fewer than 5 rdb_chain_kernel launches seen)"
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows_step:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    ours = {"rdb_chain_kernel", "conv_igemm256_kernel", "conv_igemm_kernel", "wgrad_rdb_kernel", "conv_wgrad_kernel"}
    with open(os.path.join(PROF, "r02_ncu_launch_summary.txt"), "w") as f:
        f.write("ncu launch list of ONE training step, round-2 build (bench.py, B200_GRAPHS=0; gpurun_out/r02_launches.csv),\n"
                "gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES.  %s\n"
                "%d launches (library kernels AND PyTorch's: fills, Adam, BCE, cuBLAS classifier), %.2f ms total\n\n" %
                (note, len(rows_step), tot / 1e3))
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-58s n=%4d %10.1f us %5.1f%%\n" % (n[:58], c, us, 100.0 * us / tot))
    print("launch summary: %d launches, %.2f ms" % (len(rows_step), tot / 1e3))


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_dim_x", "cluster"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__cycles_active.avg", "cycles_active"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
])


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def full_summary(name):
    rep = os.path.join(OUT, "r02_full_%s.ncu-rep" % name)
    if not os.path.exists(rep):
        return
    try:
        txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    except Exception as e:
        print("ncu failed on", rep, e)
        return
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["ncu --set full --clock-control none --cache-control none (warm caches), round-2 build; one row per captured launch "
             "(gpurun_out/r02_full_%s.ncu-rep).  Peaks: %.1f TFLOP/s sustained bf16, %.1f GB/s HBM (MEASURED_PEAKS.json)." %
             (name, PEAK_TF, PEAK_GBS), ""]
    tensor_cols = [h for h in hdr if "pipe_tensor" in h and "pct_of_peak" in h]
    for r in rows:
        kn = short(r[col["Kernel Name"]])
        vals = OrderedDict()
        for m, label in WANT.items():
            if m in col:
                vals[label] = (to_float(r[col[m]]), units[col[m]])
        dur = vals.get("duration", (None, ""))
        dur_us = None
        if dur[0] is not None:
            dur_us = dur[0] / 1e3 if dur[1] in ("ns", "nsecond") else (dur[0] if dur[1] in ("us", "usecond") else dur[0] * 1e3)

        def bytes_of(label):
            v = vals.get(label)
            if not v or v[0] is None:
                return None
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(v[1], 1)
            return v[0] * mul
        rd_b, wr_b = bytes_of("dram_read"), bytes_of("dram_write")
        line = "%-40s grid %s x %s" % (kn[:40], r[col.get("Grid Size", 0)] if "Grid Size" in col else "?",
                                       r[col.get("Block Size", 0)] if "Block Size" in col else "?")
        if dur_us:
            line += "  %9.1f us" % dur_us
        if rd_b is not None and wr_b is not None and dur_us:
            line += "  DRAM read %.1f MB write %.1f MB = %.0f GB/s (%.1f%% of HBM peak)" % (
                rd_b / 1e6, wr_b / 1e6, (rd_b + wr_b) / dur_us / 1e3, 100.0 * (rd_b + wr_b) / dur_us / 1e3 / PEAK_GBS)
        for label in ("tensor_pipe_pct", "hmma_pct", "dram_pct", "l2_hit_pct", "occupancy_pct", "regs", "dyn_smem", "cluster"):
            v = vals.get(label)
            if v and v[0] is not None:
                line += "  %s %.4g" % (label, v[0])
        extra = [("%s=%s" % (h.split(".")[0].replace("sm__", ""), r[col[h]])) for h in tensor_cols[:4]]
        lines.append(line)
        if extra:
            lines.append("      tensor-pipe counters: " + "  ".join(extra))
    with open(os.path.join(PROF, "r02_ncu_full_%s.txt" % name), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    launch_summary()
    for n in ("rdb_chain", "wgrad_rdb", "conv_igemm256", "conv_wgrad", "bn", "thin", "l1"):
        full_summary(n)
