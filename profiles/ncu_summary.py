#!/usr/bin/env python
"""Summarise ncu output brought back from the GPU box into the small text files committed under profiles/.

  python profiles/ncu_summary.py launches gpurun_out/X_launches.csv          > profiles/rNN_launches.txt
  python profiles/ncu_summary.py kernel   gpurun_out/X.ncu-rep [launch_idx]  > profiles/rNN_<kernel>_ncu.txt
  python profiles/ncu_summary.py hotspots gpurun_out/X.ncu-rep [launch_idx]  > profiles/rNN_<kernel>_source_hotspots.txt

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count, total, share).
`hotspots` walks the SASS program of one profiled launch in segments of 100 instructions and prints where the warp-stall
samples sit (segment share + its three hottest instructions), then the hottest instruction with what precedes it.
`kernel` prints the metrics of one profiled launch of an `ncu --set full` report that the roofline discussion in
DESIGN.md uses: duration, DRAM bytes, pipe utilisation, issue/stall breakdown, occupancy limits, and the
per-opcode warp-stall samples from the source page (needs -lineinfo / --import-source on).
"""
import csv
import re
import subprocess
import sys
from collections import Counter, defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = defaultdict(lambda: [0, 0.0, "", ""])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("s2s::<unnamed>::", "").replace("s2s::", "")
        a = agg[name]
        a[0] += 1; a[1] += v; a[2] = r[gi]; a[3] = r[bi]
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time "
          f"(ncu serialises launches and flushes caches: compare SHARES, not absolutes)")
    print(f"{'kernel':40s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s}  last grid/block")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {n:8d} {t / 1e3:10.1f} {100 * t / tot:6.1f}% {t / n / 1e3:9.1f}  {g} {b}")


KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_bytes\.sum|sm__cycles_elapsed\.max|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__inst_executed_pipe_(xu|alu|fma|lsu|tmem|tc|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(tensor|tc|shared|fma|alu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__warps_eligible\.avg\.per_cycle_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
    r"smsp__sass_inst_executed_op_tmem_(ldt|stt)\.sum|launch__(grid_size|block_size|registers_per_thread|"
    r"shared_mem_per_block|occupancy_limit_\w+|waves_per_multiprocessor)|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum)$")


def kernel(path, idx=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    r = body[idx]
    ni = hdr.index("Kernel Name")
    print(f"# {path}: launch {idx} of {len(body)} profiled; kernel {r[ni]}")
    for h, u, v in zip(hdr, units, r):
        if KEEP.match(h):
            print(f"{h:90s} {v:>16s} {u}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, x in enumerate(rows) if x and x[0] == "Kernel Name"]
    if not starts:
        return
    s = starts[min(idx, len(starts) - 1)]
    e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
    h2 = rows[s + 1]
    si, ei = h2.index("Warp Stall Sampling (All Samples)"), h2.index("Instructions Executed")
    samp, execd = Counter(), Counter()
    for x in rows[s + 2:e]:
        toks = x[1].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        samp[op] += int(x[si]); execd[op] += int(x[ei])
    tot = sum(samp.values()) or 1
    print(f"\n# warp-stall samples by SASS opcode ({tot} samples, {e - s - 2} instructions)")
    for op, n in samp.most_common(16):
        print(f"{op:28s} {100 * n / tot:5.1f}%  warp-instructions executed {execd[op]}")
    mn = [op for op in execd if re.match(r"UTC\w*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|MUFU", op)]
    print("# Blackwell-native opcodes present:", ", ".join(f"{op} x{execd[op]}" for op in sorted(mn)))


def hotspots(path, idx=0, seg=100, floor=0.015):
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    starts = [i for i, x in enumerate(rows) if x and x[0] == "Kernel Name"]
    s = starts[min(idx, len(starts) - 1)]
    e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
    hdr, body = rows[s + 1], [x for x in rows[s + 2:e] if len(x) > 4]
    si, ci = hdr.index("# Samples"), hdr.index("Source")
    n = [int(x[si]) for x in body]
    tot = sum(n) or 1
    print(f"# {path}, launch {idx} ({rows[s][1][:60]}...): warp-stall samples along the SASS program")
    print(f"# {len(body)} instructions, {tot} samples; segments of {seg} instructions with >= {100 * floor:.1f} % of the "
          f"samples, top three instructions of each")
    for a in range(0, len(body), seg):
        share = sum(n[a:a + seg]) / tot
        if share < floor:
            continue
        top = sorted(range(a, min(a + seg, len(body))), key=lambda i: -n[i])[:3]
        print(f"[{a:5d}-{a + seg:5d}] {100 * share:5.1f}%  " +
              " | ".join(f"{body[i][ci].strip()[:44]} ({100 * n[i] / tot:.1f}%)" for i in top))
    spins = ("BRA", "SYNCS", "BAR", "NOP", "ISETP", "BSYNC", "FENCE")      # barrier / mbarrier wait loops
    cand = [i for i in range(len(body)) if not any(t in body[i][ci] for t in spins)]
    hot = max(cand, key=lambda i: n[i])
    print(f"\n# the hottest non-wait instruction ({100 * n[hot] / tot:.1f} % of the samples) and what precedes it")
    for i in range(max(hot - 12, 0), min(hot + 3, len(body))):
        print(f"{i:5d} {n[i]:>6d}  {body[i][ci].strip()[:80]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "hotspots":
        hotspots(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
