#!/usr/bin/env python
"""Turn the two ncu artefacts of a round into the tables under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            -> markdown launch table (stdout)
  python tools/ncu_summary.py full /tmp/raw.csv [--traffic profiles/ncu_traffic.json] [--images 4096]
        raw.csv = `ncu -i capture.ncu-rep --page raw --csv`; prints the per-kernel markdown table and
        (with --traffic) rewrites the DRAM bytes-per-launch file bench.py reads for roofline.traffic.
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    name = re.sub(r"^void ", "", name)
    name = name.replace("hohk::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def base(name):
    return re.sub(r"<.*$", "", short(name))


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[short(r[k])].append(float(r[v].replace(",", "")) / 1e6)
    total = sum(sum(x) for x in agg.values())
    print("| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|")
    for name, xs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{name}` | {len(xs)} | {sum(xs):.3f} | {max(xs):.3f} | {100 * sum(xs) / total:.1f}% |")
    print(f"\ntotal device time {total:.1f} ms over {sum(len(x) for x in agg.values())} launches")


COLS = [("ms", "gpu__time_duration.sum", 1e-6), ("DRAM rd GB", "dram__bytes_read.sum", None),
        ("DRAM wr GB", "dram__bytes_write.sum", None),
        ("DRAM %peak", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("regs", "launch__registers_per_thread", 1), ("CTA/SM (smem)", "launch__occupancy_limit_shared_mem", 1),
        ("warps/SMSP", "sm__warps_active.avg.per_cycle_active", 0.25),
        ("IPC/SM", "sm__inst_executed.avg.per_cycle_active", 1),
        ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1),
        ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
        ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1),
        ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
        ("branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", 1),
        ("smem conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
        ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
        ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
        ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
        ("fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
        ("lsu wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1)]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * mult.get(unit, 1)


def full(path, traffic_path=None, images=4096):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    k = hdr.index("Kernel Name")
    t = hdr.index("gpu__time_duration.sum")

    def dur_ms(r):
        return float(r[t].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[t]]

    best = {}
    for r in rows[2:]:
        if len(r) <= t or not r[t]:
            continue
        n = short(r[k])
        if n not in best or dur_ms(r) > dur_ms(best[n]):
            best[n] = r
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |\n|" + "---|" * (len(COLS) + 1))
    traffic = {}
    for n, r in sorted(best.items(), key=lambda kv: -dur_ms(kv[1])):
        cells = []
        for label, metric, scale in COLS:
            i = hdr.index(metric)
            if metric.startswith("dram__bytes"):
                cells.append(f"{to_bytes(r[i], units[i]) / 1e9:.2f}")
            elif metric == "gpu__time_duration.sum":
                cells.append(f"{dur_ms(r):.2f}")
            else:
                try:
                    cells.append(f"{float(r[i].replace(',', '')) * scale:.2f}")
                except ValueError:  # empty, or "no data"
                    cells.append("-")
        print(f"| `{n}` | " + " | ".join(cells) + " |")
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        b = base(n)
        if b not in traffic or dur_ms(r) > traffic[b]["ms_under_ncu"]:
            traffic[b] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                          "ms_under_ncu": dur_ms(r), "kernel": n, "images": images}
    if traffic_path:
        # stamp with the kernel source the capture was taken from: bench.py quotes the figures only while it is unchanged
        import hashlib
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        h = hashlib.sha1()
        for f in ("hoh_kernels.cuh", "hoh_api.cu", "hoh_format.cuh"):
            h.update(open(os.path.join(root, "hoh-ans_b200", "csrc", f), "rb").read())
        traffic["_kernel_source_sha1"] = h.hexdigest()
        json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        tp = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
        im = int(sys.argv[sys.argv.index("--images") + 1]) if "--images" in sys.argv else 4096
        full(sys.argv[2], tp, im)
