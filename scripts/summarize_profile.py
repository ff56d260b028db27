#!/usr/bin/env python
"""Turn the ncu artefacts a GPU-box visit leaves in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profile.py r1a            # reads gpurun_out/launches.csv + prof_grouped.ncu-rep
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(tag):
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return None
    rows = [l for l in open(path) if not l.startswith("==")]
    d = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(rows))):
        d.setdefault(r["Kernel Name"], []).append(float(r["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in d.values())
    lines = ["# %s: kernel launch list of `bench.py --steps 6 --warmup 3` under ncu" % tag,
             "(`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and",
             "serialised: compare SHARES, not absolutes)", "", "| kernel | launches | mean ns | share of GPU time |", "|---|---|---|---|"]
    for k, v in d.items():
        lines.append("| `%s` | %d | %.0f | %.1f %% |" % (k[:110], len(v), sum(v) / len(v), 100 * sum(v) / total))
    open(os.path.join(OUT, "%s_launches.md" % tag), "w").write("\n".join(lines) + "\n")
    return d


def full(tag, rep="prof_grouped.ncu-rep", workload="cfg1"):
    path = os.path.join(SRC, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    out = {"kernel": data[0][name_i], "launches_captured": len(data), "metrics": {}}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals = [float(r[i].replace(",", "")) for r in data if r[i] not in ("", "n/a")]
            if vals:
                out["metrics"][k] = {"unit": units[i], "mean": sum(vals) / len(vals)}
    m = out["metrics"]

    def to_bytes(key):
        if key not in m:
            return 0.0
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m[key]["unit"]]
        return m[key]["mean"] * scale

    traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    out["dram_bytes_per_launch"] = traffic
    json.dump(out, open(os.path.join(OUT, "%s_full_%s.json" % (tag, workload)), "w"), indent=1)
    json.dump({"dram_bytes_per_launch": traffic, "source": "%s_full_%s.json" % (tag, workload),
               "kernel": out["kernel"]}, open(os.path.join(OUT, "traffic_%s.json" % workload), "w"), indent=1)
    # hottest SASS lines
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        try:
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
            body = []
            for r in rows[2:]:
                if len(r) < len(hdr) or r[0] == "Kernel Name":
                    break
                body.append(r)
            tot = sum(int(r[iS]) for r in body) or 1
            top = sorted(body, key=lambda r: -int(r[iS]))[:25]
            lines = ["# %s: hottest SASS instructions of %s (warp-stall samples)" % (tag, out["kernel"][:80]), "",
                     "| samples | share | executed | instruction |", "|---|---|---|---|"]
            for r in top:
                lines.append("| %s | %.1f %% | %s | `%s` |" % (r[iS], 100 * int(r[iS]) / tot, r[iI], r[1].strip()[:90]))
            ops = collections.Counter()
            for r in body:
                tok = r[1].split()
                op = tok[1] if tok and tok[0].startswith("@") else (tok[0] if tok else "?")
                ops[op.split(".")[0]] += int(r[iI])
            lines += ["", "Executed warp-instructions by opcode: " + ", ".join("%s %d" % kv for kv in ops.most_common(18))]
            open(os.path.join(OUT, "%s_hot_sass_%s.md" % (tag, workload)), "w").write("\n".join(lines) + "\n")
        except ValueError:
            pass
    return out


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    o = full(tag)
    if o:
        for k, v in o["metrics"].items():
            print("%-90s %14.3f %s" % (k, v["mean"], v["unit"]))
        print("dram bytes per launch", o["dram_bytes_per_launch"])
    for f in ("bench_f64.json", "bench_f32.json", "microbench.log"):
        p = os.path.join(SRC, f)
        if os.path.exists(p) and os.path.getsize(p):
            open(os.path.join(OUT, "%s_%s" % (tag, f)), "w").write(open(p).read())
