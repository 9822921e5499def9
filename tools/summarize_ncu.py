#!/usr/bin/env python3
"""Turns the ncu outputs of a gpurun call into the committed evidence under profiles/.

    python tools/summarize_ncu.py --launches gpurun_out/launches_r01.csv --rep gpurun_out/prof_r01_tps.ncu-rep \
        --tag r01 --workload "werewolf-(mafia)_p8_tps" --cmd "python bench.py --steps 240 ..."
Writes profiles/<tag>_launches.md and profiles/<tag>_full_<kernel>.md (physical DRAM traffic: tools/phys_traffic.py)."""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_us(v, u):
    return v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches", required=True)
    ap.add_argument("--rep", default="")
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--workload", default="werewolf-(mafia)_p8_tps")
    ap.add_argument("--kernel", default="k_step_w_tps")
    ap.add_argument("--cmd", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.launches)))
    hdr, data = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            e = data.setdefault(int(d["ID"]), {"k": d["Kernel Name"]})
            v = float(d["Metric Value"].replace(",", ""))
            n, u = d["Metric Name"], d["Metric Unit"]
            e[n] = to_us(v, u) if n.startswith("gpu__time") else to_bytes(v, u) if "bytes" in n else v
    agg = collections.defaultdict(list)
    for e in data.values():
        if e["k"].startswith("k_delay"):          # bench.py's head-start spin kernel: outside the timed region
            continue
        agg[e["k"]].append(e)
    tot = sum(x.get("gpu__time_duration.sum", 0) for v in agg.values() for x in v)
    out = ["# %s — ncu launch list" % a.tag, "", "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "smsp__inst_executed.sum,... --clock-control none --csv %s`" % a.cmd,
           "(per-launch times under ncu are cold-cache and serialised — compare SHARES, not absolutes; dram write bytes are "
           "write-backs that reached DRAM during the kernel, most dirty lines are still in L2 when it ends)", "",
           "| kernel | launches | total us | avg us | share of GPU time | avg DRAM read MB | avg DRAM write MB | avg warp-instructions |",
           "|---|---|---|---|---|---|---|---|"]
    traffic = None
    for k, v in sorted(agg.items(), key=lambda kv: -sum(x.get("gpu__time_duration.sum", 0) for x in kv[1])):
        t = sum(x.get("gpu__time_duration.sum", 0) for x in v)
        rd = sum(x.get("dram__bytes_read.sum", 0) for x in v) / len(v)
        wr = sum(x.get("dram__bytes_write.sum", 0) for x in v) / len(v)
        ins = sum(x.get("smsp__inst_executed.sum", 0) for x in v) / len(v)
        out.append("| `%s` | %d | %.1f | %.2f | %.3f | %.2f | %.2f | %.0f |" % (k[:60], len(v), t, t / len(v), t / tot, rd / 1e6, wr / 1e6, ins))
        if a.kernel in k and traffic is None:
            traffic = rd + wr
    out += ["", "Step-kernel launches in issue order (us, DRAM read MB, warp-instructions):", "", "```"]
    for i, e in list(data.items())[:400]:
        if a.kernel in e["k"] or "compact" in e["k"]:
            out.append("%4d %-22s %6.1f %6.2f %9.0f" % (i, e["k"].split("(")[0][-22:], e.get("gpu__time_duration.sum", 0),
                                                     e.get("dram__bytes_read.sum", 0) / 1e6, e.get("smsp__inst_executed.sum", 0)))
    out.append("```")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "%s_launches.md" % a.tag), "w") as f:
        f.write("\n".join(out) + "\n")
    # (profiles/traffic.json is written by tools/phys_traffic.py from a capture with the caches left alone; the cold-cache
    # per-launch read bytes of this listing do not include write-backs and are not used by bench.py)
    print("\n".join(out[:16]))
    print("traffic per launch:", traffic)
    if a.rep:
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        h, u = rr[0], rr[1]
        ix = {n: i for i, n in enumerate(h)}
        want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
                "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"] + [
                "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % r for r in (
                    "long_scoreboard", "wait", "short_scoreboard", "not_selected", "math_pipe_throttle", "no_instruction",
                    "branch_resolving", "barrier", "dispatch_stall", "lg_throttle", "mio_throttle")]
        lines = ["# %s — ncu `--set full` capture of `%s`" % (a.tag, a.kernel), "",
                 "Command (after the same command exited 0 without ncu): `ncu --set full --clock-control none --import-source on "
                 "-k regex:%s ... %s`" % (a.kernel, a.cmd), "", "| metric | unit | " + " | ".join("L%d" % i for i in range(len(rr) - 2)) + " |",
                 "|---|---|" + "---|" * (len(rr) - 2)]
        for w in want:
            if w in ix:
                lines.append("| %s | %s | " % (w, u[ix[w]]) + " | ".join(r[ix[w]][:8] for r in rr[2:]) + " |")
        with open(os.path.join(ROOT, "profiles", "%s_full_%s.md" % (a.tag, a.kernel)), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
