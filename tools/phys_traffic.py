#!/usr/bin/env python3
"""Physical DRAM traffic of the steady-state ring from a one-pass ncu capture with the caches left alone.

    ncu --cache-control none --clock-control none --profile-from-start off \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum \
        --csv --log-file gpurun_out/X.csv python tools/range_profile.py [...] > gpurun_out/X.log
    python tools/phys_traffic.py --csv gpurun_out/X.csv --log gpurun_out/X.log --plain gpurun_out/X_plain.json --tag r02_cfg2

With `--cache-control none` nothing is flushed or invalidated around a kernel, so the write-backs of one launch are
counted in whichever later launch they reach DRAM in: per-kernel attribution is blurred, the SUM over a region that
is many times larger than L2 is what crossed the DRAM pins.  Writes profiles/<tag>_phys_traffic.md and the entry
`<game>_p<P>_<kernel>_n<sessions>` of profiles/traffic.json (bytes per counted session-phase-step, read and write;
mean bytes per step-kernel launch), which bench.py turns into the physical DRAM rate of its own live run."""
import argparse
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON line in " + path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", required=True)
    ap.add_argument("--log", required=True, help="stdout of range_profile.py under ncu (its JSON line: counted steps of the region)")
    ap.add_argument("--plain", default="", help="stdout of the same command without ncu (region time in the concurrent state)")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--kernel", default="tps")
    a = ap.parse_args()
    run = last_json(a.log)
    plain = last_json(a.plain) if a.plain else None
    rows = list(csv.reader(open(a.csv)))
    hdr, data = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            e = data.setdefault(int(d["ID"]), {"k": d["Kernel Name"]})
            e[d["Metric Name"]] = float(d["Metric Value"].replace(",", "")) * SCALE.get(d["Metric Unit"], 1)
    agg = collections.OrderedDict()
    for e in data.values():
        name = e["k"].split("(")[0].replace("void ", "")
        g = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "inst": 0.0})
        g["n"] += 1
        g["us"] += e.get("gpu__time_duration.sum", 0)
        g["rd"] += e.get("dram__bytes_read.sum", 0)
        g["wr"] += e.get("dram__bytes_write.sum", 0)
        g["inst"] += e.get("smsp__inst_executed.sum", 0)
    rd = sum(g["rd"] for g in agg.values())
    wr = sum(g["wr"] for g in agg.values())
    inst = sum(g["inst"] for g in agg.values())
    steps = run["counted_steps"]
    step_k = [g for k, g in agg.items() if "k_step_" in k or "k_ring_" in k]
    n_step_launches = sum(g["n"] for g in step_k)
    key = "%s_p%d_%s_n%d" % (run["game"], run["players"], a.kernel, run["sessions_per_batch"])
    if run.get("launch", "streams") != "streams":              # the ring launch is profiled for DESIGN, bench.py looks up the default mode
        key += "_" + run["launch"]
    entry = {
        "read_bytes_per_step": rd / steps, "write_bytes_per_step": wr / steps, "bytes_per_step": (rd + wr) / steps,
        "bytes_per_step_launch": (rd + wr) / max(1, n_step_launches), "warp_instructions_per_step": inst / steps,
        "sessions_per_batch": run["sessions_per_batch"], "ring": run["ring"],
        "resident_bytes": run["ring"] * run["sessions_per_batch"] * run["record_bytes"],
        "counted_steps": steps, "step_launches": n_step_launches, "all_launches": sum(g["n"] for g in agg.values()),
        "algorithmic_bytes_per_step": 2 * run["record_bytes"],
        "source": "%s: sum of dram__bytes_read.sum / dram__bytes_write.sum over ALL %d launches of %d ring passes "
                  "(ncu --cache-control none, one counter pass; %s)" % (os.path.basename(a.csv), sum(g["n"] for g in agg.values()),
                                                                       run["passes"], run.get("launch", "streams")),
    }
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tj))
    except Exception:
        t = {}
    t[key] = entry
    json.dump(t, open(tj, "w"), indent=1, sort_keys=True)
    out = ["# %s — physical DRAM traffic of the steady-state ring" % a.tag, "",
           "Workload: %s, %d players, ring of %d x %d sessions, %s; region = %d passes over the ring, %d counted session-phase-steps."
           % (run["game"], run["players"], run["ring"], run["sessions_per_batch"], run.get("launch", "streams"), run["passes"], steps), "",
           "ncu: `--cache-control none --clock-control none --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__time_duration.sum,smsp__inst_executed.sum` (one pass, no cache flush between kernels: write-backs are counted when they "
           "reach DRAM, so per-kernel rows are blurred and the TOTAL is what crossed the pins).", "",
           "| kernel | launches | serialised us | DRAM read MB | DRAM write MB | warp-instructions |", "|---|---|---|---|---|---|"]
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append("| `%s` | %d | %.1f | %.1f | %.1f | %.3e |" % (k[:60], g["n"], g["us"], g["rd"] / 1e6, g["wr"] / 1e6, g["inst"]))
    out += ["", "| | bytes | per counted step |", "|---|---|---|",
            "| DRAM read | %.1f MB | %.1f B |" % (rd / 1e6, rd / steps), "| DRAM write | %.1f MB | %.1f B |" % (wr / 1e6, wr / steps),
            "| **read + write** | **%.1f MB** | **%.1f B** |" % ((rd + wr) / 1e6, (rd + wr) / steps),
            "| algorithmic (2·S, SURVEY 8d) | %.1f MB | %d B |" % (steps * 2 * run["record_bytes"] / 1e6, 2 * run["record_bytes"]),
            "| warp-instructions | %.3e | %.2f |" % (inst, inst / steps)]
    if plain:
        ms = plain["region_ms"]
        out += ["", "The same program without ncu (concurrent state, CUDA events): region %.3f ms, %.3e steps/s." % (ms, plain["steps_per_s"]),
                "Physical DRAM rate = %.1f MB / %.3f ms = **%.0f GB/s** = %.3f of the measured copy bandwidth (6544.7 GB/s), %.3f of the nominal 8000."
                % ((rd + wr) / 1e6, ms, (rd + wr) / (ms * 1e-3) / 1e9, (rd + wr) / (ms * 1e-3) / 1e9 / 6544.7, (rd + wr) / (ms * 1e-3) / 1e9 / 8000.0),
                "Issue rate = %.3e warp-instructions / %.3f ms = %.3e /s = %.2f of the scheduler peak (148 SMs x 4 x 1.965 GHz = 1.163e12 /s)."
                % (inst, ms, inst / (ms * 1e-3), inst / (ms * 1e-3) / 1.163e12)]
    open(os.path.join(ROOT, "profiles", a.tag + "_phys_traffic.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-12:]))
    print("traffic.json[%s] = %.1f B/step (read %.1f + write %.1f)" % (key, entry["bytes_per_step"], entry["read_bytes_per_step"], entry["write_bytes_per_step"]))


if __name__ == "__main__":
    main()
