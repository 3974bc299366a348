#!/usr/bin/env python3
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked text summary under profiles/.
usage: python scripts/summarise_profiles.py <tag> <launch-list.csv> <bench.json> <report.ncu-rep>..."""
import collections
import csv
import json
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def short(n):
    n = re.sub(r"\(.*", "", n).replace("b200::", "").replace("void ", "")
    return n[:90]


def launches(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    ix = {c: i for i, c in enumerate(rows[hi])}
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        out.append((short(r[ix["Kernel Name"]]), v / 1000 if u.startswith("n") else v * 1000 if u.startswith("m") else v))
    return out


def table(seg, title, out):
    tot = sum(u for _, u in seg)
    out.append(f"{title}: {len(seg)} launches, {tot:.1f} us")
    agg = collections.OrderedDict()
    for n, u in seg:
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += u
        agg[n][1] += 1
    for k, (u, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if u / tot < 0.001:
            continue
        out.append(f"  {u:9.1f} us {100 * u / tot:5.1f}%  n={c:3d}  {k}")


def main():
    tag, lpath, bpath, reps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
    b0 = json.load(open(bpath))
    out = [f"# {tag}: ncu evidence for `python bench.py` (B200, RMAT-24 ef16, {b0['config']['sources_per_gpu_per_step']} sources/batch, 3-hop chain)",
           "# launch list: ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv python bench.py --steps 2 --warmup 1"
           " --no-cpu-baseline --e2e-format csr", ""]
    recs = launches(lpath)
    table(recs, "whole capture (includes the untimed setup: RMAT generation, sorts, transpose, hot-set tables)", out)
    fills = [i for i, (n, _) in enumerate(recs) if "k_bits_fill" in n]
    if len(fills) >= 2:
        out.append("")
        table(recs[fills[0] + 1: fills[1] + 1], "one chain evaluation of the device-resident arm (between two materialise fills)", out)
    traffic = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        h, u, v = rows[0], rows[1], rows[2]
        ix = {c: i for i, c in enumerate(h)}
        name = short(v[ix["Kernel Name"]])
        out += ["", f"## ncu --set full --clock-control none: {name}"]
        for m in WANT:
            if m in ix:
                out.append(f"    {m} = {v[ix[m]]} {u[ix[m]]}")

        def by(m):
            x, unit = float(v[ix[m]]), u[ix[m]].lower()
            return x * (1e9 if unit.startswith("g") else 1e6 if unit.startswith("m") else 1e3 if unit.startswith("k") else 1)
        traffic[name.split("<")[0]] = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
    b = json.load(open(bpath))
    out += ["", "# bench.py line of the same configuration (CUDA-event timed, not under ncu):",
            "  " + json.dumps({k: b[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}),
            "  roofline: " + json.dumps(b["roofline"]), "  e2e: " + json.dumps(b["e2e"]),
            "  kernels (ms per step): " + json.dumps({k: round(x["ms"] / b["steps"], 3) for k, x in b["kernels"].items()}),
            "  cpu_baseline: " + json.dumps(b.get("cpu_baseline"))]
    out += ["", "# measured DRAM bytes per launch (ncu, read + write): " + json.dumps({k: round(x / 1e9, 3) for k, x in traffic.items()}) + " GB"]
    open(f"profiles/{tag}_launches_and_kernels.txt", "w").write("\n".join(out) + "\n")
    json.dump({"config": {"scale": 24, "edge_factor": 16, "sources": b["config"]["sources_per_gpu_per_step"]}, "dram_bytes_by_kernel": traffic},
              open(f"profiles/{tag}_traffic.json", "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
