#!/usr/bin/env python3
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked text summaries under profiles/.
usage: python scripts/summarise_r2.py <tag> --launches <csv> [--bench <bench.json>] [--bfs-launches <csv>] <report.ncu-rep>..."""
import argparse
import collections
import csv
import json
import re
import subprocess

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


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


def table(seg, title, out, floor=0.001):
    tot = sum(u for _, u in seg) or 1.0
    out.append(f"{title}: {len(seg)} launches, {tot:.1f} us")
    agg = collections.OrderedDict()
    for n, u in seg:
        agg.setdefault(n, [0.0, 0])
        agg[n][0] += u
        agg[n][1] += 1
    for k, (u, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if u / tot < floor:
            continue
        out.append(f"  {u:9.1f} us {100 * u / tot:5.1f}%  n={c:3d}  {k}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--bench")
    ap.add_argument("--bfs-launches")
    ap.add_argument("--sources", type=int, default=512)
    ap.add_argument("reps", nargs="*")
    a = ap.parse_args()
    out = [f"# {a.tag}: ncu evidence (B200, RMAT-24 ef16, {a.sources} sources/batch, 3-hop chain)",
           "# launch list: ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr",
           "# (per-launch times under ncu are cold-cache and serialised: read the SHARES, the absolute numbers are in the bench line)", ""]
    if a.launches:
        recs = launches(a.launches)
        table(recs, "whole capture (includes the untimed setup: RMAT generation, sorts, transpose, pull tables)", out)
        fills = [i for i, (n, _) in enumerate(recs) if "k_bits_fill" in n]
        if len(fills) >= 2:
            out.append("")
            table(recs[fills[0] + 1: fills[1] + 1], "one chain evaluation of the device-resident arm (between two materialise fills)", out)
    if a.bfs_launches:
        out.append("")
        recs = launches(a.bfs_launches)
        bfs = [r for r in recs if r[0].startswith("k_do_") or "Scan" in r[0] or "k_small_read" in r[0]]
        table(bfs, "BFS sweep, RMAT-26, single GPU (kernels of the BFS engine only, all levels of all sources in the capture)", out)
    traffic = {}
    for rep in a.reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        h, u = rows[0], rows[1]
        ix = {c: i for i, c in enumerate(h)}
        for v in rows[2:]:
            name = short(v[ix["Kernel Name"]])
            out += ["", f"## ncu --set full --clock-control none: {name}   ({rep.split('/')[-1]})"]
            for m in WANT:
                if m in ix:
                    out.append(f"    {m} = {v[ix[m]]} {u[ix[m]]}")

            def by(m):
                x, unit = float(v[ix[m]]), u[ix[m]].lower()
                return x * (1e9 if unit.startswith("g") else 1e6 if unit.startswith("m") else 1e3 if unit.startswith("k") else 1)
            traffic[name.split("<")[0]] = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
    if a.bench:
        b = json.loads(open(a.bench).read().strip().splitlines()[-1])
        out += ["", "# bench.py line of the same configuration (CUDA-event timed, not under ncu):",
                "  " + json.dumps({k: b[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}),
                "  roofline: " + json.dumps(b["roofline"]), "  e2e: " + json.dumps(b["e2e"]),
                "  kernels (ms per launch): " + json.dumps({k: round(x["ms"] / x["launches"], 3) for k, x in b["kernels"].items()}),
                "  cpu_baseline: " + json.dumps(b.get("cpu_baseline"))]
    out += ["", "# measured DRAM bytes per launch (ncu, read + write): " + json.dumps({k: round(x / 1e9, 3) for k, x in traffic.items()}) + " GB"]
    open(f"profiles/{a.tag}_launches_and_kernels.txt", "w").write("\n".join(out) + "\n")
    if traffic:
        json.dump({"config": {"scale": 24, "edge_factor": 16, "sources": a.sources}, "dram_bytes_by_kernel": traffic},
                  open(f"profiles/{a.tag}_traffic.json", "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
