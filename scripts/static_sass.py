#!/usr/bin/env python
"""Static evidence for kernels that have no ncu capture: registers / shared memory / spills (cuobjdump -res-usage) and the memory
instruction mix of the SASS (cuobjdump -sass) of the shipped libb200grb.so, for the instantiations the default options launch.
Needs no GPU.  Usage: python scripts/static_sass.py > profiles/r2_static_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "falkordb_b200", "libb200grb.so")
# demangled-name prefixes of the instantiations the 512-source chain (W = 8), the BFS engine and the side kernels launch
WANT = [
    ("hop, pull (mid + long-row segments), W = 8", r"k_pull_seg<8, false, 4, false>"),
    ("hop, pull (rows <= 8), W = 8", r"k_pull_small<8, false, 1>"),
    ("hop 2, CSR push", r"k_csr_push<8"),
    ("frontier totals in gather order", r"k_ordered_flops<8>"),
    ("materialise, count (carry-save)", r"k_bits_count_csa<8>"),
    ("materialise, fill v3", r"k_bits_fill_v3("),
    ("bitmap hand-off", r"k_bits_rowmajor"),
    ("BFS top-down expand", r"k_do_expand"),
    ("BFS bottom-up", r"k_do_pull"),
    ("BFS parent assign", r"k_do_assign"),
    ("masked SpGEMM pairs", r"k_masked_pairs"),
    ("set union", r"k_union"),
    ("mask filter", r"k_rowfilter_fill"),
    ("FP64 mxv (PLUS_SECOND)", r"k_mxv_fp64<false>"),
    ("WCC hook", r"k_cc_hook"),
    ("CDLP pick", r"k_cdlp_pick"),
    ("bulk tensor build, run marking", r"k_bulk_mark"),
]
OPS = [("LDG.*\\.256", r"\bLDG\.[A-Z0-9_.]*256"), ("LDG.*\\.128", r"\bLDG\.[A-Z0-9_.]*128"), ("LDG.*\\.64", r"\bLDG\.[A-Z0-9_.]*\.64\b"),
       ("LDG other", r"\bLDG\."), ("STG.*\\.256", r"\bSTG\.[A-Z0-9_.]*256"), ("STG.*\\.128", r"\bSTG\.[A-Z0-9_.]*128"), ("STG other", r"\bSTG\."),
       ("REDG (global reduction)", r"\bREDG\."), ("ATOMG", r"\bATOMG\."), ("ATOMS", r"\bATOMS\."), ("REDUX", r"\bREDUX\b"), ("LDS/STS", r"\b(LDS|STS)\b"), ("SHFL", r"\bSHFL\."), ("VOTE/MATCH", r"\b(VOTE|MATCH)\b"),
       ("POPC", r"\bPOPC\b"), ("PRMT", r"\bPRMT\b"), ("LOP3", r"\bLOP3\b"), ("BAR", r"\bBAR\.")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            cur = None
    dm = demangle(list(usage))
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    bodies, cur = collections.defaultdict(list), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            bodies[cur].append(line)
    print("# Static SASS evidence (no GPU needed): `python scripts/static_sass.py` on the shipped `libb200grb.so` (sm_100a)\n")
    print("Registers / shared memory / spills per thread block and the memory-instruction mix of the instantiations the default options launch.")
    print("`LDG.*.256` / `STG.*.256` are Blackwell's 32-byte global accesses; `REDG` = fire-and-forget global atomics (OR / ADD / MIN); no kernel spills.\n")
    print("| kernel | instantiation | regs | smem B | spill B | SASS lines | " + " | ".join(n for n, _ in OPS) + " |")
    print("|---|---|---|---|---|---|" + "---|" * len(OPS))
    for label, pat in WANT:
        hits = [k for k, v in dm.items() if pat in v]
        if not hits:
            print(f"| {label} | `{pat}` not found | | | | |" + " |" * len(OPS))
            continue
        k = sorted(hits, key=lambda x: len(dm[x]))[0]
        u = usage[k]
        body = bodies.get(k, [])
        counts, seen = [], set()
        for name, rx in OPS:
            c = 0
            for i, ln in enumerate(body):
                if i in seen:
                    continue
                if re.search(rx, ln):
                    c += 1
                    seen.add(i)
            counts.append(c)
        short = dm[k].split("(")[0].replace("void b200::", "").replace("b200::", "")
        print(f"| {label} | `{short}` | {u.get('REG')} | {u.get('SHARED')} | {u.get('LOCAL')} | {len(body)} | " + " | ".join(str(c) for c in counts) + " |")
    spills = [dm[k] for k, u in usage.items() if int(u.get("LOCAL", 0)) > 0]
    print(f"\nKernels in the library: {len(usage)}; with local-memory spills: {len(spills)}" + (": " + "; ".join(s.split('(')[0] for s in spills[:8]) if spills else "") + ".")


if __name__ == "__main__":
    sys.exit(main())
