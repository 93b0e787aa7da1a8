"""Issue-slot budget of a kernel per 128-row tile from a saved `ncu --set full --import-source on` report.

    python tools/ncu_issue_slots.py gpurun_out/r01_full_edge_bwd2_kernel.ncu-rep [rows=5992002]

Reads the source page (`ncu -i REPORT --page source --csv --print-source cuda,sass`), sums "Instructions Executed" per
source line and splits them into the mbarrier wait loops (`wait_clk` in csrc/mgn_tile.cuh, `mbar_wait` in
csrc/mgn_tc.cuh) and everything else; a B200 SM issues at most one warp instruction per cycle and sub-partition, so
(instructions per tile) / 4 is a lower bound on the tile time in cycles.  Runs without a GPU."""
import csv
import subprocess
import sys

rep = sys.argv[1]
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 5992002
tiles = (n_rows + 127) // 128
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, check=True).stdout
hdr, cur, data = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 3 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        data.append((cur, int(r[0]), r))
ix = {n: i for i, n in enumerate(hdr)}
IE, S = ix["Instructions Executed"], ix["# Samples"]


def _ranges():
    """line ranges of the polling loops, found in the sources (wait_clk in mgn_tile.cuh, mbar_try_wait* / mbar_wait in
    mgn_tc.cuh) so that edits above them do not silently break the split"""
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "modulus_b200", "csrc")
    out = {}
    for fn, pats in (("mgn_tile.cuh", [r"bool wait_clk\("]), ("mgn_tc.cuh", [r"uint32_t mbar_try_wait", r"bool mbar_wait\("])):
        lines = open(os.path.join(root, fn)).read().split("\n")
        spans = []
        for i, ln in enumerate(lines):
            if any(re.search(p_, ln) for p_ in pats):
                j = i
                while j < len(lines) and not lines[j].startswith("}"):
                    j += 1
                spans.append((i + 1, j + 1))
        out[fn] = spans
    return out


_R = _ranges()


def is_wait(c, l):
    return any(a <= l <= b for a, b in _R.get(c, []))


tot = sum(int(r[IE]) for _, _, r in data)
spin = sum(int(r[IE]) for c, l, r in data if is_wait(c, l))
samples = sum(int(r[S]) for _, _, r in data)
wsamples = sum(int(r[S]) for c, l, r in data if is_wait(c, l))
print(f"report {rep}: {tiles} tiles")
print(f"warp instructions per tile: {tot / tiles:.0f} total = {(tot - spin) / tiles:.0f} work + {spin / tiles:.0f} in mbarrier wait loops "
      f"({100 * spin / tot:.0f} %)")
print(f"per SM sub-partition and tile: {(tot - spin) / tiles / 4:.0f} work + {spin / tiles / 4:.0f} wait-loop issue slots")
print(f"stall samples inside the wait loops: {100 * wsamples / samples:.0f} %")
by = {}
for c, l, r in data:
    if not is_wait(c, l):
        by[c] = by.get(c, 0) + int(r[IE])
print("work instructions per tile by source file:")
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:6]:
    print(f"  {k:28s} {v / tiles:8.0f}")
print("hottest source lines outside the wait loops (stall samples):")
for c, l, r in sorted((d for d in data if not is_wait(d[0], d[1])), key=lambda d: -int(d[2][S]))[:8]:
    print(f"  {c}:{l:<4d} {100 * int(r[S]) / samples:4.1f} %  {r[1].strip()[:90]}")
