"""SASS opcode census of the built library (runs on the CPU: cuobjdump -sass): per kernel the counts of the Blackwell-only
instruction families that show the path is hand-written tcgen05 / TMA code, plus register / spill figures.
    python tools/sass_census.py > profiles/rNN_sass_census.md
Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UTMALDG / UTMASTG = TMA tensor load / store, LDGSTS = cp.async, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = fp32x2,
HSET2 / HMNMX2 / HFMA2(.BF16_V2) = packed bf16."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "modulus_b200", "lib", "libmgn_b200.so")
FAM = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "LDGSTS", "SYNCS", "FFMA2", "FADD2", "FMUL2",
       "HSET2", "HMNMX2", "HFMA2", "HADD2", "F2FP", "SHFL", "LDS", "STS", "FFMA", "FADD", "IMAD"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    regs[m.group(1)] = (m.group(2), m.group(3))
rows = []
for f in re.split(r"\n\s*Function : ", sass)[1:]:
    name = f.split("\n")[0].strip()
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f):
        ops[m.group(1)] += 1
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(.*", "", dem).replace("mgn::", "")
    rows.append((dem, sum(ops.values()), ops, regs.get(name, ("?", "?"))))
rows.sort(key=lambda r: -r[2]["UTCHMMA"] * 100000 - r[1])
print("# SASS opcode census of modulus_b200/lib/libmgn_b200.so (sm_100a), `python tools/sass_census.py`\n")
print("Static instruction counts per kernel (unrolled loops count once per copy).  Kernels with tcgen05 first.\n")
hdr = ["kernel", "instr", "regs", "stack B"] + FAM
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
for dem, tot, ops, (r, st) in rows:
    if tot < 50:
        continue
    print("| `" + dem[:70] + "` | " + " | ".join([str(tot), r, st] + [str(ops[k]) if ops[k] else "" for k in FAM]) + " |")
