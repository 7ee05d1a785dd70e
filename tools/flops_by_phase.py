#!/usr/bin/env python
"""Executed FP64 operations of one captured k_solve launch, by source function (phase), from the source page of an
`ncu --set full --import-source on` report + the line table of the library that ran (-lineinfo):

    python tools/flops_by_phase.py <report.ncu-rep> <lib.so> <git rev of the sources> <instances> <algorithmic flops per solve> [kernel prefix]

Counts predicated-on THREAD instructions of DFMA (2 flop), DADD / DMUL (1 flop) and DSETP/other D* (0) per SASS line and
attributes each line to the innermost source function (the function whose definition precedes the line in its file at the
given revision).  Output: share of the executed FP64 flops per function and the executed / algorithmic ratio."""
import collections, csv, io, re, subprocess, sys, shutil
from pathlib import Path

rep, lib, rev, inst, algo = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), float(sys.argv[5])
kname = sys.argv[6] if len(sys.argv) > 6 else "_Z7k_solveILb0EE"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, ith = hdr.index("Address"), hdr.index("Source"), hdr.index("Predicated-On Thread Instructions Executed")
tmp = Path("/tmp/ftmpc_sass2"); shutil.rmtree(tmp, ignore_errors=True); tmp.mkdir()
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", str(next(tmp.glob("*.cubin")))], capture_output=True, text=True).stdout
off2src, cur, on = {}, None, False
for l in sass.splitlines():
    if l.startswith(".text."):
        on = kname in l
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,6})\*/", l)
    if m:
        off2src[int(m.group(1), 16)] = cur
fn_cache = {}
def functions(fname):
    if fname not in fn_cache:
        out = subprocess.run(["git", "show", f"{rev}:fault-tolerant-mpc_b200/csrc/{fname}"], capture_output=True, text=True).stdout
        fs = []
        for i, l in enumerate(out.splitlines(), 1):
            m = re.match(r"(?:FT_HD|FT_D|__device__|__host__|template|static|inline)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
            if m and not l.startswith(" ") and "template" not in l.split("(")[0].split()[-1:]:
                fs.append((i, m.group(1)))
        fn_cache[fname] = fs
    return fn_cache[fname]
def owner(k):
    if not k:
        return "?"
    f, ln = k
    name = "?"
    for i, n in functions(f):
        if i <= ln:
            name = n
    return f"{f}:{name}"
base = int(rows[2][ia], 16)
flops = collections.Counter()
ops = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    s = r[isrc]
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?(D[A-Z0-9]+)", s)
    if not m:
        continue
    op = m.group(1)
    w = 2 if op.startswith("DFMA") else (1 if op.startswith(("DADD", "DMUL")) else 0)
    n = int(r[ith])
    ops[op.split(".")[0]] += n
    if w:
        flops[owner(off2src.get(int(r[ia], 16) - base))] += w * n
tot = sum(flops.values())
print(f"executed FP64 flops of the captured launch: {tot:.4e}  = {tot / inst / 1e6:.2f} MFLOP per solve ({inst} instances)")
print(f"algorithmic (SURVEY 8d model, bench.py): {algo / 1e6:.2f} MFLOP per solve  ->  executed / algorithmic = {tot / inst / algo:.2f}")
print("thread instructions by opcode:", ", ".join(f"{k} {v:.3e}" for k, v in ops.most_common(8)))
print("\nshare of the executed FP64 flops by source function:")
for k, v in flops.most_common(24):
    print(f"  {100 * v / tot:6.2f} %  {v / inst / 1e6:7.3f} MFLOP/solve  {k}")
