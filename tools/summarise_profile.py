#!/usr/bin/env python
"""Turn the scratch artefacts of one `tools/gpu_round.sh <tag>` pass (gpurun_out/) into the tracked summaries under
profiles/:  <tag>_bench.json, <tag>_launches.csv + <tag>_launch_shares.txt, <tag>_k_solve_ncu_full_summary.csv,
<tag>_k_solve_stalls_by_source.txt (warp-stall samples of the full capture attributed to source lines / barrier call
sites) and traffic.json (DRAM bytes of the captured launch, read by bench.py for roofline.traffic).
Usage: python tools/summarise_profile.py <tag> [instances_in_capture] [mangled kernel prefix]"""
import collections, csv, io, json, re, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
KEEP = """dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed launch__block_size launch__grid_size
launch__occupancy_limit_shared_mem launch__registers_per_thread launch__shared_mem_per_block_dynamic
sass__inst_executed_local_loads sass__inst_executed_local_stores sass__inst_executed_shared_loads sass__inst_executed_shared_stores
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active""".split()


def launches(tag):
    src = OUT / f"launches_{tag}.csv"
    if not src.exists():
        return
    shutil.copy(src, PROF / f"{tag}_launches.csv")
    rows = [r for r in csv.reader(l for l in src.read_text().splitlines() if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik])[:70]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --batch 592 --steps 1 --warmup 3 --no-latency --no-cpu-baseline"]
    lines += [f"{k:70s} launches={cnt[k]:4d} total_ms={v:10.3f} share={v / total:.4f}" for k, v in tot.most_common(12)]
    (PROF / f"{tag}_launch_shares.txt").write_text("\n".join(lines) + "\n")


def full(tag, instances):
    rep = OUT / f"prof_{tag}.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out = [["metric", "unit", "value"], ["Kernel Name", "", vals[col["Kernel Name"]]]]
    for m in KEEP:
        if m in col:
            out.append([m, units[col[m]], vals[col[m]].replace(",", "")])
    with open(PROF / f"{tag}_k_solve_ncu_full_summary.csv", "w", newline="") as f:
        csv.writer(f).writerows(out)
    d = {r[0]: r for r in out}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(d["dram__bytes_read.sum"][2]) * scale[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][2]) * scale[d["dram__bytes_write.sum"][1]]
    (PROF / "traffic.json").write_text(json.dumps({
        "kernel": "k_solve", "capture": f"ncu --set full, bench.py --batch 592 --steps 1, 4th k_solve launch (profiles/{tag}_k_solve_ncu_full_summary.csv)",
        "instances_in_capture": instances, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "dram_bytes_per_solve": (rd + wr) / instances}, indent=1) + "\n")
    # warp-stall samples by source line (needs the in-tree library the capture ran with: -lineinfo)
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, isamp = hdr.index("Address"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lib = ROOT / "fault-tolerant-mpc_b200" / "csrc" / "libftmpc.so"
    tmp = Path("/tmp/ftmpc_sass"); shutil.rmtree(tmp, ignore_errors=True); tmp.mkdir()
    subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=tmp, capture_output=True)
    sass = subprocess.run(["nvdisasm", "-g", "-c", str(next(tmp.glob("*.cubin")))], capture_output=True, text=True).stdout
    kname = KNAME
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
    base = int(rows[2][ia], 16)
    per_line, per_stall, total = collections.Counter(), collections.defaultdict(collections.Counter), 0
    bars = []
    offs = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        off = int(r[ia], 16) - base
        s = int(r[isamp]); total += s
        k = off2src.get(off)
        per_line[k] += s
        for i in stall:
            per_stall[k][hdr[i]] += int(r[i])
        offs.append((off, k, int(r[hdr.index("stall_barrier")])))
    for j, (off, k, b) in enumerate(offs):
        if b > 0.002 * total:
            ctx = []
            i = j - 1
            while i >= 0 and len(ctx) < 2:
                kk = offs[i][1]
                if kk and kk[0] != "ftmpc_block.cuh" and kk not in ctx:
                    ctx.append(kk)
                i -= 1
            bars.append((b, ctx))
    lines = [f"warp-stall samples of the full capture ({total} samples; share of all samples = share of warp time)", "",
             "barrier waits by call site (source lines executed just before the barrier):"]
    for b, ctx in sorted(bars, key=lambda x: -x[0])[:14]:
        lines.append(f"  {100 * b / total:5.2f} %  after " + ", ".join(f"{f}:{n}" for f, n in ctx))
    lines.append(f"  total stall_barrier: {100 * sum(o[2] for o in offs) / total:.1f} %")
    lines += ["", "source lines by samples (innermost inlined frame), top stall reasons:"]
    for k, s in per_line.most_common(70):
        top = ", ".join(f"{n.replace('stall_', '')} {100 * v / max(s, 1):.0f}%" for n, v in per_stall[k].most_common(4))
        lines.append(f"  {100 * s / total:5.2f} %  {k[0] if k else '?'}:{k[1] if k else 0}   [{top}]")
    (PROF / f"{tag}_k_solve_stalls_by_source.txt").write_text("\n".join(lines) + "\n")


KNAME = "_Z7k_solveILb0EE"    # mangled prefix of the captured kernel (k_solve2: "_Z8k_solve2")

if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 3:
        KNAME = sys.argv[3]
    inst = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    PROF.mkdir(exist_ok=True)
    b = OUT / f"bench_{tag}.json"
    if b.exists():
        shutil.copy(b, PROF / f"{tag}_bench.json")
    c = OUT / f"clocks_{tag}.csv"
    if c.exists():
        shutil.copy(c, PROF / f"{tag}_clocks.csv")
    launches(tag)
    full(tag, inst)
    print("profiles/ updated for", tag)
