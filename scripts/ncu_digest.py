"""Digest of an .ncu-rep: headline metrics, executed-instruction mix by opcode and the top stall lines.
python scripts/ncu_digest.py file.ncu-rep [n_top]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct", "launch__grid_size", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled"]
for h, u, v in zip(hdr, units, vals):
    if any(h == w or (w.endswith("pct") and h.startswith(w)) or (w == "smsp__average_warps_issue_stalled" and h.startswith(w)) for w in want):
        if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.2:
            continue
        print(f"{h:92s} {v:>18s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    if len(r) <= iexec:
        continue
    try:
        data.append((r[isrc], int(r[isamp] or 0), int(r[iexec] or 0)))
    except ValueError:
        pass
te, ts = sum(d[2] for d in data), sum(d[1] for d in data)
op, ops = collections.Counter(), collections.Counter()
for s, sm, ex in data:
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", s)
    o = m.group(2) if m else "?"
    op[o] += ex
    ops[o] += sm
print(f"\nexecuted warp-instructions {te}, samples {ts}; by opcode (executed %, samples %):")
for o, c in op.most_common(ntop):
    print(f"  {o:12s} {100*c/te:5.1f}%  {100*ops[o]/ts:5.1f}%")
