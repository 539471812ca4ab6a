"""Summarise ncu outputs into profiles/: (1) a launch-list CSV (--metrics gpu__time_duration.sum) -> per-kernel
totals and shares; (2) a --set full report -> a few headline metrics per kernel.
usage: python scripts/ncu_summary.py launches <csv> <out.md> "<command>"
       python scripts/ncu_summary.py full <rep> <out.md>"""
import csv, subprocess, sys, collections, re

def launches(path, out, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        name = re.sub(r"\(.*", "", r[ki])[:90]
        tot[name] += v_ms
        cnt[name] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\nCommand: `{cmd}` on one B200.\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / total:.1f}% |\n")

def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic"]
    tens = [h for h in hdr if "pipe_tensor" in h and "pct" in h]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}` (--clock-control none)\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')][:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in want + tens:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
            f.write("\n")

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
