"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / bulk-copy path (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), UTMALDG (tensor-map
TMA), plus FFMA and register / shared-memory use from the ELF resource usage.  Runs without a GPU:
    python scripts/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "soc_matching_b200", "libsocm_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTMALDG", "SYNCS", "FFMA", "MUFU", "ATOMG", "REDG", "RED."]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[cur][mn] += 1
        counts[cur]["total"] += 1
usage = {}
fn = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and fn:
        usage[fn] = (int(m.group(1)), int(m.group(2)))


def demangle(n):
    try:
        return re.sub(r"\(int\)", "", subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()).split("(")[0]
    except Exception:
        return n


print("# SASS summary of libsocm_b200.so (cuobjdump -sass, sm_100a)\n")
print("| kernel | instr | " + " | ".join(MNEMONICS) + " | regs | static smem |")
print("|---|---:|" + "---:|" * len(MNEMONICS) + "---:|---:|")
for fnm, c in counts.items():
    r = usage.get(fnm, ("", ""))
    print(f"| `{demangle(fnm)}` | {c['total']} | " + " | ".join(str(c[m]) for m in MNEMONICS) + f" | {r[0]} | {r[1]} |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nTotals: " + ", ".join(f"{m} {tot[m]}" for m in MNEMONICS))
print("\nUTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st, UBLKCP = cp.async.bulk "
      "(1-D bulk copies through the TMA engine), UTCBAR = tcgen05.commit -> mbarrier.  No UTMALDG: the operands are "
      "pre-packed tapes / feature blocks moved by 1-D bulk copies, no tensor maps are needed.")
