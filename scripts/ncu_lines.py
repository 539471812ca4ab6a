"""Top source lines of an .ncu-rep by warp-stall samples: python scripts/ncu_lines.py file.ncu-rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, data = None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) < 9 or r[0] in ("Line No", ""):
        continue
    try:
        data.append((fname, int(r[0]), r[1], int(r[6] or 0), int(r[7] or 0)))
    except ValueError:
        pass
ts = sum(d[3] for d in data)
te = sum(d[4] for d in data)
print("samples", ts, "executed", te)
for d in sorted(data, key=lambda x: -x[3])[:ntop]:
    print(f"{100*d[3]/ts:5.1f}% ex {100*d[4]/te:5.1f}%  {d[0]}:{d[1]:<4d} {d[2].strip()[:100]}")
