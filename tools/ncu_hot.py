"""Top SASS lines by stall samples from an .ncu-rep source page. Usage: ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
r = list(csv.reader(lines[1:]))
h = r[0]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
rows = [(int(x[isamp] or 0), i, x[isrc].strip(), int(x[iex] or 0)) for i, x in enumerate(r[1:]) if len(x) > isamp]
tot = sum(a for a, *_ in rows)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot, "instr lines", len(rows))
for a, i, s, e in sorted(rows, reverse=True)[:N]:
    print(f"{100*a/tot:5.1f}%  line {i:5d}  exec {e:9d}  {s[:110]}")
