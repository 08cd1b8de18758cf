"""Summarise the SASS-level source page of an .ncu-rep: top instructions by stall samples and totals per stall reason.
   python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        samples = int(r[col["# Samples"]])
    except ValueError:
        continue
    data.append((samples, r))
total = sum(s for s, _ in data) or 1
print(f"total samples {total}, instructions {len(data)}")
tot = {n: 0 for n in stall_cols}
for s, r in data:
    for n in stall_cols:
        try:
            tot[n] += int(r[col[n]])
        except ValueError:
            pass
print("by reason:", ", ".join(f"{n[6:]}={v} ({100 * v / total:.0f}%)" for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
print(f"{'idx':>5s} {'samples':>8s} {'%':>5s} {'exec':>9s}  instruction / dominant stalls")
for idx, (s, r) in sorted(enumerate(data), key=lambda t: -t[1][0])[:top]:
    st = sorted(((int(r[col[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:3]
    print(f"{idx:5d} {s:8d} {100 * s / total:5.1f} {r[col['Instructions Executed']]:>9s}  {r[col['Source']].strip()[:70]:70s} "
          + " ".join(f"{n}:{v}" for v, n in st if v))
if "--regions" in sys.argv:
    step = 100
    print("samples per 100-instruction region:")
    for i in range(0, len(data), step):
        s = sum(x for x, _ in data[i:i + step])
        if s * 100 >= total:
            ops = {}
            for _, r in data[i:i + step]:
                op = r[col["Source"]].strip().split()
                op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
                ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
            top_ops = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
            print(f"  [{i:5d},{i + step:5d}) {s:7d} {100 * s / total:5.1f}%  {top_ops}")
