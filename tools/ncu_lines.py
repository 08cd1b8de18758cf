"""Stall samples per CUDA source line of an .ncu-rep captured with --import-source on (kernels built with -lineinfo).
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
lines = []
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if len(r) > 8 and r[0].isdigit():
        try:
            lines.append((int(r[6]), cur_file, int(r[0]), r[1].strip(), int(r[7] or 0)))
        except ValueError:
            pass
total = sum(l[0] for l in lines) or 1
print(f"total samples {total}")
for s, f, ln, src, ex in sorted(lines, reverse=True)[:top]:
    print(f"{s:7d} {100 * s / total:5.1f}%  {f}:{ln:<5d} exec={ex:<9d} {src[:110]}")
