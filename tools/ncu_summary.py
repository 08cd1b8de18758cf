"""Key metrics of every kernel in an .ncu-rep (from `ncu --set full`), one block per launch.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (sm__pipe_tensor_cycles_active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
print(f"# {rep}: ncu --set full --clock-control none (per launch; cold-ish caches, replayed passes)")
for r in rows[2:]:
    print(f"\nkernel: {r[col['Kernel Name']][:110]}")
    for k, label in KEYS:
        if k in col:
            print(f"  {label:58s} {r[col[k]]:>16s} {units[col[k]]}")
