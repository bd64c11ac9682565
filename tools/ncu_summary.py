"""Summarise an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / bench.py quote.

  python tools/ncu_summary.py <report.ncu-rep> [more.ncu-rep ...] > profiles/rNN_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"== {rep}: no data")
        continue
    hdr, units = rows[0], rows[1]
    print(f"== {rep}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"-- {name[:110]}")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {label:34s} {r[i]:>16s} {units[i]}")
