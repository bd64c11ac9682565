"""Join an ncu report's per-instruction stall samples with CUDA source lines (nvdisasm -g line info of the in-tree .so).

  python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [top_n]
Prints the hottest source lines (sum of stall samples over their SASS instructions) and the hottest instructions.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
by_instr = len(sys.argv) > 4 and sys.argv[4] == "instr"   # rank lines by executed warp instructions instead of samples

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may be in the report: take the first whose name matches
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = next(b for b in blocks if pat in b["name"])
hdr, data = blk["rows"][0], blk["rows"][1:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "sunet_tf_b200", "libsunet_b200.so")], cwd=tmp, capture_output=True)
mangled = None
lines = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    secs = re.split(r"\n//-+ \.text\.", txt)
    for s in secs[1:]:
        name = s.split(" ", 1)[0]
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
        if pat in dem:
            cand = []
            curline = None
            for ln in s.splitlines():
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if m:
                    curline = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                    cand.append(curline)
            if len(cand) == len(data) or (not lines and mangled is None):
                if len(cand) == len(data):
                    mangled = name
                lines = cand
if not lines or len(lines) != len(data):
    print(f"warning: {len(lines)} disassembled instructions vs {len(data)} profiled (kernel {mangled})")
n = min(len(lines), len(data))
tot = sum(int(r[isamp]) for r in data)
by_line = {}
for i in range(n):
    k = lines[i]
    d = by_line.setdefault(k, [0, 0, {}])
    d[0] += int(data[i][isamp])
    d[1] += int(data[i][iex])
    for c in stall_cols:
        v = int(data[i][c] or 0)
        if v:
            d[2][hdr[c]] = d[2].get(hdr[c], 0) + v
print(f"kernel {blk['name'][:100]}\ntotal samples {tot}")
srcs = {}
tot_ex = sum(v[1] for v in by_line.values())
print(f"total warp instructions {tot_ex}")
for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][1 if by_instr else 0])[:topn]:
    if k is None:
        print(f"{v[0]:6d} {100*v[0]/tot:5.1f}%  <no line>")
        continue
    fn, ln = k
    if fn not in srcs:
        p = os.path.join(ROOT, "sunet_tf_b200", "csrc", fn)
        srcs[fn] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[fn][ln - 1].strip()[:90] if ln - 1 < len(srcs[fn]) else ""
    st = ", ".join(f"{a[6:]}={b}" for a, b in sorted(v[2].items(), key=lambda ab: -ab[1])[:3])
    print(f"{v[0]:6d} {100*v[0]/tot:5.1f}% | {v[1]:9d} {100*v[1]/tot_ex:5.1f}%  {fn}:{ln:<4d} {text}   [{st}]")
