"""Decode the event trace a timing build of mlp_proj_fused writes with SUNET_MLP_TRACE=1 (tools/phase_timing.sh).
  python tools/decode_mlp_trace.py gpurun_out/trace.log [C=96] [which launch=1] [t_lo] [n events]"""
import sys
path = sys.argv[1]
C = sys.argv[2] if len(sys.argv) > 2 else "96"
which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t_lo = int(sys.argv[4]) if len(sys.argv) > 4 else 70000
nmax = int(sys.argv[5]) if len(sys.argv) > 5 else 140
lines = open(path).read().splitlines()
idx = [k for k, l in enumerate(lines) if l.startswith(f"TRACE mlp_proj_fused<{C}>")][which]
ev = []
for l in lines[idx + 1:]:
    if not l.startswith("TR "):
        break
    _, t, w, e = l.split()
    ev.append((int(t), int(w), int(e, 16)))
ev.sort()
names = {0: "p_full ok", 1: "epi0 done (x1_ready)", 2: "h_full ok", 3: "gelu math done", 4: "tmem ld done (h_empty)", 5: "store done (gelu_done)",
         6: "y_full ok", 7: "out done", 0x5f: "mma0 begin", 0x6f: "fc1thr: sees h_full(0) complete", 0x60: "mma0 issued (p_full commit)"}
for j in range(8):
    names[0x50 + j] = f"fc1({j}) issued"
    names[0x70 + j] = f"fc2({j}) issued"
thr = {1: {0x40: "fc1thr: x1_ready ok", 0x41: "fc1thr: h_empty ok", 0x42: "fc1thr: r1_full ok"},
       18: {0x40: "fc2thr: gelu_done ok", 0x41: "fc2thr: r2_full ok", 0x42: "fc2thr: x_full ok", 0x43: "fc2thr: y_empty ok"}}
cnt = 0
for t, w, e in ev:
    if t < t_lo:
        continue
    print(f"{t - t_lo:7d} w{w:2d} {thr.get(w, {}).get(e, names.get(e, hex(e)))}")
    cnt += 1
    if cnt >= nmax:
        break
