"""Run bench.py once per prebuilt library variant (tools/build_variant.py) on the GPU box and print one comparison table.

  python tools/ab_variants.py [--steps 20] base name1 label2:libname2:ENV=1,ENV2=0 ...
Each run: `SUNET_LIB_PATH=<variant .so> python bench.py --no-cpu-baseline --no-parity --no-anyres --profile-json ...`; the table
lists images/s and the average device time of every distinct kernel shape (kind, algorithmic flops) of one forward."""
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
steps = "20"
if args and args[0] == "--steps":
    steps, args = args[1], args[2:]
names = args or ["base"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
res = OrderedDict()
for spec in names:   # label[:lib[:ENV=V,ENV=V]]  (lib "base" or empty = the default library)
    parts = spec.split(":")
    nm = parts[0]
    libname = parts[1] if len(parts) > 1 and parts[1] else nm
    env = dict(os.environ)
    if libname != "base":
        env["SUNET_LIB_PATH"] = os.path.join(ROOT, "sunet_tf_b200", "variants", f"libsunet_{libname}.so")
    if len(parts) > 2:
        for kv in parts[2].split(","):
            k, v = kv.split("=")
            env[k] = v
    pj = os.path.join(ROOT, "gpurun_out", f"ab_{nm}_kernels.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-parity", "--no-anyres", "--steps", steps,
                        "--profile-json", pj], capture_output=True, text=True, env=env, cwd=ROOT)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not line:
        print(f"== {nm}: FAILED rc={r.returncode}\n{r.stderr[-1500:]}")
        continue
    d = json.loads(line[-1])
    with open(os.path.join(ROOT, "gpurun_out", f"ab_{nm}.json"), "w") as fh:
        fh.write(line[-1] + "\n")
    shapes = OrderedDict()
    for kind, ms, fl, by in json.load(open(pj))["launch_list"]:
        k = (kind, fl, by)
        a = shapes.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    res[nm] = (d, shapes)
    print(f"== {nm}: {d['value']:.1f} img/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}", flush=True)
if res:
    first = next(iter(res))
    keys = list(res[first][1].keys())
    print("\nkernel shape (kind, GF, MB) x launches : avg us per variant")
    for k in keys:
        cnt = res[first][1][k][0]
        if res[first][1][k][1] < 0.03:
            continue
        row = "  ".join(f"{nm}={res[nm][1][k][1] / res[nm][1][k][0] * 1e3:7.1f}" if k in res[nm][1] else f"{nm}=   n/a" for nm in res)
        print(f"{k[0]:16s} {k[1] / 1e9:8.2f} GF {k[2] / 1e6:7.1f} MB x{cnt:3d} : {row}")
