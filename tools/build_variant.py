"""Build a named variant of libsunet_b200.so with extra nvcc flags, HERE (no GPU needed), for A/B runs on the GPU box:

  python tools/build_variant.py <name> [-DFLAG=1 ...]      -> sunet_tf_b200/variants/libsunet_<name>.so

Only the sources whose text mentions one of the -D macros are recompiled; the rest are taken from csrc/build (the default build).
Variants are git-ignored (*.so) and travel with the gpurun snapshot; tools/ab_variants.py runs bench.py once per variant."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sunet_tf_b200 import _build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
macros = [re.match(r"-D(\w+)", f).group(1) for f in flags if f.startswith("-D")]
_build.build()
vdir = os.path.join(ROOT, "sunet_tf_b200", "variants")
odir = os.path.join(vdir, "obj_" + name)
os.makedirs(odir, exist_ok=True)
nvcc = _build._nvcc()
objs = []
procs = []
for src in _build.SOURCES:
    path = os.path.join(_build.CSRC, src)
    text = open(path).read()
    hdrs = "".join(open(os.path.join(_build.CSRC, h)).read() for h in os.listdir(_build.CSRC) if h.endswith((".cuh", ".h")))
    if any(m in text or m in hdrs for m in macros):   # a macro used in a header recompiles every source
        obj = os.path.join(odir, src.replace(".cu", ".o"))
        procs.append((src, subprocess.Popen([nvcc, *_build.BASE_FLAGS, *flags, "-c", path, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    else:
        obj = os.path.join(_build.OBJ_DIR, src.replace(".cu", ".o"))
    objs.append(obj)
for src, pr in procs:
    out, _ = pr.communicate()
    if pr.returncode != 0:
        raise SystemExit(f"nvcc failed for {src}:\n{out}")
lib = os.path.join(vdir, f"libsunet_{name}.so")
r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
if r.returncode != 0:
    raise SystemExit(r.stdout + r.stderr)
print(lib, "recompiled:", [s for s, _ in procs])
