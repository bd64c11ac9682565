"""In-tree build of libsunet_b200.so (nvcc, sm_100a only).  No JIT cache: the .so sits next to this file so that
it travels with the source tree; it is rebuilt when a source is newer than the library or the compile flags changed.

Safe under torchrun: every process may call build(); an exclusive file lock serialises them, object and link outputs
go to per-process temporary names and are renamed into place, and the processes that waited find a fresh library and
return without compiling."""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "csrc", "build")
LIB_PATH = os.path.join(HERE, "libsunet_b200.so")
STAMP_PATH = os.path.join(OBJ_DIR, "flags.stamp")
LOCK_PATH = os.path.join(HERE, ".build.lock")
SOURCES = ["error.cu", "gemm_tcgen05.cu", "attn_core.cu", "attn_core_tc.cu", "attn_fused.cu", "mlp_fused.cu", "mlp_row.cu", "proj_ln.cu", "tail_fused.cu", "elementwise.cu", "tiles.cu", "model.cu"]
BASE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_flags():
    # SUNET_NVCC_EXTRA: e.g. -DSUNET_KERNEL_TIMING=1 for the phase-cycle counters
    return BASE_FLAGS + os.environ.get("SUNET_NVCC_EXTRA", "").split()


def _flags_digest():
    return hashlib.sha256(" ".join(nvcc_flags()).encode()).hexdigest()


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libsunet_b200.so cannot be built")
    return exe


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "sunet_b200.h"))
    if any(os.path.getmtime(d) > lib_m for d in deps if os.path.exists(d)):
        return True
    # a library without a stamp (shipped prebuilt to a box without nvcc history) is taken as built with the default flags
    if os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as fh:
            return fh.read().strip() != _flags_digest()
    return bool(os.environ.get("SUNET_NVCC_EXTRA", "").split())


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libsunet_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():   # another process built it while this one waited for the lock
                return LIB_PATH
            tag = f".{os.getpid()}.tmp"
            flags = nvcc_flags()

            def compile_one(src):
                obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
                tmp_obj = obj + tag
                r = subprocess.run([nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", tmp_obj], capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
                if verbose and r.stderr.strip():
                    print(r.stderr, file=sys.stderr)
                os.replace(tmp_obj, obj)
                return obj

            with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
                objs = list(ex.map(compile_one, SOURCES))
            tmp = LIB_PATH + tag
            r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH + tag, "w") as fh:
                fh.write(_flags_digest() + "\n")
            os.replace(STAMP_PATH + tag, STAMP_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
