"""Build libjmc_b200.so (sm_100a only) in-tree with nvcc.

    python jetmontecarlo_b200/csrc/build.py [--force] [--verbose]

Each translation unit has its own floating-point contract:
  jmc_exact.cu   --fmad=false   FP64 in NumPy operation order (replay parity, bit-exact bins)
  jmc_fast.cu    FMA + fast intrinsics: the FP32 log-space throughput path
  jmc_shower.cu  --fmad=false   veto-algorithm shower (FP64 chain, checkable against the oracle)
  jmc_process.cu --fmad=false   hit-or-miss process sampling, pairwise observables, EWOC histograms
  jmc_sudakov.cu --fmad=false   per-event inverse-transform sampling from tabulated conditional cdfs
  jmc_host.cu    host plumbing and the host-buffer entry points
The objects are linked into jetmontecarlo_b200/lib/libjmc_b200.so, which
travels to the GPU box with the repository snapshot (it is git-ignored).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(LIB_DIR, "libjmc_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", os.path.join(ROOT, "include")]
# compile-time experiment switches (scripts/fast_shootout.sh, scripts/shower_shootout.sh); unset = the defaults in the sources
SWITCHES = ("JMC_FAST_CHUNK", "JMC_FAST_MIN_BLOCKS", "JMC_FAST_PAIRS", "JMC_FAST_PAIRS_2W", "JMC_FAST_PIN_KEYS", "JMC_PHILOX_SPLIT", "JMC_ACC_MODE",
            "JMC_HIST_MODE", "JMC_HIST_UNROLL", "JMC_SHOWER_GROUP", "JMC_SHOWER_BIN_GROUP", "JMC_SHOWER_CHUNK",
            "JMC_SHOWER_MIN_BLOCKS", "JMC_SHOWER_TRIALS", "JMC_SHOWER_SHARE")
_DEFS = ["-D%s=%s" % (k, os.environ[k]) for k in SWITCHES if os.environ.get(k)]
UNITS = {
    "jmc_exact.cu": ["--fmad=false"] + _DEFS,
    "jmc_shower.cu": ["--fmad=false"] + _DEFS,
    "jmc_process.cu": ["--fmad=false"] + _DEFS,
    "jmc_sudakov.cu": ["--fmad=false"] + _DEFS,
    "jmc_fast.cu": ["--fmad=true"] + _DEFS,
    "jmc_host.cu": [],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    names = sorted(f for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".py")))
    for k in SWITCHES:
        h.update(os.environ.get(k, "").encode())
    # names are hashed relative to the repository so that the stamp survives a copy of the tree (the GPU box
    # receives the built library with the snapshot and must load it as is)
    files = [(name, os.path.join(HERE, name)) for name in names]
    files.append(("include/jmc_b200.h", os.path.join(ROOT, "include", "jmc_b200.h")))
    for name, path in files:
        with open(path, "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed; returns the path of the shared library."""
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "libjmc_b200.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    units = {u: f for u, f in UNITS.items() if os.path.exists(os.path.join(HERE, u))}

    # experiments that touch one translation unit: JMC_BUILD_ONLY=jmc_shower.cu reuses the other objects
    only = [u for u in os.environ.get("JMC_BUILD_ONLY", "").split(",") if u]

    def compile_one(item):
        unit, flags = item
        obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
        if only and unit not in only and os.path.exists(obj):
            return obj
        cmd = [nvcc, *ARCH, *COMMON, *flags, *(["-Xptxas", "-v"] if verbose else []),
               "-c", os.path.join(HERE, unit), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, res.stdout, res.stderr))
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(units)) as pool:
        objs = list(pool.map(compile_one, units.items()))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
