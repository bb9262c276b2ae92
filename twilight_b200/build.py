"""Builds libtwilight_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libtwilight_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # The DP must reproduce the reference's float results bit for bit: nothing may be contracted implicitly.
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "twilight_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libtwilight_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    # build stamp: which sources (sha1), which compiler, when — so a "does it build" check leaves evidence that it compiled
    import hashlib
    import json
    import time
    h = hashlib.sha1()
    for path in sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(CSRC, "*.hpp"))):
        h.update(open(path, "rb").read())
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1:]
    with open(os.path.join(PKG, "BUILD_INFO.json"), "w") as f:
        json.dump({"library": os.path.basename(LIB), "built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "nvcc": ver, "flags": NVCC_FLAGS,
                   "sources": [os.path.basename(p) for p in sources()], "sources_sha1": h.hexdigest(), "bytes": os.path.getsize(LIB)}, f, indent=1)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
