"""In-tree build of the native library: ramsesgpu_b200/lib/libramsesgpu_b200.so (+ the
ramsesgpu_b200_main executable).  nvcc cross-compiles sm_100a without a GPU.

    python -m ramsesgpu_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# RG_VARIANT=<tag> builds an experimental flavour next to the product (lib_<tag>/, build_<tag>/) with the extra
# preprocessor defines of RG_VARIANT_DEFINES (e.g. "-DRG_EXP_INT_CLAMP -DRG_EXP_LIMITER_V1"); load it with
# RG_LIB_PATH=ramsesgpu_b200/lib_<tag>/libramsesgpu_b200.so for A/B timing.  Unset: the product, as always.
VARIANT = os.environ.get("RG_VARIANT", "")
LIBDIR = os.path.join(PKG, "lib" + ("_" + VARIANT if VARIANT else ""))
OBJDIR = os.path.join(PKG, "build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(LIBDIR, "libramsesgpu_b200.so")
MAIN = os.path.join(LIBDIR, "ramsesgpu_b200_main")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
CUFLAGS = ["--expt-relaxed-constexpr", "-Xptxas", "-v"]
if VARIANT:
    COMMON = COMMON + os.environ.get("RG_VARIANT_DEFINES", "").split()


def _sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f == "main.cpp":
            continue
        if f.endswith(".cu") or f.endswith(".cpp"):
            out.append(os.path.join(CSRC, f))
    return out


def _headers_digest():
    h = hashlib.sha1()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh", ".hpp")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(ARCH + COMMON + CUFLAGS).encode())
    return h.hexdigest()


def _compile(src, stamp, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src) + ".o")
    tag = obj + ".stamp"
    with open(src, "rb") as fh:
        want = hashlib.sha1(fh.read()).hexdigest() + stamp
    if os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == want:
        return obj, ""
    cmd = [NVCC] + ARCH + COMMON + (CUFLAGS if src.endswith(".cu") else ["-x", "cu"] + CUFLAGS[:1]) + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    log = p.stdout.decode()
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
    with open(tag, "w") as fh:
        fh.write(want)
    if verbose:
        print(" ".join(cmd))
    return obj, log


def build(force=False, verbose=False):
    """Builds what is out of date.  Safe to call from several processes at once (pytest-xdist workers, the ranks of a
    torchrun launch): an exclusive file lock serialises them, the late comers find everything up to date."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    with open(os.path.join(OBJDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    stamp = _headers_digest() + ("force%d" % os.getpid() if force else "")
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, stamp, verbose), srcs))
    objs = [o for o, _ in results]
    logs = "".join(l for _, l in results)
    if logs:
        with open(os.path.join(OBJDIR, "ptxas.log"), "w") as fh:
            fh.write(logs)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB + ".tmp"] + objs + ["-lcudart", "-ldl"])
        os.replace(LIB + ".tmp", LIB)
    main_src = os.path.join(CSRC, "main.cpp")
    if os.path.exists(main_src) and (force or not os.path.exists(MAIN) or os.path.getmtime(MAIN) < max(newest, os.path.getmtime(main_src))):
        subprocess.check_call([NVCC] + ARCH + COMMON + [main_src, "-o", MAIN + ".tmp", "-L", LIBDIR, "-lramsesgpu_b200",
                                                         "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
        os.replace(MAIN + ".tmp", MAIN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
