"""Builds the oracle (TEST INFRASTRUCTURE ONLY):
  oracle/liboracle_f64.so, oracle/liboracle_f32.so  -- the C restatement (always)
  oracle/_ref/euler_cpu[_f32]                        -- the unmodified reference CPU executable,
                                                        only when /root/reference is present
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RAMSES_REFERENCE", "/root/reference")
SRCS = sorted(f for f in os.listdir(HERE) if f.startswith("oracle_") and f.endswith(".c"))
# -ffp-contract=off: same arithmetic as the reference's g++ -O3 x86-64 build (no FMA contraction)
CFLAGS = ["-O2", "-std=gnu99", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall",
          "-Wno-unused-variable", "-Wno-unused-function"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_restatement(verbose=False):
    deps = [os.path.join(HERE, s) for s in SRCS] + [os.path.join(HERE, "oracle.h")]
    out = []
    for name, defs in (("liboracle_f64.so", []), ("liboracle_f32.so", ["-DORACLE_FLOAT"])):
        target = os.path.join(HERE, name)
        if _stale(target, deps):
            cmd = ["gcc"] + CFLAGS + defs + [os.path.join(HERE, s) for s in SRCS] + ["-o", target, "-lm"]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        out.append(target)
    return out


def build_reference(verbose=False):
    """g++ on the reference's own sources where they lie (oracle/Makefile.ref); skipped when the
    reference tree is absent (GPU box: the prebuilt binaries travel with the snapshot)."""
    if not os.path.isdir(os.path.join(REF, "src", "hydro")):
        return None
    for tgt in ("all", "f32"):
        cmd = ["make", "-s", "-f", os.path.join(HERE, "Makefile.ref"), "-j", str(os.cpu_count() or 4),
               "REF=" + REF, "OUT=" + os.path.join(HERE, "_ref"), tgt]
        subprocess.check_call(cmd, cwd=ROOT, stdout=None if verbose else subprocess.DEVNULL)
    return os.path.join(HERE, "_ref", "euler_cpu")


if __name__ == "__main__":
    print(build_restatement(verbose=True))
    if "--no-ref" not in sys.argv:
        print(build_reference(verbose=True))
