"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ramsesgpu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(native, s), "missing export %s" % s
    from ramsesgpu_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes table covers exactly the header


def test_no_device_means_error_not_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from ramsesgpu_b200 import MHDRunGodunov, _lib
    with pytest.raises(_lib.RgError) as e:
        MHDRunGodunov("[mesh]\nnx=8\nny=8\nnz=8\n[MHD]\nenable=true\n")
    assert e.value.code == _lib.RG_ERR_NO_DEVICE


def test_slab_extent_covers_domain(native):
    from ramsesgpu_b200 import slab_extent
    for nz, n in ((256, 4), (1024, 8), (10, 3), (7, 7)):
        ext = [slab_extent(nz, n, r) for r in range(n)]
        assert ext[0][1] == 0
        for (a, o), (b, p) in zip(ext, ext[1:]):
            assert o + a == p
        assert ext[-1][1] + ext[-1][0] == nz
        assert max(e[0] for e in ext) - min(e[0] for e in ext) <= 1


def test_sass_is_sm100(native):
    """the shipped .so carries sm_100a SASS (cuobjdump) -- skipped when the CUDA toolkit is absent"""
    import shutil
    import subprocess
    from ramsesgpu_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("no cuobjdump")
    out = subprocess.run([exe, "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    assert "sm_100a" in out


def test_header_is_plain_c_and_links(native, tmp_path):
    """the boundary is a C ABI: the public header compiles as C99 (no C++/torch types) and a C caller links
    against the shared library; without a device the program gets RG_ERR_NO_DEVICE, computes the initial
    condition on the host and exits cleanly"""
    import shutil
    import subprocess
    from ramsesgpu_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "ramsesgpu_b200.h"
int main(void) {
  const char* ini = "[mesh]\nnx=8\nny=6\nnz=4\nboundary_xmin=3\nboundary_xmax=3\nboundary_ymin=3\nboundary_ymax=3\n"
                    "boundary_zmin=3\nboundary_zmax=3\n[hydro]\nproblem=Orszag-Tang\ngamma0=1.66\n[MHD]\nenable=true\n";
  rg_layout L;
  if (rg_initial_condition_host(ini, 0, 0, 1, NULL, 0, &L) != RG_OK) { puts(rg_last_error()); return 2; }
  size_t n = (size_t)L.isize * L.jsize * L.ksize * L.nvar;
  double* U = (double*)malloc(n * sizeof(double));
  if (rg_initial_condition_host(ini, 0, 0, 1, U, n * sizeof(double), &L) != RG_OK) { puts(rg_last_error()); return 3; }
  printf("layout %d %d %d nvar %d gw %d rho %.17g\n", L.isize, L.jsize, L.ksize, L.nvar, L.ghost_width,
         U[(size_t)L.ghost_width * L.isize * L.jsize + (size_t)L.ghost_width * L.isize + L.ghost_width]);
  rg_handle h = NULL;
  int rc = rg_create(ini, 0, &h);
  printf("rg_create -> %d (%s)\n", rc, rc ? rg_last_error() : "ok");
  if (rc == RG_OK) rg_destroy(h);
  free(U);
  return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-lramsesgpu_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
    text = out.stdout.decode()
    assert out.returncode == 0, text
    assert "layout 14 12 10 nvar 8 gw 3" in text
    assert "rho 0.21928367177348035" in text          # (1.66f)^2 / 4 pi, SURVEY 8(c)
    import torch
    if not torch.cuda.is_available():
        assert "rg_create -> 2" in text               # RG_ERR_NO_DEVICE, no CPU fallback


def test_struct_layouts_match_the_header(native, tmp_path):
    """rg_layout / rg_stats are filled by the library and read through ctypes mirrors: same size and the same field
    offsets as the C compiler gives the header's structs (a field added on one side only would corrupt memory silently)"""
    import shutil
    import subprocess
    from ramsesgpu_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    text = open(os.path.join(ROOT, "include", "ramsesgpu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "ramsesgpu_b200.h"', 'int main(void) {']
    mirrors = {"rg_layout": _lib.RgLayout, "rg_stats": _lib.RgStats}
    for name in mirrors:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                first, *rest = decl.split(",")
                fields += [first.split()[-1]] + [r.strip() for r in rest]
        assert fields == [f[0] for f in mirrors[name]._fields_], name
        prog.append('  printf("%s %%zu", sizeof(%s));' % (name, name))
        prog += ['  printf(" %%zu", offsetof(%s, %s));' % (name, f) for f in fields]
        prog.append('  printf("\\n");')
    prog += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()
    for line in subprocess.run([str(exe)], stdout=subprocess.PIPE).stdout.decode().splitlines():
        name, size, *offs = line.split()
        m = mirrors[name]
        assert int(size) == C.sizeof(m), name
        assert [int(o) for o in offs] == [getattr(m, f[0]).offset for f in m._fields_], name


def test_sanitizer_cases_are_written(tmp_path):
    """tools/sanitize_cases.py (parameter files of the compute-sanitizer pass) still finds its fixtures"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_cases.py"), str(tmp_path)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert r.returncode == 0, r.stdout.decode()
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".ini"))
    assert len(files) >= 13 and any(f.endswith("_f32.ini") for f in files)
    assert "nstepmax=3" in open(tmp_path / "ot3d_16_s10.ini").read()
