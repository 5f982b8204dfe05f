"""Readers for the two reference output formats the parity checks use, and the L2-relative norm.

* ``.xsm``  : one ASCII header line ``Binary 1 NXxNY[xNZ] N(B byte reals)`` followed by the raw
  inner cells of ONE variable (reference HydroRunBase.cpp:2520-2562).
* ``.vti``  : hand-written raw-appended VTK ImageData; per variable a ``uint32`` byte count and the
  raw inner cells, x fastest (reference HydroRunBase.cpp:2974-3033).
* ``l2_relative`` : ``sqrt(sum((a-b)^2)/sum(a^2))`` -- Python-3/NumPy restatement of
  test/computeL2relatif.py.in:43-50, extended to 3D and to every variable.
"""
import re

import numpy as np

VAR_NAMES_MHD = ["density", "energy", "mx", "my", "mz", "bx", "by", "bz"]


def read_xsm(path):
    with open(path, "rb") as f:
        header = f.readline().decode()
        raw = f.read()
    dims = [int(x) for x in header.split()[2].split("x")]
    nbytes = int(re.search(r"\((\d+) byte", header).group(1))
    dtype = np.float64 if nbytes == 8 else np.float32
    n = int(np.prod(dims))
    return np.frombuffer(raw, dtype=dtype, count=n).reshape(dims[::-1]).copy()


def read_vti(path):
    """Returns {name: ndarray[(nz,) ny, nx]} of the inner cells of every variable."""
    with open(path, "rb") as f:
        blob = f.read()
    head_end = blob.index(b"<AppendedData")
    head = blob[:head_end].decode()
    ext = [int(x) for x in re.search(r'WholeExtent="([^"]+)"', head).group(1).split()]
    nx, ny, nz = ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1
    arrays = re.findall(r'<DataArray type="(Float32|Float64)" Name="([^"]+)" format="appended" offset="(\d+)"', head)
    start = blob.index(b"_", head_end) + 1
    out = {}
    for typ, name, off in arrays:
        dtype = np.float64 if typ == "Float64" else np.float32
        o = start + int(off)
        nbytes = int(np.frombuffer(blob, dtype=np.uint32, count=1, offset=o)[0])
        a = np.frombuffer(blob, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize, offset=o + 4)
        out[name] = a.reshape((nz, ny, nx) if nz > 1 else (ny, nx)).copy()
    return out


def l2_relative(ref, other):
    ref = np.asarray(ref, dtype=np.float64)
    other = np.asarray(other, dtype=np.float64)
    den = float(np.sum(ref * ref))
    num = float(np.sum((ref - other) ** 2))
    if den == 0.0:
        return 0.0 if num == 0.0 else float(np.sqrt(num))
    return float(np.sqrt(num / den))


def ini_override(text, overrides):
    """Returns ini ``text`` with ``overrides = {section: {key: value}}`` applied (keys replaced in place
    or appended to their section; sections created when missing).  Section/key match is
    case-insensitive like the reference's INIReader::makeKey."""
    lines = text.splitlines()
    todo = {s.lower(): {k.lower(): (k, v) for k, v in kv.items()} for s, kv in overrides.items()}
    out, sec = [], None
    last_line_of_section = {}
    for ln in lines:
        st = ln.strip()
        if st.startswith("[") and "]" in st:
            sec = st[1:st.index("]")].strip().lower()
        elif "=" in st and not st.startswith(("#", ";")) and sec in todo:
            key = st.split("=", 1)[0].strip().lower()
            if key in todo[sec]:
                k, v = todo[sec].pop(key)
                ln = "%s=%s" % (k, v)
        out.append(ln)
        if sec is not None:
            last_line_of_section[sec] = len(out)
    for sec, kv in todo.items():
        if not kv:
            continue
        new = ["%s=%s" % (k, v) for k, v in kv.values()]
        if sec in last_line_of_section:
            at = last_line_of_section[sec]
            out[at:at] = new
            for s2 in last_line_of_section:
                if last_line_of_section[s2] >= at and s2 != sec:
                    last_line_of_section[s2] += len(new)
        else:
            orig = [s for s in overrides if s.lower() == sec][0]
            out += ["", "[%s]" % orig] + new
    return "\n".join(out) + "\n"
