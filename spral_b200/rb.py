"""Rutherford-Boeing files (assembled matrices), enough of spral_rutherford_boeing
(src/rutherford_boeing.f90) to feed real test matrices to the engine and to write the
benchmark matrices out: `rb_read` mirrors rb_read (:217-470) for types [r|i|p][s|u|z|r]a --
a symmetric matrix comes back as its lower triangle in 1-based CSC, which is what
ssids_analyse takes -- and `rb_write` mirrors rb_write (:640-731: header lines
"(a72,a8)", "(i14,3(1x,i13))", "(a3,11x,i14,3(1x,i13))", "(a16,a16,a20)", default value
format "(3e24.16)").  Elemental matrices and the supplementary right-hand sides of the
format are not handled."""
import re

import numpy as np


def _parse_format(fmt):
    """'(10i8)', '(3e24.16)', '(1p,4d20.12)' -> (items per line, field width)."""
    f = fmt.strip().lower().replace(" ", "")
    f = re.sub(r"\d+p,?", "", f)                        # scale factors do not change the layout
    m = re.search(r"\(?(\d*)\s*([ifedg])\s*(\d+)", f)
    if not m:
        raise ValueError(f"unsupported Fortran format {fmt!r}")
    return int(m.group(1) or 1), int(m.group(3))


def _read_fixed(lines, pos, count, fmt, conv):
    per, width = _parse_format(fmt)
    out = []
    while len(out) < count:
        line = lines[pos].rstrip("\n")
        pos += 1
        for k in range(per):
            if len(out) == count:
                break
            field = line[k * width:(k + 1) * width]
            if not field.strip():
                break
            out.append(conv(field))
    return out, pos


def _to_float(field):
    return float(field.strip().lower().replace("d", "e"))


def rb_peek(path):
    """Header information as rb_peek (:74-215): dict with title, id, type_code, m, n, nnz."""
    with open(path) as fh:
        l1, l2, l3 = fh.readline(), fh.readline(), fh.readline()
    t = l3[:3].lower()
    m, n, nnz, nelt = (int(x) for x in l3[3:].split()[:4])
    return {"title": l1[:72].rstrip(), "id": l1[72:80].strip(), "type_code": t, "m": m, "n": n, "nnz": nnz,
            "lines": [int(x) for x in l2.split()[:4]]}


def rb_read(path):
    """(n, ptr, row, val, info): 1-based int64 ptr, int32 row, float64 val (ones for pattern
    matrices).  Symmetric / skew / Hermitian files hold one triangle; entries above the diagonal are
    mirrored so that the result is the LOWER triangle, columns sorted by row."""
    info = rb_peek(path)
    t = info["type_code"]
    if len(t) != 3 or t[2] != "a" or t[0] not in "rip":
        raise ValueError(f"unsupported Rutherford-Boeing type {t!r} (assembled real / integer / pattern only)")
    with open(path) as fh:
        lines = fh.readlines()
    ptrfmt, indfmt, valfmt = lines[3][:16], lines[3][16:32], lines[3][32:52]
    n, nnz = info["n"], info["nnz"]
    pos = 4
    ptr, pos = _read_fixed(lines, pos, n + 1, ptrfmt, int)
    row, pos = _read_fixed(lines, pos, nnz, indfmt, int)
    if t[0] == "p":
        val = [1.0] * nnz
    else:
        val, pos = _read_fixed(lines, pos, nnz, valfmt, _to_float if t[0] == "r" else lambda s: float(int(s)))
    ptr, row, val = np.asarray(ptr, np.int64), np.asarray(row, np.int32), np.asarray(val, np.float64)
    if t[1] in "szh":                                   # one triangle stored: normalise to the lower one
        import scipy.sparse as sp
        col = np.repeat(np.arange(1, n + 1), np.diff(ptr))
        lo_r, lo_c = np.maximum(row, col), np.minimum(row, col)
        A = sp.coo_matrix((val, (lo_r - 1, lo_c - 1)), shape=(n, n)).tocsc()
        A.sort_indices()
        ptr, row, val = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int32) + 1, A.data.astype(np.float64)
    return n, ptr, row, val, info


def rb_write(path, n, ptr, row, val=None, title="", ident="0", type_code=None, val_format="(3e24.16)"):
    """Writes a square matrix given in 1-based CSC (the lower triangle for the symmetric types) as
    rb_write (:640-731) does; type_code defaults to 'rsa' with values, 'psa' without."""
    ptr = np.asarray(ptr, np.int64)
    row = np.asarray(row, np.int32)
    nnz = int(ptr[n]) - 1
    t = type_code or ("rsa" if val is not None else "psa")

    def int_format(maxval):
        prec = len(str(int(maxval))) + 1
        return max(1, 80 // prec), prec

    pp, pw = int_format(ptr[n])
    rp, rw = int_format(max(n, 1))
    vper, vwidth = _parse_format(val_format)
    vdig = int(re.search(r"\.(\d+)", val_format).group(1))
    ptrcrd = -(-(n + 1) // pp)
    indcrd = -(-nnz // rp)
    valcrd = -(-nnz // vper) if val is not None else 0
    with open(path, "w") as fh:
        fh.write(f"{title[:72]:<72}{ident[:8]:<8}\n")
        fh.write(f"{ptrcrd + indcrd + valcrd:14d} {ptrcrd:13d} {indcrd:13d} {valcrd:13d}\n")
        fh.write(f"{t:3}{'':11}{n:14d} {n:13d} {nnz:13d} {0:13d}\n")
        fh.write(f"{f'({pp}i{pw})':<16}{f'({rp}i{rw})':<16}{val_format:<20}\n")
        for data, per, width in ((ptr[:n + 1], pp, pw), (row[:nnz], rp, rw)):
            for k in range(0, len(data), per):
                fh.write("".join(f"{int(x):{width}d}" for x in data[k:k + per]) + "\n")
        if val is not None:
            v = np.asarray(val, np.float64)[:nnz]
            for k in range(0, nnz, vper):
                fh.write("".join(f"{x:{vwidth}.{vdig}E}" for x in v[k:k + vper]) + "\n")
