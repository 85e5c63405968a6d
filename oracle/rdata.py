"""ORACLE / TEST INFRASTRUCTURE — not product code.

Minimal reader for R's ``save()`` format (bzip2/gzip/xz + ``RDX3`` XDR serialisation), just
enough to decode the reference's packaged fixtures ``data/BD_Obs.rda`` and ``data/BD_Coord.rda``
(data.table objects: VECSXP of REALSXP/STRSXP columns with names/class/row.names attributes and a
``.internal.selfref`` external pointer).  R is not installed in the build container, so the
fixtures are decoded by hand once by ``tests/golden/make_golden.py`` and frozen as ``.npz``.

Only ``tests/`` and the golden generator may import this module.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

NA_INTEGER = -(2 ** 31)


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64(self) -> float:
        v = struct.unpack_from(">d", self.b, self.o)[0]
        self.o += 8
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


class RObject:
    """A decoded R value with its attribute pairlist (dict)."""

    def __init__(self, kind, value, attrs=None):
        self.kind, self.value, self.attrs = kind, value, attrs or {}

    def __repr__(self):
        return f"RObject({self.kind}, n={len(self.value) if hasattr(self.value, '__len__') else self.value})"


def _read_item(r: _Reader):
    flags = r.i32()
    typ = flags & 0xFF
    has_attr = bool(flags & (1 << 9))
    has_tag = bool(flags & (1 << 10))

    if typ == 254:  # NILVALUE_SXP
        return None
    if typ == 255:  # REFSXP
        idx = flags >> 8
        if idx == 0:
            idx = r.i32()
        return r.refs[idx - 1]
    if typ == 253:  # GLOBALENV
        return RObject("env", "global")
    if typ == 242:  # EMPTYENV
        return RObject("env", "empty")
    if typ == 1:  # SYMSXP
        name = _read_item(r)
        sym = RObject("sym", name.value if isinstance(name, RObject) else name)
        r.refs.append(sym)
        return sym
    if typ in (2, 6):  # LISTSXP / LANGSXP pairlist -> python list of (tag, value)
        out = []
        attrs = None
        while True:
            if has_attr:
                attrs = _read_item(r)
            tag = _read_item(r) if has_tag else None
            car = _read_item(r)
            out.append((tag.value if isinstance(tag, RObject) else None, car))
            # CDR: peek flags of the next cell
            nflags = struct.unpack_from(">i", r.b, r.o)[0]
            ntyp = nflags & 0xFF
            if ntyp == 254:
                r.o += 4
                break
            if ntyp not in (2, 6):
                cdr = _read_item(r)
                out.append((None, cdr))
                break
            r.o += 4
            has_attr = bool(nflags & (1 << 9))
            has_tag = bool(nflags & (1 << 10))
        return RObject("pairlist", out)
    if typ == 9:  # CHARSXP
        n = r.i32()
        if n == -1:
            return RObject("char", None)
        return RObject("char", r.raw(n).decode("utf-8", "replace"))
    if typ == 10:  # LGLSXP
        n = r.i32()
        v = [r.i32() for _ in range(n)]
        obj = RObject("lgl", v)
    elif typ == 13:  # INTSXP
        n = r.i32()
        v = [r.i32() for _ in range(n)]
        obj = RObject("int", v)
    elif typ == 14:  # REALSXP
        n = r.i32()
        raw = r.raw(8 * n)
        obj = RObject("real", raw)  # keep the XDR bytes so NA payloads survive
    elif typ == 16:  # STRSXP
        n = r.i32()
        v = [_read_item(r).value for _ in range(n)]
        obj = RObject("str", v)
    elif typ == 19:  # VECSXP
        n = r.i32()
        v = [_read_item(r) for _ in range(n)]
        obj = RObject("list", v)
    elif typ == 22:  # EXTPTRSXP
        obj = RObject("extptr", None)
        r.refs.append(obj)
        _read_item(r)  # prot
        _read_item(r)  # tag
    elif typ == 24:  # RAWSXP
        n = r.i32()
        obj = RObject("raw", r.raw(n))
    else:
        raise NotImplementedError(f"SEXP type {typ} at offset {r.o}")
    if has_attr:
        a = _read_item(r)
        if a is not None:
            obj.attrs = {k: v for k, v in a.value}
    return obj


def load_rda(path: str) -> dict:
    """Return {object name: RObject} for every object stored by save()."""
    with open(path, "rb") as f:
        head = f.read(6)
        f.seek(0)
        blob = f.read()
    if head[:3] == b"BZh":
        blob = bz2.decompress(blob)
    elif head[:2] == b"\x1f\x8b":
        blob = gzip.decompress(blob)
    elif head[:6] == b"\xfd7zXZ\x00":
        blob = lzma.decompress(blob)
    if blob[:5] != b"RDX3\n":
        raise ValueError(f"not an RDX3 file: {blob[:5]!r}")
    r = _Reader(blob)
    r.o = 5
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    r.i32()  # serialisation version (3)
    r.i32()  # writer R version
    r.i32()  # min reader R version
    n = r.i32()
    r.raw(n)  # native encoding
    top = _read_item(r)
    return {k: v for k, v in top.value}


def real_vector(obj: RObject):
    """REALSXP payload as float64 numpy array (bit-exact, NA_real_ payload preserved)."""
    import numpy as np

    return np.frombuffer(obj.value, dtype=">f8").astype("<f8")


def data_frame(obj: RObject) -> dict:
    """data.frame / data.table -> {column name: numpy array or list of str}."""
    names = obj.attrs["names"].value
    out = {}
    for nm, col in zip(names, obj.value):
        if col.kind == "real":
            out[nm] = real_vector(col)
            out[nm + "::class"] = col.attrs.get("class").value if "class" in col.attrs else None
        elif col.kind == "str":
            out[nm] = list(col.value)
        elif col.kind == "int":
            if "levels" in col.attrs:  # factor
                lev = col.attrs["levels"].value
                out[nm] = [lev[i - 1] if i != NA_INTEGER else None for i in col.value]
            else:
                out[nm] = list(col.value)
        else:
            out[nm] = col
    return out
