"""TEST INFRASTRUCTURE ONLY (oracle/): minimal reader for R's RDX3 `.rda` files.

Used once, in this container, by oracle/make_golden.py to turn the reference's
datasets (/root/reference/data/holes.rda, stripes.rda, holes_bm.rda;
documented at R/data.R:1-55) into the small fixtures under tests/golden/.
Handles what those files contain: gzip'd XDR serialisation version 3 with
REALSXP / INTSXP / LGLSXP / STRSXP / VECSXP / pairlist attributes / symbol
references / compact-sequence ALTREP.  Data frames come back as dicts of numpy
columns, lists as python lists (or dicts when named).
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

_NIL, _REF, _ALTREP = 254, 255, 238
_SYM, _LIST, _CHAR, _LGL, _INT, _REAL, _STR, _VEC = 1, 2, 9, 10, 13, 14, 16, 19
_LANG, _GLOBALENV, _EMPTYENV, _BASEENV, _NAMESPACE, _PACKAGE = 6, 253, 242, 241, 249, 250


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def int(self):
        (v,) = struct.unpack_from(">i", self.b, self.o)
        self.o += 4
        return v

    def length(self):
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def bytes(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == _NIL:
            return None
        if t in (_GLOBALENV, _EMPTYENV, _BASEENV):
            return None
        if t == _REF:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t == _SYM:
            name = self.item()
            self.refs.append(name)
            return name
        if t in (_NAMESPACE, _PACKAGE):
            self.int()
            n = self.int()
            v = [self.item() for _ in range(n)]
            self.refs.append(v)
            return v
        if t == _CHAR:
            n = self.int()
            return None if n == -1 else self.bytes(n).decode("utf-8", "replace")
        if t in (_LIST, _LANG):
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                t2 = flags & 0xFF
                if t2 == _NIL:
                    break
                if t2 not in (_LIST, _LANG):
                    self.o -= 4
                    out.append((None, self.item()))
                    break
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if t == _ALTREP:
            info = self.item()
            state = self.item()
            self.item()  # attributes
            cls = info[0][1] if info else ""
            if cls == "compact_intseq":
                n, start, step = (int(x) for x in np.asarray(state))
                return np.arange(start, start + n * step, step, dtype=np.int64)
            if cls == "compact_realseq":
                n, start, step = np.asarray(state)
                return start + step * np.arange(int(n), dtype=np.float64)
            return state
        if t in (_LGL, _INT):
            n = self.length()
            v = np.frombuffer(self.bytes(4 * n), dtype=">i4").astype(np.int64)
        elif t == _REAL:
            n = self.length()
            v = np.frombuffer(self.bytes(8 * n), dtype=">f8").astype(np.float64)
        elif t == _STR:
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == _VEC:
            n = self.length()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError("SEXP type %d at offset %d" % (t, self.o))
        attrs = dict((k, a) for k, a in self.item()) if has_attr else {}
        return _finish(v, attrs, t)


def _finish(v, attrs, t):
    names = attrs.get("names")
    cls = attrs.get("class")
    if t == _VEC and cls and "data.frame" in cls:
        return {nm: col for nm, col in zip(names, v)}
    if t == _VEC and names:
        return {nm if nm else str(k): col for k, (nm, col) in enumerate(zip(names, v))}
    if "dim" in attrs and isinstance(v, np.ndarray):
        dim = tuple(int(d) for d in attrs["dim"])
        return v.reshape(dim, order="F")
    return v


def read_rda(path):
    """Return {object name: value} for every object saved in `path`."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:5] != b"RDX3\n" or raw[5:7] != b"X\n":
        raise ValueError("not an XDR RDX3 file: %r" % raw[:7])
    r = _Reader(raw)
    r.o = 7
    version, _, _ = r.int(), r.int(), r.int()
    if version == 3:
        r.bytes(r.int())  # native encoding
    top = r.item()
    return {tag: val for tag, val in top}


def frame_to_matrix(frame, columns=None):
    """Stack data-frame columns (dict from read_rda) into an n x k float64 matrix."""
    cols = list(frame.keys()) if columns is None else columns
    return np.column_stack([np.asarray(frame[c], dtype=np.float64) for c in cols]), cols
