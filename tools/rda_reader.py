"""Minimal reader for R `.rda` files (bz2/gzip/xz + RDX2/RDX3 XDR serialisation).

Used once, in the build container, to decode the reference's bundled datasets
(/root/reference/data/*.rda, documented at /root/reference/R/data.R:1-65) into the small
fixtures committed under tests/golden/ (see tools/make_fixtures.py).  Not used at run time.
"""
import bz2, gzip, lzma, struct


class _R:
    def __init__(self, b):
        self.b, self.p, self.refs = b, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]; self.p += 4; return v

    def f64(self):
        v = struct.unpack_from(">d", self.b, self.p)[0]; self.p += 8; return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254: return None                      # NILVALUE
        if t == 253: return "<globalenv>"
        if t == 255: return self.refs[(flags >> 8) - 1] if (flags >> 8) else self.refs[self.i32() - 1]
        if t == 1:                                    # SYMSXP
            s = self.item(); self.refs.append(s); return s
        if t == 2:                                    # LISTSXP (pairlist)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32(); t = flags & 0xFF
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
                if t == 254: break
                assert t == 2, t
            return out
        if t == 9:                                    # CHARSXP
            n = self.i32()
            if n == -1: return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace"); self.p += n; return s
        if t in (10, 13):                             # LGLSXP / INTSXP
            n = self.i32(); v = [self.i32() for _ in range(n)]
            v = [None if x == -2147483648 else x for x in v]
        elif t == 14:                                 # REALSXP
            n = self.i32(); v = list(struct.unpack_from(">%dd" % n, self.b, self.p)); self.p += 8 * n
        elif t == 16:                                 # STRSXP
            n = self.i32(); v = [self.item() for _ in range(n)]
        elif t == 19:                                 # VECSXP
            n = self.i32(); v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError("SEXP type %d at %d" % (t, self.p))
        attrs = dict(self.item()) if has_attr else {}
        if t == 19 and "names" in attrs:
            d = dict(zip(attrs["names"], v)); d["__attrs__"] = attrs; return d
        if t == 13 and "levels" in attrs:
            return [attrs["levels"][x - 1] if x is not None else None for x in v]
        return v


def read_rda(path):
    raw = open(path, "rb").read()
    for dec in (bz2.decompress, gzip.decompress, lzma.decompress, lambda x: x):
        try:
            b = dec(raw); break
        except Exception:
            continue
    assert b[:5] in (b"RDX2\n", b"RDX3\n"), b[:5]
    r = _R(b); r.p = 5
    assert b[r.p:r.p + 2] == b"X\n"; r.p += 2
    ver = r.i32(); r.i32(); r.i32()
    if ver == 3:
        n = r.i32(); r.p += n
    return dict((k, v) for k, v in r.item())
