"""Reader for R's XDR serialization (version 2): `.RData` files (gzip, RDX2) and knitr lazy-load caches (`.rdb` / `.rdx`).

The reference ships its genealogies as R objects (`data/Ebola_Sierra_Leone2014.RData`: a `phylo` tree) and its users keep
theirs in `.RData` files; this module reads them without R, so that `read_phylo(path)` -> `genealogy.summarize_phylo` ->
`mcmc.General_MCMC_with_ESlice` is the whole path from the reference's data file to chains on the GPU (SURVEY.md section 8(f.2),
Appendix D).  Covers what those files contain: NULL, symbols, pairlists, character / logical / integer / real vectors,
lists and their names / dim / class attributes; no closures, environments, ALTREP or bytecode.  Host-side NumPy only.
tests/golden/make_golden.py uses the same reader to turn the reference's knitr caches into the .npz fixtures.
"""
import gzip
import struct
import zlib

import numpy as np

NILVALUE, REFSXP = 254, 255


class RObj:
    """value + attributes (names/dim/class) of one R object."""

    def __init__(self, value, attr=None):
        self.value = value
        self.attr = attr or {}

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attr={list(self.attr)})"


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def header(self):
        assert self.b[self.o:self.o + 2] == b"X\n", self.b[self.o:self.o + 8]
        self.o += 2
        version = self.i32()
        self.i32()
        self.i32()
        assert version == 2, version

    def charsxp(self):
        flags = self.i32()
        assert flags & 0xFF == 9, flags
        n = self.i32()
        if n == -1:
            return None
        s = self.b[self.o:self.o + n].decode("latin-1")
        self.o += n
        return s

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE:
            return None
        if t == REFSXP:
            return self.refs[(flags >> 8) - 1]
        if t == 1:  # SYMSXP
            s = self.charsxp()
            self.refs.append(s)
            return s
        if t == 2:  # LISTSXP (pairlist) -> list of (tag, value); iterative over cdr
            out = []
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
                if t == NILVALUE:
                    break
                assert t == 2, t
            return out
        if t == 9:
            self.o -= 4
            return self.charsxp()
        if t in (10, 13):
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype(np.int32)
            self.o += 4 * n
        elif t == 14:
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype(np.float64)
            self.o += 8 * n
        elif t == 16:
            n = self.i32()
            v = [self.charsxp() for _ in range(n)]
        elif t == 19:
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.o}")
        attr = {}
        if has_attr:
            for tag, val in self.item():
                attr[tag] = val
        return RObj(v, attr)


def simplify(x):
    """RObj tree -> dicts / lists / numpy arrays (column-major dims applied)."""
    if isinstance(x, RObj):
        v = x.value
        names = x.attr.get("names")
        dim = x.attr.get("dim")
        if isinstance(v, np.ndarray):
            if dim is not None:
                return v.reshape(tuple(int(d) for d in dim.value), order="F")
            return v
        if isinstance(v, list) and v and all(isinstance(e, (str, type(None))) for e in v):
            return v
        if isinstance(v, list):
            vals = [simplify(e) for e in v]
            if names is not None:
                nm = names.value
                return {(n if n else f"_{i}"): e for i, (n, e) in enumerate(zip(nm, vals))}
            return vals
        return v
    if isinstance(x, list):
        return [(t, simplify(v)) for t, v in x]
    return x


def read_rdb(rdb_path, rdx_path):
    """knitr lazy-load cache: returns {variable name: object}."""
    with open(rdx_path, "rb") as f:
        r = _Reader(gzip.decompress(f.read()))
    r.header()
    idx = simplify(r.item())
    with open(rdb_path, "rb") as f:
        blob = f.read()
    out = {}
    for name, (off, length) in ((k, v.astype(int)) for k, v in idx["variables"].items()):
        rec = blob[off:off + length]
        raw = zlib.decompress(rec[4:])
        rr = _Reader(raw)
        rr.header()
        out[name] = simplify(rr.item())
    return out


def read_rdata(path):
    """.RData (RDX2, gzip): returns {name: object}."""
    with open(path, "rb") as f:
        raw = gzip.decompress(f.read())
    assert raw[:5] == b"RDX2\n", raw[:8]
    r = _Reader(raw[5:])
    r.header()
    return {tag: simplify(val) for tag, val in r.item()}


def read_phylo(path, name=None):
    """The `phylo` tree stored in an .RData file as {edge (n x 2, 1-based), edge_length, Nnode, ntip}: the arguments of
    genealogy.summarize_phylo.  `name`: the R variable (default: the only / first object of class "phylo")."""
    objs = read_rdata(path)
    keys = [name] if name else list(objs)
    for k in keys:
        t = objs[k]
        if isinstance(t, dict) and "edge" in t and "edge.length" in t:
            edge = np.asarray(t["edge"], dtype=np.int64)
            nnode = int(np.ravel(t["Nnode"])[0])
            return {"edge": edge, "edge_length": np.asarray(t["edge.length"], dtype=np.float64), "Nnode": nnode,
                    "ntip": int(len(t["tip.label"])) if "tip.label" in t else int(edge.max()) - nnode}
    raise ValueError(f"no phylo object in {path}")

