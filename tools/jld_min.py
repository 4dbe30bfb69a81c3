"""Minimal reader for the JLD (HDF5 superblock v0) files shipped with the reference.

Test/fixture tooling only.  There is no h5py / libhdf5 in this image, so this walks the
few HDF5 structures the reference's data files actually use:

  * 512-byte JLD user block, then an HDF5 v0 superblock (all file addresses are relative
    to base address 512);
  * version-1 object headers with continuation blocks;
  * root-group links stored either as link messages (type 0x06, `381raypaths.jld`) or in a
    fractal-heap direct block (`381traces.jld`, `model.jld`);
  * contiguous (class 1) and compact (class 0) data layouts;
  * arrays of 8-byte object references pointing at scalar / vector datasets.

It is NOT a general HDF5 reader.
"""
from __future__ import annotations

import re
import struct

import numpy as np

BASE = 512  # JLD user block


class JLDFile:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if not self.buf.startswith(b"Julia data file (HDF5)"):
            raise ValueError(f"{path}: not a JLD file")
        if self.buf[BASE:BASE + 8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError(f"{path}: HDF5 signature not found behind the user block")
        self.links = self._root_links()

    # ---- object headers -------------------------------------------------------------
    def messages(self, addr: int):
        """Return [(type, payload bytes)] of the v1 object header at file address `addr`."""
        b = self.buf
        a = addr + BASE
        if b[a] != 1:
            raise ValueError(f"object header at {addr}: version {b[a]} unsupported")
        nmsg, = struct.unpack_from("<H", b, a + 2)
        hsize, = struct.unpack_from("<I", b, a + 8)
        out = []
        blocks = [(a + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(out) < nmsg:
                t, s, _fl = struct.unpack_from("<HHB", b, p)
                d = b[p + 8:p + 8 + s]
                out.append((t, d))
                if t == 0x10:  # continuation
                    o, ln = struct.unpack_from("<QQ", d)
                    blocks.append((o + BASE, ln))
                p += 8 + s
        return out

    def raw(self, addr: int) -> bytes:
        """Raw bytes of the dataset whose object header is at `addr`."""
        for t, d in self.messages(addr):
            if t == 8:  # data layout, version 3
                cls = d[1]
                if cls == 0:  # compact
                    sz, = struct.unpack_from("<H", d, 2)
                    return d[4:4 + sz]
                if cls == 1:  # contiguous
                    o, ln = struct.unpack_from("<QQ", d, 2)
                    return self.buf[o + BASE:o + BASE + ln]
                raise ValueError(f"layout class {cls} unsupported")
        raise ValueError(f"object at {addr} has no layout message")

    def shape(self, addr: int):
        """HDF5 dims (the reverse of the Julia dims)."""
        for t, d in self.messages(addr):
            if t == 1:  # dataspace v1/v2
                ver, rank, flags = d[0], d[1], d[2]
                off = 8 if ver == 1 else 4
                return tuple(struct.unpack_from("<Q", d, off + 8 * i)[0] for i in range(rank))
        return ()

    # ---- root group ---------------------------------------------------------------------
    def _root_links(self):
        b = self.buf
        # v0 superblock: root symbol-table entry at BASE+56: link-name offset, object header address
        root_addr, = struct.unpack_from("<Q", b, BASE + 56 + 8)
        links = {}
        heap_addr = None
        for t, d in self.messages(root_addr):
            if t == 6:  # link message
                name, addr = self._parse_link(d, 0)[:2]
                links[name] = addr
            elif t == 2:  # link info: fractal heap address of dense storage
                flags = d[1]
                off = 2 + (8 if flags & 1 else 0)
                heap_addr, = struct.unpack_from("<Q", d, off)
        if not links and heap_addr not in (None, 0xFFFFFFFFFFFFFFFF):
            # dense storage: scan every fractal-heap direct block for link records
            for m in re.finditer(b"FHDB", b):
                i = m.start() + 21
                while i < len(b) and b[i] == 1:
                    try:
                        name, addr, i = self._parse_link(b, i)
                    except (UnicodeDecodeError, struct.error):
                        break
                    links[name] = addr
        return links

    @staticmethod
    def _parse_link(b, i):
        flags = b[i + 1]
        j = i + 2
        if flags & 0x08:
            j += 1  # link type
        if flags & 0x04:
            j += 8  # creation order
        if flags & 0x10:
            j += 1  # charset
        nbytes = 1 << (flags & 3)
        ln = int.from_bytes(b[j:j + nbytes], "little")
        j += nbytes
        name = bytes(b[j:j + ln]).decode("ascii")
        addr, = struct.unpack_from("<Q", b, j + ln)
        return name, addr, j + ln + 8

    # ---- typed helpers --------------------------------------------------------------------
    def f64(self, name_or_addr) -> np.ndarray:
        addr = self.links[name_or_addr] if isinstance(name_or_addr, str) else name_or_addr
        a = np.frombuffer(self.raw(addr), "<f8")
        shp = self.shape(addr)
        return a.reshape(shp) if shp and int(np.prod(shp)) == a.size else a

    def refs(self, name_or_addr) -> np.ndarray:
        addr = self.links[name_or_addr] if isinstance(name_or_addr, str) else name_or_addr
        return np.frombuffer(self.raw(addr), "<u8")

    def f64_via_refs(self, name: str) -> np.ndarray:
        """Dataset of object references, each pointing at a scalar f64 dataset (Any[] arrays)."""
        return np.array([struct.unpack("<d", self.raw(int(r))[:8])[0] for r in self.refs(name)])
