"""Readers/writers for the containers exchanged with the reference harness (oracle/ref_harness.c).

TEST INFRASTRUCTURE: imported by tests/ and by tests/golden/make_golden.py only.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

FLAG_TEMPORAL, FLAG_DISTRIBUTION, FLAG_VOLUME = 1, 2, 4


@dataclass
class RefProperty:
    name: str
    flags: int
    dim: tuple
    num_values: int
    perframe: dict = field(default_factory=dict)  # frame -> values (dense np.float32)
    full: np.ndarray | None = None
    full_range: tuple | None = None
    meta: dict = field(default_factory=dict)  # (kind, beg) -> dict(min_value,...)
    aggregate: dict | None = None  # multi-valued temporals: per-frame mean / var / ext


def read_refout(path: str) -> dict:
    """Parse an MDREFOUT file -> {name: RefProperty}."""
    b = open(path, "rb").read()
    assert b[:8] == b"MDREFOUT", b[:8]
    ver, nprops = struct.unpack_from("<II", b, 8)
    off = 16
    props = []
    for _ in range(nprops):
        name = b[off:off + 64].split(b"\0")[0].decode(); off += 64
        flags, = struct.unpack_from("<I", b, off); off += 4
        dim = struct.unpack_from("<4i", b, off); off += 16
        nv, = struct.unpack_from("<Q", b, off); off += 8
        props.append(RefProperty(name, flags, dim, nv))
    while off < len(b):
        kind, pi, beg, end = struct.unpack_from("<IIqq", b, off); off += 24
        mn, mx, r0, r1, r2, r3 = struct.unpack_from("<6f", b, off); off += 24
        storage, = struct.unpack_from("<I", b, off); off += 4
        cnt, = struct.unpack_from("<Q", b, off); off += 8
        p = props[pi]
        if storage == 0:
            vals = np.frombuffer(b, dtype=np.float32, count=cnt, offset=off).copy(); off += 4 * cnt
        else:
            rec = np.frombuffer(b, dtype=np.dtype([("i", "<u4"), ("v", "<f4")]), count=cnt, offset=off); off += 8 * cnt
            vals = np.zeros(p.num_values, dtype=np.float32)
            vals[rec["i"]] = rec["v"]
        meta = dict(min_value=mn, max_value=mx, min_range=(r0, r1), max_range=(r2, r3))
        if kind == 0:
            p.perframe[beg] = vals
        elif kind == 2:   # per-frame aggregates of a multi-valued temporal: mean[F] | var[F] | (min, max)[F]
            na = cnt // 4; p.aggregate = dict(mean=vals[:na], var=vals[na:2 * na], ext=vals[2 * na:].reshape(na, 2))
        else:
            p.full = vals; p.full_range = (beg, end)
        p.meta[(kind, beg)] = meta
    return {p.name: p for p in props}


def read_sysinfo(path: str) -> dict:
    b = open(path, "rb").read()
    assert b[:8] == b"MDSYSINF"
    n, = struct.unpack_from("<Q", b, 8); off = 16
    mass = np.frombuffer(b, np.float32, n, off).copy(); off += 4 * n
    z = np.frombuffer(b, np.uint32, n, off).copy(); off += 4 * n
    names = [b[off + 8 * i: off + 8 * i + 8].split(b"\0")[0].decode() for i in range(n)]; off += 8 * n
    ncomp, = struct.unpack_from("<Q", b, off); off += 8
    comp_off = np.frombuffer(b, np.uint32, ncomp + 1, off).copy(); off += 4 * (ncomp + 1)
    noff, nconn = struct.unpack_from("<QQ", b, off); off += 16
    conn_off = np.frombuffer(b, np.uint32, noff, off).copy(); off += 4 * noff
    conn_idx = np.frombuffer(b, np.int32, nconn, off).copy(); off += 4 * nconn
    cell = struct.unpack_from("<6dI", b, off); off += 56  # x,xy,xz,y,yz,z,flags (+pad)
    x = np.frombuffer(b, np.float32, n, off).copy(); off += 4 * n
    y = np.frombuffer(b, np.float32, n, off).copy(); off += 4 * n
    zc = np.frombuffer(b, np.float32, n, off).copy(); off += 4 * n
    return dict(n=n, mass=mass, z=z, names=names, comp_off=comp_off, conn_off=conn_off, conn_idx=conn_idx,
                cell=cell, x=x, y=y, zc=zc)


def write_raw_traj(path: str, frames_xyz: np.ndarray, cells: np.ndarray, flags: np.ndarray) -> None:
    """frames_xyz: [F,3,N] float32; cells: [F,6] float64 (x,xy,xz,y,yz,z); flags: [F] uint32."""
    F, three, N = frames_xyz.shape
    assert three == 3
    with open(path, "wb") as f:
        f.write(b"MDRAWTRJ"); f.write(struct.pack("<QQ", F, N))
        for i in range(F):
            f.write(np.asarray(cells[i], np.float64).tobytes())
            f.write(struct.pack("<II", int(flags[i]), 0))
            f.write(np.ascontiguousarray(frames_xyz[i], np.float32).tobytes())


def write_gro(path: str, resid, resname, atomname, xyz_A: np.ndarray, box_A) -> None:
    """Minimal .gro writer (nm, 3 decimals). xyz_A: [N,3] in Angstrom."""
    n = len(atomname)
    with open(path, "w") as f:
        f.write("generated\n%d\n" % n)
        for i in range(n):
            f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (resid[i] % 100000, resname[i], atomname[i], (i + 1) % 100000,
                                                          xyz_A[i, 0] * 0.1, xyz_A[i, 1] * 0.1, xyz_A[i, 2] * 0.1))
        if len(box_A) == 3:
            f.write("%10.5f%10.5f%10.5f\n" % tuple(b * 0.1 for b in box_A))
        else:  # gro triclinic order: v1x v2y v3z v1y v1z v2x v2z v3x v3y
            f.write(" ".join("%10.5f" % (b * 0.1) for b in box_A) + "\n")


def read_raw_traj(path: str):
    """-> (frames [F,3,N] float32, cells [F,6] float64, flags [F] uint32)."""
    b = np.memmap(path, dtype=np.uint8, mode="r")
    assert bytes(b[:8]) == b"MDRAWTRJ"
    F, N = struct.unpack("<QQ", bytes(b[8:24]))
    fb = 56 + 12 * N
    frames = np.empty((F, 3, N), np.float32); cells = np.empty((F, 6), np.float64); flags = np.empty(F, np.uint32)
    for i in range(F):
        o = 24 + i * fb
        cells[i] = np.frombuffer(b[o:o + 48].tobytes(), np.float64)
        flags[i] = struct.unpack("<I", b[o + 48:o + 52].tobytes())[0]
        frames[i] = np.frombuffer(b[o + 56:o + fb].tobytes(), np.float32).reshape(3, N)
    return frames, cells, flags


def write_raw_traj(path: str, frames, cells, flags):
    """inverse of read_raw_traj (MDRAWTRJ container read by oracle/ref_harness.c)"""
    frames = np.ascontiguousarray(frames, np.float32); F, _, N = frames.shape
    with open(path, "wb") as f:
        f.write(b"MDRAWTRJ"); f.write(struct.pack("<QQ", F, N))
        for i in range(F):
            f.write(np.asarray(cells[i], np.float64).tobytes()); f.write(struct.pack("<II", int(flags[i]), 0)); f.write(frames[i].tobytes())
