"""ctypes binding of the plain-C oracle (oracle/md_oracle.c). TEST INFRASTRUCTURE ONLY.

Builds oracle/build/liboracle.so on first use (gcc, strict IEEE flags).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

ORTHO, TRICLINIC, PBC_X, PBC_Y, PBC_Z, PBC_ALL = 1, 2, 4, 8, 16, 28


class UnitCell(C.Structure):
    _fields_ = [("x", C.c_double), ("xy", C.c_double), ("xz", C.c_double), ("y", C.c_double), ("yz", C.c_double),
                ("z", C.c_double), ("flags", C.c_uint32)]

    @staticmethod
    def ortho(x, y, z, flags=ORTHO | PBC_ALL):
        return UnitCell(float(x), 0.0, 0.0, float(y), 0.0, float(z), flags)

    @staticmethod
    def from_params(x, xy, xz, y, yz, z, flags):
        return UnitCell(float(x), float(xy), float(xz), float(y), float(yz), float(z), int(flags))


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "build", "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("md_oracle.c", "md_oracle.h")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
            build()
        _lib = C.CDLL(path)
        _lib.mdo_rdf_frame.restype = C.c_uint64
        _lib.mdo_sdf_frame.restype = C.c_uint64
        _lib.mdo_count_pairs.restype = C.c_uint64
        _lib.mdo_distance.restype = C.c_float
        _lib.mdo_angle.restype = C.c_float
        _lib.mdo_dihedral.restype = C.c_float
        for n in ("mdo_distance_pos", "mdo_angle_pos", "mdo_dihedral_pos", "mdo_min_distance"): getattr(_lib, n).restype = C.c_float
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, np.int32)


def rdf_frame(x, y, z, ref_idx, trg_idx, cell: UnitCell, min_cutoff, max_cutoff, ref_pos=None, excl_off=None, excl_idx=None):
    x, y, z = _f32(x), _f32(y), _f32(z)
    ref_idx, trg_idx, ref_pos = _i32(ref_idx), _i32(trg_idx), _f32(ref_pos)
    n_ref = len(ref_pos) // 3 if ref_pos is not None and ref_pos.ndim == 1 else (len(ref_pos) if ref_pos is not None else len(ref_idx))
    eo = None if excl_off is None else np.ascontiguousarray(excl_off, np.uint32)
    ei = _i32(excl_idx)
    bins = np.zeros(1024, np.float32); w = np.zeros(1024, np.float32)
    total = lib().mdo_rdf_frame(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float),
                                _p(ref_idx, C.c_int32), _p(ref_pos, C.c_float), C.c_size_t(n_ref),
                                _p(trg_idx, C.c_int32), C.c_size_t(len(trg_idx)), C.byref(cell),
                                C.c_float(min_cutoff), C.c_float(max_cutoff), _p(eo, C.c_uint32), _p(ei, C.c_int32),
                                _p(bins, C.c_float), _p(w, C.c_float))
    return bins, w, int(total)


def rdf_frame_args(x, y, z, mass, ref, trg, cell, min_cutoff, max_cutoff):
    """rdf of one frame with either argument an atom index array or a LIST of index arrays (an array of selections: one centre of mass per selection,
    extract_com; as reference also the exclusion masks, tested against the target's index — its ORDINAL when the targets are centres of mass,
    rdf_cb_excl_mask md_script_functions.inl:5252). The target points are handed to the oracle as coordinate arrays of their own."""
    x, y, z = _f32(x), _f32(y), _f32(z)
    eo = ei = None
    if isinstance(ref, list): rp, eo, ei = group_com(x, y, z, mass, ref)
    else: r = _i32(ref); rp = np.ascontiguousarray(np.stack([x[r], y[r], z[r]], axis=1), np.float32)
    if isinstance(trg, list):
        tp, _, _ = group_com(x, y, z, mass, trg)
        tx, ty, tz = (np.ascontiguousarray(tp[:, k]) for k in range(3)); ti = np.arange(len(trg), dtype=np.int32)
    else: tx, ty, tz, ti = x, y, z, _i32(trg)
    return rdf_frame(tx, ty, tz, None, ti, cell, min_cutoff, max_cutoff, ref_pos=rp, excl_off=eo, excl_idx=ei)


def group_com(x, y, z, mass, groups):
    """groups: list of int32 index arrays -> (AoS positions [n,3], offsets, concatenated indices)"""
    x, y, z, mass = _f32(x), _f32(y), _f32(z), _f32(mass)
    off = np.zeros(len(groups) + 1, np.uint32); off[1:] = np.cumsum([len(g) for g in groups])
    idx = np.ascontiguousarray(np.concatenate(groups), np.int32)
    out = np.zeros((len(groups), 3), np.float32)
    lib().mdo_group_com(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(mass, C.c_float), _p(idx, C.c_int32), _p(off, C.c_uint32),
                        C.c_size_t(len(groups)), _p(out, C.c_float))
    return out, off, idx


def sdf_frame(x, y, z, init_xyz, mass, struct_idx, trg_idx, conn_off, conn_idx, cell: UnitCell, cutoff, vol=None, want_matrices=False):
    x, y, z = _f32(x), _f32(y), _f32(z)
    ix, iy, iz = (_f32(a) for a in init_xyz)
    mass = _f32(mass)
    struct_idx = _i32(struct_idx); n_struct, ssize = struct_idx.shape
    trg_idx = _i32(trg_idx)
    conn_off = np.ascontiguousarray(conn_off, np.uint32); conn_idx = _i32(conn_idx)
    if vol is None:
        vol = np.zeros(128 ** 3, np.float32)
    mats = np.zeros((n_struct, 16), np.float32) if want_matrices else None
    n = lib().mdo_sdf_frame(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(ix, C.c_float), _p(iy, C.c_float), _p(iz, C.c_float),
                            _p(mass, C.c_float), _p(struct_idx, C.c_int32), C.c_size_t(n_struct), C.c_size_t(ssize),
                            _p(trg_idx, C.c_int32), C.c_size_t(len(trg_idx)), _p(conn_off, C.c_uint32), _p(conn_idx, C.c_int32),
                            C.c_size_t(len(conn_off)), C.byref(cell), C.c_float(cutoff), _p(vol, C.c_float), _p(mats, C.c_float))
    return (vol, int(n), mats) if want_matrices else (vol, int(n))


def density_frame(x, y, z, mass, idx, init_cell: UnitCell, axis):
    x, y, z, mass, idx = _f32(x), _f32(y), _f32(z), _f32(mass), _i32(idx)
    bins = np.zeros(1024, np.float32); w = np.zeros(1024, np.float32)
    lib().mdo_density_frame(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(mass, C.c_float), _p(idx, C.c_int32),
                            C.c_size_t(len(idx)), C.byref(init_cell), C.c_int(axis), _p(bins, C.c_float), _p(w, C.c_float))
    return bins, w


def distance(x, y, z, a, b, cell):
    x, y, z = _f32(x), _f32(y), _f32(z)
    return float(lib().mdo_distance(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), C.c_int32(a), C.c_int32(b), C.byref(cell)))


def angle(x, y, z, a, b, c):
    x, y, z = _f32(x), _f32(y), _f32(z)
    return float(lib().mdo_angle(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), C.c_int32(a), C.c_int32(b), C.c_int32(c)))


def dihedral(x, y, z, a, b, c, d, cell):
    x, y, z = _f32(x), _f32(y), _f32(z)
    return float(lib().mdo_dihedral(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), C.c_int32(a), C.c_int32(b), C.c_int32(c), C.c_int32(d), C.byref(cell)))


def count_pairs(x, y, z, ref_idx, trg_idx, cell, cell_ext, cutoff):
    x, y, z = _f32(x), _f32(y), _f32(z); ref_idx, trg_idx = _i32(ref_idx), _i32(trg_idx)
    return int(lib().mdo_count_pairs(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(ref_idx, C.c_int32), C.c_size_t(len(ref_idx)),
                                     _p(trg_idx, C.c_int32), C.c_size_t(len(trg_idx)), C.byref(cell), C.c_double(cell_ext), C.c_double(cutoff)))


def arg_position(x, y, z, mass, arg, cell):
    """argument of distance/angle/dihedral: int -> the atom's position, index array -> centre of mass, list of index arrays (an ARRAY of
    selections) -> centre of the selections' centres (coordinate_extract_com)"""
    x, y, z, mass = _f32(x), _f32(y), _f32(z), _f32(mass)
    if isinstance(arg, list) and len(arg) > 1:
        sels = [_i32(g) for g in arg]; off = np.zeros(len(sels) + 1, np.uint32); off[1:] = np.cumsum([len(g) for g in sels])
        idx = np.ascontiguousarray(np.concatenate(sels), np.int32); out = np.zeros(3, np.float32)
        lib().mdo_arg_position_parts(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(mass, C.c_float), _p(idx, C.c_int32), _p(off, C.c_uint32),
                                     C.c_size_t(len(sels)), C.byref(cell), _p(out, C.c_float))
        return out
    if isinstance(arg, list): arg = arg[0]
    direct = np.ndim(arg) == 0
    idx = np.ascontiguousarray([int(arg)] if direct else arg, np.int32); out = np.zeros(3, np.float32)
    lib().mdo_arg_position(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(mass, C.c_float), _p(idx, C.c_int32), C.c_size_t(len(idx)),
                           C.c_int(1 if direct else 0), C.byref(cell), _p(out, C.c_float))
    return out


def distance_args(x, y, z, mass, a, b, cell):
    pa, pb = arg_position(x, y, z, mass, a, cell), arg_position(x, y, z, mass, b, cell)
    return np.float32(lib().mdo_distance_pos(_p(pa, C.c_float), _p(pb, C.c_float), C.byref(cell)))


def angle_args(x, y, z, mass, a, b, c, cell):
    p = [arg_position(x, y, z, mass, k, cell) for k in (a, b, c)]
    return np.float32(lib().mdo_angle_pos(_p(p[0], C.c_float), _p(p[1], C.c_float), _p(p[2], C.c_float)))


def dihedral_args(x, y, z, mass, a, b, c, d, cell):
    p = np.ascontiguousarray(np.stack([arg_position(x, y, z, mass, k, cell) for k in (a, b, c, d)]), np.float32)
    return np.float32(lib().mdo_dihedral_pos(_p(p, C.c_float), C.byref(cell)))


def xtc_frame_offsets(blob):
    blob = np.ascontiguousarray(blob, np.uint8); offs = np.zeros(4096, np.int64)
    lib().mdo_xtc_frame_offsets.restype = C.c_size_t
    n = lib().mdo_xtc_frame_offsets(_p(blob, C.c_uint8), C.c_size_t(blob.size), _p(offs, C.c_int64), C.c_size_t(offs.size))
    return offs[:n + 1].copy()


def xtc_decode_frame(blob, beg, end, num_atoms):
    """-> (ok, xyz [3, n], UnitCell, step, time)"""
    blob = np.ascontiguousarray(blob, np.uint8); xyz = np.zeros((3, num_atoms), np.float32); cell = UnitCell(); st = C.c_int32(); tm = C.c_float()
    ok = lib().mdo_xtc_decode_frame(C.c_void_p(blob.ctypes.data + int(beg)), C.c_size_t(int(end - beg)), C.c_size_t(num_atoms),
                                    C.c_void_p(xyz[0].ctypes.data), C.c_void_p(xyz[1].ctypes.data), C.c_void_p(xyz[2].ctypes.data), C.byref(cell), C.byref(st), C.byref(tm))
    return bool(ok), xyz, cell, int(st.value), float(tm.value)


def min_distance(x, y, z, a, b, cell):
    x, y, z = _f32(x), _f32(y), _f32(z); a, b = _i32(a), _i32(b)
    return np.float32(lib().mdo_min_distance(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(a, C.c_int32), C.c_size_t(len(a)), _p(b, C.c_int32), C.c_size_t(len(b)), C.byref(cell)))


def shape_weights(x, y, z, mass, idx, cell):
    x, y, z = _f32(x), _f32(y), _f32(z); idx = _i32(idx); out = np.zeros(3, np.float32)
    m = None if mass is None else _f32(mass)
    lib().mdo_shape_weights(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), None if m is None else _p(m, C.c_float), _p(idx, C.c_int32), C.c_size_t(len(idx)), C.byref(cell), _p(out, C.c_float))
    return out


def plane_frame(x, y, z, idx, conn_off, conn_idx, cell):
    x, y, z = _f32(x), _f32(y), _f32(z); idx = _i32(idx); co = np.ascontiguousarray(conn_off, np.uint32); ci = _i32(conn_idx); out = np.zeros(4, np.float32)
    lib().mdo_plane_frame(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(idx, C.c_int32), C.c_size_t(len(idx)), _p(co, C.c_uint32), _p(ci, C.c_int32),
                          C.c_size_t(len(co)), C.byref(cell), _p(out, C.c_float))
    return out


def distance_pair(x, y, z, a, b, cell):
    x, y, z = _f32(x), _f32(y), _f32(z); a, b = _i32(a), _i32(b); out = np.zeros(len(a) * len(b), np.float32)
    lib().mdo_distance_pair(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(a, C.c_int32), C.c_size_t(len(a)), _p(b, C.c_int32), C.c_size_t(len(b)), C.byref(cell), _p(out, C.c_float))
    return out


def distance_pair_args(x, y, z, mass, a, b, cell):
    """distance_pair(a, b) where a / b is an index array (atoms) or a list of index arrays (array of selections -> one extract_com centre each)"""
    def positions(arg):
        if isinstance(arg, (list, tuple)): return np.asarray(group_com(x, y, z, mass, arg)[0], np.float32).reshape(-1, 3)   # extract_com :857
        arg = np.asarray(arg, np.int32); return np.stack([_f32(x)[arg], _f32(y)[arg], _f32(z)[arg]], axis=1)
    pa, pb = np.ascontiguousarray(positions(a), np.float32), np.ascontiguousarray(positions(b), np.float32); out = np.zeros(len(pa) * len(pb), np.float32)
    lib().mdo_distance_pair_pos(_p(pa, C.c_float), C.c_size_t(len(pa)), _p(pb, C.c_float), C.c_size_t(len(pb)), C.byref(cell), _p(out, C.c_float))
    return out


def aggregate(values):
    """-> (min, max, mean, var) as the reference folds one frame of a multi-valued temporal"""
    v = _f32(values); out = np.zeros(4, np.float32)
    lib().mdo_aggregate(_p(v, C.c_float), C.c_size_t(len(v)), _p(out, C.c_float))
    return out


def rmsd_frame(x, y, z, init_xyz, mass, idx, conn_off, conn_idx, cell):
    x, y, z, mass = _f32(x), _f32(y), _f32(z), _f32(mass); ix, iy, iz = _f32(init_xyz[0]), _f32(init_xyz[1]), _f32(init_xyz[2])
    idx = _i32(idx); co = np.ascontiguousarray(conn_off, np.uint32); ci = _i32(conn_idx)
    lib().mdo_rmsd_frame.restype = C.c_double
    return np.float32(lib().mdo_rmsd_frame(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), _p(ix, C.c_float), _p(iy, C.c_float), _p(iz, C.c_float),
                                           _p(mass, C.c_float), _p(idx, C.c_int32), C.c_size_t(len(idx)), _p(co, C.c_uint32), _p(ci, C.c_int32), C.c_size_t(len(co)), C.byref(cell)))


def within(x, y, z, sel, radius, cell, rmin=0.0):
    """-> ascending atom indices within `radius` (and at least `rmin`) of any atom of `sel` (sel itself removed)"""
    x, y, z = _f32(x), _f32(y), _f32(z); sel = _i32(sel); mask = np.zeros(len(x), np.uint8)
    lib().mdo_within_range.restype = C.c_size_t
    n = lib().mdo_within_range(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), C.c_size_t(len(x)), _p(sel, C.c_int32), C.c_size_t(len(sel)), C.c_float(rmin), C.c_float(radius), C.byref(cell), _p(mask, C.c_uint8))
    idx = np.nonzero(mask)[0].astype(np.int32); assert len(idx) == n
    return idx
