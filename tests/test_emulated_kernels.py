"""Kernels whose threads do not cooperate, EXECUTED on the CPU from the product's own .cu source (tests/emul: g++ + a small CUDA shim) and
compared with the reference's golden values. This is how device code written without GPU time left is checked before its first launch:
it proves the source's logic and operation order (every operation on these paths is IEEE add/mul/div/sqrt, identical on host and device);
it cannot prove launch configuration or memory behaviour, which is what tests/test_zz_gpu_new_ops.py is for."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import oracle_lib as O  # noqa: E402
from helpers import load_golden, golden_system, sel_element, cell_from_row, dense_from_sparse  # noqa: E402


class _Cell(C.Structure):   # mdgpu_unitcell_t
    _fields_ = [("x", C.c_double), ("xy", C.c_double), ("xz", C.c_double), ("y", C.c_double), ("yz", C.c_double), ("z", C.c_double), ("flags", C.c_uint32)]


@pytest.fixture(scope="module")
def emul():
    import build_emul
    return C.CDLL(build_emul.build("sdf"))


@pytest.fixture(scope="module")
def emul_props():
    import build_emul
    return C.CDLL(build_emul.build("props"))


def _cells(g):
    F = g["frames"].shape[0]; cells = (_Cell * F)()
    for f in range(F):
        x, xy, xz, y, yz, z = (float(v) for v in g["cells"][f]); cells[f] = _Cell(x, xy, xz, y, yz, z, int(g["cell_flags"][f]))
    return cells


def unwrap_pairs(count, conn_off, conn_idx):
    """(child, parent) pairs in the order md_util_unwrap_vec4(xyzw, NULL, count, bond, cell) visits them (md_util.c:8738-8819): breadth
    first over the bonds of GLOBAL atoms 0..count-1 (the local index is used as the atom index there), neighbours >= count skipped."""
    na = len(conn_off) - 1; visited = np.zeros(na + 1, bool); out = []
    for seed in range(count):
        if seed >= na or visited[seed]: continue
        visited[seed] = True; queue = [seed]; qh = 0
        while qh < len(queue):
            cur = queue[qh]; qh += 1
            for k in range(conn_off[cur], conn_off[cur + 1]):
                nx = int(conn_idx[k])
                if nx < 0 or nx >= count or visited[nx]: continue
                out.append((nx, cur)); visited[nx] = True; queue.append(nx)
    return np.asarray(out, np.int32).reshape(-1, 2)


def run_rmsd(lib, g, s, idx):
    frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape
    cells = (_Cell * F)()
    for f in range(F):
        x, xy, xz, y, yz, z = (float(v) for v in g["cells"][f]); cells[f] = _Cell(x, xy, xz, y, yz, z, int(g["cell_flags"][f]))
    idx = np.ascontiguousarray(idx, np.int32); mass = np.ascontiguousarray(s["mass"], np.float32)
    pairs = np.ascontiguousarray(unwrap_pairs(len(idx), s["conn_off"], s["conn_idx"]))
    out = np.zeros(F, np.float32); fp = C.POINTER(C.c_float); ip = C.POINTER(C.c_int32)
    lib.emul_rmsd.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), fp, C.c_size_t, fp, ip, C.c_uint32, ip, C.c_uint32, fp]
    rc = lib.emul_rmsd(frames.ctypes.data_as(fp), 3 * na, na, F, cells, frames[0].ctypes.data_as(fp), na, mass.ctypes.data_as(fp),
                       idx.ctypes.data_as(ip), len(idx), pairs.ctypes.data_as(ip), len(pairs), out.ctypes.data_as(fp))
    assert rc == 0
    return out


def test_k_rmsd_source_matches_the_reference(emul):
    """k_rmsd (viamd_b200/csrc/sdf.cu) run on the CPU: bit-equal to the reference's rmsd() on the orthorhombic water box, the 1ALA
    trajectory (50 frames, 153 atoms) and the triclinic cell that changes every frame (contiguous and scattered selections)."""
    g = load_golden("water6.npz"); s = golden_system(g)
    assert np.array_equal(run_rmsd(emul, g, s, np.arange(30)), g["rm__full"])
    g = load_golden("ala50.npz"); s = golden_system(g)
    assert np.array_equal(run_rmsd(emul, g, s, np.arange(153)), g["rma__full"])
    g = load_golden("tric6.npz"); s = golden_system(g); r = load_golden("tric6_rmsd.npz")
    for key, idx in (("rmt", np.arange(30)), ("rma", np.arange(99, 160)), ("rmo", sel_element(s, 8))):
        assert np.array_equal(run_rmsd(emul, g, s, idx), r[f"{key}__full"]), key
    assert np.array_equal(run_rmsd(emul, g, s, np.zeros(0, np.int32)), np.zeros(g["frames"].shape[0], np.float32))


FP, IP = C.POINTER(C.c_float), C.POINTER(C.c_int32)


def test_emulation_reproduces_a_gpu_validated_kernel(emul_props):
    """Check of the emulation itself: k_temporal has passed on the B200 against these goldens; run on the CPU from the same source it must
    reproduce them too (distance bit-equal; angle / dihedral bit-equal as well here, because host and reference share glibc's acosf / atan2f)."""
    g = load_golden("water6.npz"); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
    emul_props.emul_temporal.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), C.c_int, IP, FP]
    for key, op, atoms in (("d", 6, (0, 9, 0, 0)), ("a", 7, (0, 1, 2, 0)), ("t", 8, (0, 3, 6, 9))):
        out = np.zeros(F, np.float32); at = np.asarray(atoms, np.int32)
        assert emul_props.emul_temporal(frames.ctypes.data_as(FP), 3 * na, na, F, cells, op, at.ctypes.data_as(IP), out.ctypes.data_as(FP)) == 0
        assert np.array_equal(out, g[f"{key}__full"]), key


def test_k_distance_pair_source_matches_the_reference(emul_props):
    """k_distance_pair (viamd_b200/csrc/props.cu) run on the CPU thread by thread: bit-equal to the reference's distance_pair() matrices in the
    orthorhombic box and in the triclinic cell that changes every frame (5 x 11 and 3 x 216 pairs per frame)."""
    p = load_golden("pairs6.npz")
    emul_props.emul_distance_pair.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), IP, C.c_uint32, IP, C.c_uint32, FP]
    for tag, name in (("w", "water6.npz"), ("t", "tric6.npz")):
        g = load_golden(name); s = golden_system(g); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
        for key, a, b in (("dp", np.arange(0, 5, dtype=np.int32), np.arange(19, 30, dtype=np.int32)), ("dpo", np.arange(0, 3, dtype=np.int32), sel_element(s, 8))):
            a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32); out = np.zeros(F * len(a) * len(b), np.float32)
            assert emul_props.emul_distance_pair(frames.ctypes.data_as(FP), 3 * na, na, F, cells, a.ctypes.data_as(IP), len(a), b.ctypes.data_as(IP), len(b), out.ctypes.data_as(FP)) == 0
            assert np.array_equal(out, p[f"{tag}_{key}__full"]), (tag, key)


def test_k_plane_and_k_com_rows_sources_match_the_reference(emul, emul_props):
    """k_plane (sdf.cu) and k_com_rows (props.cu) run on the CPU: plane() of 30 bonded atoms and of all oxygens, com() of one atom, ortho and
    triclinic — bit-equal to the reference. (com() of a selection takes its position from k_arg_com, which is GPU-validated through distance();
    here that slot is fed with the reference's own values to check the row store.)"""
    p = load_golden("pairs6.npz")
    emul.emul_plane.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), IP, C.c_uint32, IP, C.c_uint32, FP]
    emul_props.emul_com_rows.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.c_int, FP, FP]
    for tag, name in (("w", "water6.npz"), ("t", "tric6.npz")):
        g = load_golden(name); s = golden_system(g); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
        for key, idx in (("pl", np.arange(30, dtype=np.int32)), ("plo", sel_element(s, 8))):
            idx = np.ascontiguousarray(idx, np.int32); pairs = np.ascontiguousarray(unwrap_pairs(len(idx), s["conn_off"], s["conn_idx"])); out = np.zeros(4 * F, np.float32)
            assert emul.emul_plane(frames.ctypes.data_as(FP), 3 * na, na, F, cells, idx.ctypes.data_as(IP), len(idx), pairs.ctypes.data_as(IP), len(pairs), out.ctypes.data_as(FP)) == 0
            assert np.array_equal(out, p[f"{tag}_{key}__full"]), (tag, key)
        out = np.zeros(3 * F, np.float32)
        assert emul_props.emul_com_rows(frames.ctypes.data_as(FP), 3 * na, na, F, 4, None, out.ctypes.data_as(FP)) == 0
        assert np.array_equal(out, p[f"{tag}_ci__full"])
        pos = np.zeros((F, 4, 3), np.float32); pos[:, 0, :] = p[f"{tag}_ca__full"].reshape(F, 3); out = np.zeros(3 * F, np.float32)
        assert emul_props.emul_com_rows(frames.ctypes.data_as(FP), 3 * na, na, F, 0, pos.ctypes.data_as(FP), out.ctypes.data_as(FP)) == 0
        assert np.array_equal(out, p[f"{tag}_ca__full"])


def test_cooperative_emulation_reproduces_a_gpu_validated_kernel(emul_props):
    """Check of emul_launch (threads of a block as host threads, barriers, warp shuffles): k_min_distance has passed on the B200 against
    these goldens and must reproduce them when run this way (256 threads per block, shuffle reduction, shared memory, __syncthreads)."""
    emul_props.emul_min_distance.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), IP, C.c_uint32, IP, C.c_uint32, FP]
    g = load_golden("water6.npz"); s = golden_system(g); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
    h = np.nonzero(np.asarray(s["z"]) == 1)[0].astype(np.int32)
    for key, a, b in (("dmn", np.arange(0, 3), np.arange(99, 648)), ("dmx", np.arange(0, 30), np.arange(99, 151)), ("dmh", h, np.arange(299, 400))):
        a = np.ascontiguousarray(a, np.int32); b = np.ascontiguousarray(b, np.int32); out = np.zeros(F, np.float32)
        assert emul_props.emul_min_distance(frames.ctypes.data_as(FP), 3 * na, na, F, cells, a.ctypes.data_as(IP), len(a), b.ctypes.data_as(IP), len(b), out.ctypes.data_as(FP)) == 0
        assert np.array_equal(out, g[f"{key}__full"]), key
    g = load_golden("tric6.npz"); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
    a = np.arange(0, 30, dtype=np.int32); b = np.arange(99, 151, dtype=np.int32); out = np.zeros(F, np.float32)
    assert emul_props.emul_min_distance(frames.ctypes.data_as(FP), 3 * na, na, F, cells, a.ctypes.data_as(IP), len(a), b.ctypes.data_as(IP), len(b), out.ctypes.data_as(FP)) == 0
    assert np.array_equal(out, g["dmt__full"])


@pytest.fixture(scope="module")
def emul_within():
    import build_emul
    lib = C.CDLL(build_emul.build("within", ["cells", "within"]))
    lib.emul_within_count.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), C.c_uint32, IP, C.c_uint32, C.c_float, C.c_uint32, FP]
    return lib


def _within(lib, g, sel, radius, flags=None):
    frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
    if flags is not None:
        for f in range(F): cells[f].flags = flags
    sel = np.ascontiguousarray(sel, np.int32); out = np.zeros(F, np.float32)
    rc = lib.emul_within_count(frames.ctypes.data_as(FP), 3 * na, na, F, cells, na, sel.ctypes.data_as(IP), len(sel), radius, 1 << 16, out.ctypes.data_as(FP))
    assert rc == 0
    return out


def test_count_within_pipeline_matches_the_reference(emul_within):
    """count(within(radius, selection)) end to end on the CPU: the cell-list kernels of cells.cu (k_frame_geom, k_bin_points, k_scan_cells,
    k_scatter_points, k_aabb — GPU-validated) followed by the new k_within_mark / k_within_count, with the launch sequences of the product.
    Equal to the reference's counts on the water goldens; equal to the (reference-pinned) oracle in the triclinic cell that changes every
    frame, with one axis non-periodic (grid fitted to the data), and for an empty selection."""
    g = load_golden("water6.npz")
    assert np.array_equal(_within(emul_within, g, np.arange(0, 3), 4.0), g["cw__full"])
    assert np.array_equal(_within(emul_within, g, np.arange(9, 12), 7.5), g["cw2__full"])
    assert np.array_equal(_within(emul_within, g, np.zeros(0, np.int32), 4.0), np.zeros(g["frames"].shape[0], np.float32))
    for name, flags, sel, radius in (("tric6.npz", None, np.arange(0, 30), 5.0), ("tric6.npz", None, np.arange(100, 103), 8.5),
                                     ("water6.npz", 1 | 4 | 8, np.arange(0, 12), 6.5), ("water6.npz", 1 | 8 | 16, np.arange(300, 340), 3.0)):
        g = load_golden(name); got = _within(emul_within, g, sel, radius, flags)
        for f in range(g["frames"].shape[0]):
            x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f] if flags is None else flags)
            assert got[f] == len(O.within(x, y, z, np.asarray(sel, np.int32), radius, cell)), (name, flags, radius, f)


def test_sdf_pipeline_emulated_matches_the_reference_voxels():
    """sdf() end to end on the CPU from the product's sources: target cell list (cells.cu), reference-frame fit per structure (k_sdf_ref0,
    k_sdf_fit: unwrap, covariance, svd3, Kabsch) and the AABB gather + voxel scatter (k_sdf_scatter: ballots, shuffles, REDUX.OR, ring
    compaction) — all GPU-validated kernels, here as a CPU regression net. Per-frame voxels bit-equal to the reference, orthorhombic and
    triclinic; a batch of all frames gives their sum."""
    import build_emul
    lib = C.CDLL(build_emul.build("sdfpipe", ["cells", "sdf"]))
    UP = C.POINTER(C.c_uint32)
    lib.emul_sdf.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), FP, C.c_size_t, FP, IP, C.c_uint32, C.c_uint32, IP, C.c_uint32, IP, C.c_uint32,
                             C.c_float, C.c_uint32, UP, C.POINTER(C.c_ulonglong)]
    for name, key in (("water6.npz", "v"), ("tric6.npz", "vt")):
        g = load_golden(name); s = golden_system(g); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g)
        structs = np.ascontiguousarray(np.arange(60, dtype=np.int32)); trg = np.ascontiguousarray(sel_element(s, 8), np.int32); mass = np.ascontiguousarray(s["mass"], np.float32)
        pairs = np.ascontiguousarray(unwrap_pairs(3, s["conn_off"], s["conn_idx"]))
        init = np.ascontiguousarray(frames[0])

        def run(fr, cl, nf):
            vol = np.zeros(128 ** 3, np.uint32); tot = (C.c_ulonglong * nf)()
            rc = lib.emul_sdf(fr.ctypes.data_as(FP), 3 * na, na, nf, cl, init.ctypes.data_as(FP), na, mass.ctypes.data_as(FP), structs.ctypes.data_as(IP), 20, 3,
                              trg.ctypes.data_as(IP), len(trg), pairs.ctypes.data_as(IP), len(pairs), 5.0, 1 << 16, vol.ctypes.data_as(UP), tot)
            assert rc == 0
            return vol, np.array(list(tot), np.uint64)

        total = np.zeros(128 ** 3, np.uint64)
        for f in range(F):
            one = (_Cell * 1)(cells[f])
            vol, tot = run(np.ascontiguousarray(frames[f:f + 1]), one, 1)
            ref = dense_from_sparse(g[f"{key}__pf{f}_idx"], g[f"{key}__pf{f}_val"])
            assert np.array_equal(vol.astype(np.float32), ref), (name, f)
            assert tot[0] == int(ref.sum()); total += vol
        vol, tot = run(frames, cells, F)
        assert np.array_equal(vol.astype(np.uint64), total) and tot.sum() == total.sum()


def test_selection_arguments_and_group_centres_emulated(emul_props):
    """k_arg_com (the 8-lane replay of the AVX2 reference's periodic centre of mass, warp shuffles, Cephes sincos) + k_temporal, and
    k_group_com — GPU-validated kernels as a CPU regression net: distances between centres of mass bit-equal to the reference (water and
    1ALA goldens), angles / dihedrals too (host libm = the reference's), group centres equal to the oracle."""
    PP = C.POINTER(C.c_int32) * 4
    emul_props.emul_temporal_args.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), FP, C.c_int, PP, C.c_uint32 * 4, C.c_int * 4, FP]
    emul_props.emul_group_com.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, IP, C.POINTER(C.c_uint32), C.c_uint32, FP, FP]

    def run(g, s, op, args):
        frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g); mass = np.ascontiguousarray(s["mass"], np.float32)
        arrs = [np.ascontiguousarray([a] if np.ndim(a) == 0 else a, np.int32) for a in args] + [np.zeros(1, np.int32)] * (4 - len(args))
        ptrs = PP(*[a.ctypes.data_as(IP) for a in arrs]); cnt = (C.c_uint32 * 4)(*[len(a) for a in arrs]); direct = (C.c_int * 4)(*([int(np.ndim(a) == 0) for a in args] + [1] * (4 - len(args))))
        out = np.zeros(F, np.float32)
        assert emul_props.emul_temporal_args(frames.ctypes.data_as(FP), 3 * na, na, F, cells, mass.ctypes.data_as(FP), op, ptrs, cnt, direct, out.ctypes.data_as(FP)) == 0
        return out

    g = load_golden("water6.npz"); s = golden_system(g); co = s["comp_off"]; res = lambda r: np.arange(co[r - 1], co[r], dtype=np.int32)
    assert np.array_equal(run(g, s, 6, (res(1), res(5))), g["dc__full"])
    assert np.array_equal(run(g, s, 6, (np.arange(0, 30), np.arange(99, 151))), g["dg__full"])
    assert np.array_equal(run(g, s, 6, (np.arange(0, 30), 199)), g["dm__full"])
    assert np.array_equal(run(g, s, 7, (res(1), res(2), res(3))), g["ac__full"])
    assert np.array_equal(run(g, s, 8, (res(1), res(2), res(3), res(4))), g["tc__full"])
    g = load_golden("ala50.npz"); s = golden_system(g); co = s["comp_off"]
    assert np.array_equal(run(g, s, 6, (res(1), res(15))), g["dr__full"])
    assert np.array_equal(run(g, s, 8, (res(1), res(5), 99, res(15))), g["tr__full"])

    g = load_golden("water6.npz"); s = golden_system(g); frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape
    groups = [np.arange(co0, co1, dtype=np.int32) for co0, co1 in zip(s["comp_off"][:20], s["comp_off"][1:21])]
    idx = np.ascontiguousarray(np.concatenate(groups)); off = np.zeros(21, np.uint32); off[1:] = np.cumsum([len(x) for x in groups])
    mass = np.ascontiguousarray(s["mass"], np.float32); out = np.zeros((F, 20, 3), np.float32)
    assert emul_props.emul_group_com(frames.ctypes.data_as(FP), 3 * na, na, F, idx.ctypes.data_as(IP), off.ctypes.data_as(C.POINTER(C.c_uint32)), 20, mass.ctypes.data_as(FP), out.ctypes.data_as(FP)) == 0
    for f in range(F):
        pos, _, _ = O.group_com(*g["frames"][f], s["mass"], groups)
        assert np.array_equal(out[f], np.asarray(pos, np.float32).reshape(20, 3)), f


@pytest.mark.parametrize("case", ("water6", "lowprec", "tric6", "small5", "wide12", "huge12"))
def test_xtc_decode_emulated(case):
    """The device XTC decoder (k_xtc_scan: warp per frame, shared-memory staged stream, speculative 32-group rounds with a serial fallback;
    k_xtc_decode: thread per group, 64/128-bit unpacking) run on the CPU from xtc.cu: equal to the reference reader's decode on the five
    stream classes it decodes correctly, to the written data on the two it does not — GPU-validated kernels as a CPU regression net."""
    import build_emul
    lib = C.CDLL(build_emul.build("xtc"))
    lib.emul_xtc_decode.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, FP]
    g = load_golden("xtc_cases.npz"); blob = np.ascontiguousarray(g[case + "__xtc"]); na = int(g[case + "__na"]); F = len(g[case + "__cells"])
    offs = O.xtc_frame_offsets(blob); assert len(offs) == F + 1
    offs = np.ascontiguousarray(offs, np.uint64); out = np.zeros((F, 3, na), np.float32)
    assert lib.emul_xtc_decode(blob.ctypes.data, offs.ctypes.data_as(C.POINTER(C.c_uint64)), F, na, out.ctypes.data_as(FP)) == 0
    for f in range(F):
        if case in ("wide12", "huge12"): assert np.abs(out[f] - g[case + "__orig"][f]).max() <= 0.02
        else: assert np.array_equal(out[f], g[case + "__frames"][f]), (case, f)


def test_rdf_hot_path_emulated_matches_the_reference_bins():
    """rdf() — the hot path — end to end on the CPU from the product's sources: cell lists, k_rdf_cull (exact target cull into per-home-cell
    candidate lists), k_rdf_pairs_v2 (packed pair loop, per-lane hit queue, branch-free drain, symmetric pairs counted twice, dynamic
    home-cell scheduling), the scalar k_rdf_pairs with exclusion masks for centre-of-mass references, k_rdf_finalize. Per-frame integer
    bins and pair totals bit-equal to the reference: same selection on both sides (symmetric mode), different selections with a min:max
    cutoff, centre-of-mass references, orthorhombic and the triclinic cell that changes every frame."""
    import build_emul
    lib = C.CDLL(build_emul.build("rdfpipe", ["cells", "props", "rdf"]))
    UP = C.POINTER(C.c_uint32)
    lib.emul_rdf.argtypes = [FP, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(_Cell), FP, IP, C.c_uint32, UP, C.c_uint32, IP, C.c_uint32,
                             C.c_float, C.c_float, C.c_int, C.c_uint32, UP, C.POINTER(C.c_ulonglong)]

    def run(g, s, ref, trg, cmin, cmax, groups=None):
        frames = np.ascontiguousarray(g["frames"], np.float32); F, _, na = frames.shape; cells = _cells(g); mass = np.ascontiguousarray(s["mass"], np.float32)
        trg = np.ascontiguousarray(trg, np.int32); keep = np.zeros((F, 1024), np.uint32); tot = (C.c_ulonglong * F)()
        if groups is not None:
            ref = np.ascontiguousarray(np.concatenate(groups), np.int32); off = np.zeros(len(groups) + 1, np.uint32); off[1:] = np.cumsum([len(x) for x in groups])
            offp, ng, sym = off.ctypes.data_as(UP), len(groups), 0
        else:
            ref = np.ascontiguousarray(ref, np.int32); offp, ng, sym = None, 0, int(np.array_equal(ref, trg))
        rc = lib.emul_rdf(frames.ctypes.data_as(FP), 3 * na, na, F, cells, mass.ctypes.data_as(FP), ref.ctypes.data_as(IP), len(ref), offp, ng, trg.ctypes.data_as(IP), len(trg),
                          cmin, cmax, sym, 1 << 16, keep.ctypes.data_as(UP), tot)
        assert rc == 0
        return keep, np.array(list(tot), np.uint64)

    def check(keep, tot, g, key):
        for f in range(keep.shape[0]):
            want = g[f"{key}__pf"][f, :1024]
            assert np.array_equal(keep[f].astype(np.float32), want), (key, f)
            assert tot[f] == int(want.sum()), (key, f)

    g = load_golden("water6.npz"); s = golden_system(g); o = sel_element(s, 8); h = sel_element(s, 1); co = s["comp_off"]
    check(*run(g, s, o, o, 0.0, 6.0), g, "r")
    check(*run(g, s, o, h, 1.5, 6.0), g, "rh")
    check(*run(g, s, None, o, 0.0, 5.0, groups=[np.arange(co[r], co[r + 1], dtype=np.int32) for r in range(20)]), g, "rc")
    g = load_golden("tric6.npz"); s = golden_system(g); o = sel_element(s, 8); h = sel_element(s, 1); co = s["comp_off"]
    check(*run(g, s, o, o, 0.0, 6.0), g, "rt")
    check(*run(g, s, o, h, 2.0, 7.0), g, "rth")
    check(*run(g, s, None, h, 0.0, 5.0, groups=[np.arange(co[r], co[r + 1], dtype=np.int32) for r in range(30)]), g, "rtc")
