"""The plain-C oracle (oracle/md_oracle.c) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, strict build). Bit-exact for bins / voxels / weights / distances."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import load_golden, cell_from_row, dense_from_sparse, golden_system, sel_element


@pytest.fixture(scope="module")
def water():
    g = load_golden("water6.npz"); return g, golden_system(g)


@pytest.fixture(scope="module")
def ala():
    g = load_golden("ala50.npz"); return g, golden_system(g)


def test_rdf_water_bins_and_weights_bitexact(water):
    g, s = water; o = sel_element(s, 8); h = sel_element(s, 1)
    for f in range(g["frames"].shape[0]):
        x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f])
        bins, w, tot = O.rdf_frame(x, y, z, o, o, cell, 0.0, 6.0)
        assert np.array_equal(bins, g["r__pf"][f, :1024]) and np.array_equal(w, g["r__pf"][f, 1024:])
        assert tot == int(g["r__pf"][f, :1024].sum()) > 0
        bins, w, _ = O.rdf_frame(x, y, z, o, h, cell, 1.5, 6.0)   # min:max form
        assert np.array_equal(bins, g["rh__pf"][f, :1024]) and np.array_equal(w, g["rh__pf"][f, 1024:])


def test_sdf_water_voxels_bitexact(water):
    g, s = water; o = sel_element(s, 8)
    structs = np.arange(60, dtype=np.int32).reshape(20, 3)
    init = g["frames"][0]
    for f in range(g["frames"].shape[0]):
        x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f])
        vol, n = O.sdf_frame(x, y, z, init, s["mass"], structs, o, s["conn_off"], s["conn_idx"], cell, 5.0)
        ref = dense_from_sparse(g[f"v__pf{f}_idx"], g[f"v__pf{f}_val"])
        assert n == int(ref.sum()) > 0
        assert np.array_equal(vol, ref)


def test_density_water_bitexact(water):
    g, s = water; o = sel_element(s, 8)
    c0 = cell_from_row(g["cells"][0], g["cell_flags"][0])
    for f in range(g["frames"].shape[0]):
        x, y, z = g["frames"][f]
        for axis, key in ((2, "dz"), (0, "dx")):
            bins, w = O.density_frame(x, y, z, s["mass"], o, c0, axis)
            assert np.array_equal(bins, g[f"{key}__pf"][f, :1024]) and np.array_equal(w, g[f"{key}__pf"][f, 1024:])


def test_temporals_water(water):
    g, s = water
    for f in range(g["frames"].shape[0]):
        x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f])
        assert np.float32(O.distance(x, y, z, 0, 9, cell)) == g["d__full"][f]
        assert np.float32(O.angle(x, y, z, 0, 1, 2)) == g["a__full"][f]            # same libm here -> exact
        assert np.float32(O.dihedral(x, y, z, 0, 3, 6, 9, cell)) == g["t__full"][f]


def test_config1_distance_1ala(ala):
    """BASELINE config 1: d = distance(1,10) on datasets/1ALA-500.pdb. values[0] = 2.770258, sum(500) = 1493.846763."""
    g, s = ala
    assert abs(float(g["d500__full"][0]) - 2.770258) < 1e-6 and abs(float(g["d500__full"].astype(np.float64).sum()) - 1493.846763) < 2e-3
    for f in range(g["frames"].shape[0]):
        x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f])
        assert np.float32(O.distance(x, y, z, 0, 9, cell)) == g["d__full"][f] == g["d500__full"][f]
        assert np.float32(O.angle(x, y, z, 0, 4, 8)) == g["a__full"][f]
        assert np.float32(O.dihedral(x, y, z, 4, 6, 8, 14, cell)) == g["t__full"][f]


def test_rdf_and_density_1ala(ala):
    g, s = ala; c = sel_element(s, 6); o = sel_element(s, 8)
    c0 = cell_from_row(g["cells"][0], g["cell_flags"][0])
    for f in range(0, g["frames"].shape[0], 7):
        x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f])
        bins, w, _ = O.rdf_frame(x, y, z, c, o, cell, 0.0, 10.0)
        assert np.array_equal(bins, g["rc__pf"][f, :1024]) and np.array_equal(w, g["rc__pf"][f, 1024:])
        b, w = O.density_frame(x, y, z, s["mass"], c, c0, 2)
        assert np.array_equal(b, g["dz__pf"][f, :1024])


def test_membrane_config4_shape():
    """BASELINE config 4 at reduced size: lipid-tail rdf() and density_z/x profiles, bit-exact vs the reference."""
    g = load_golden("membrane6.npz"); s = golden_system(g)
    names = np.array(s["names"]); c2 = np.nonzero(np.char.startswith(names, "C2"))[0].astype(np.int32); allat = np.arange(len(names), dtype=np.int32)
    c0 = cell_from_row(g["cells"][0], g["cell_flags"][0])
    for f in range(g["frames"].shape[0]):
        x, y, z = g["frames"][f]; cell = cell_from_row(g["cells"][f], g["cell_flags"][f])
        bins, w, tot = O.rdf_frame(x, y, z, c2, c2, cell, 0.0, 12.0)
        assert np.array_equal(bins, g["rt__pf"][f, :1024]) and np.array_equal(w, g["rt__pf"][f, 1024:]) and tot > 0
        for key, sel, axis in (("dz", c2, 2), ("dall", allat, 2), ("dxall", allat, 0)):
            b, _ = O.density_frame(x, y, z, s["mass"], sel, c0, axis)
            assert np.array_equal(b, g[f"{key}__pf"][f, :1024]), key


def test_svd3_reconstructs():
    rng = np.random.default_rng(3)
    import ctypes as C
    for _ in range(20):
        A = rng.normal(size=(3, 3)).astype(np.float32)
        U = np.zeros((3, 3), np.float32); S = np.zeros((3, 3), np.float32); V = np.zeros((3, 3), np.float32)
        O.lib().mdo_svd3(A.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
        assert np.allclose(U @ S @ V.T, A, atol=2e-5)
        assert np.allclose(U @ U.T, np.eye(3), atol=2e-5) and np.allclose(V @ V.T, np.eye(3), atol=2e-5)
