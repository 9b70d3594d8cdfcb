"""CPU-side checks: the C ABI library loads and exports every symbol include/mdgpu.h declares, it fails loudly without a
GPU (no CPU fallback), script lowering, synthetic-data determinism, and the world_size-2 (gloo) frame-shard merge."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vb():
    from viamd_b200 import build
    build.build()
    import viamd_b200 as vb
    return vb


def test_abi_exports_every_declared_symbol(vb):
    hdr = open(os.path.join(ROOT, "include", "mdgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mdgpu_[a-z0-9_]+)\s*\(", hdr))
    names = {n for n in names if not n.endswith("_t")}
    assert len(names) >= 30
    lib = vb.lib()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    # layout of the structs shared with the reference
    assert C.sizeof(vb.UnitCell) == 56                     # md_unitcell_t: 6 doubles + flags (+pad)
    from viamd_b200.api import FrameHeader
    assert C.sizeof(FrameHeader) == 8 + 8 + 8 + 56         # md_trajectory_frame_header_t


def test_no_cpu_fallback(vb):
    if vb.device_count() > 0:
        pytest.skip("GPU present")
    s = vb.water_system(2)
    o = np.arange(0, 24, 3, dtype=np.int32)
    with pytest.raises(vb.MdgpuError, match="no CUDA device"):
        vb.Plan(s, [vb.rdf("r", o, o, 5.0)], 1)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "viamd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "md_oracle" not in src and "oracle_lib" not in src and "liboracle" not in src, f
                # nor the CPU emulation the tests use (tests/emul): the product has no switch that could load it
                assert "libmdgpu_emul" not in src and "fake_cudart" not in src and "emul_launch" not in src and "MDG_HOST_EMULATION" not in src, f
    for f in ("bench.py", "__graft_entry__.py"):
        src = open(os.path.join(ROOT, f)).read()
        assert "libmdgpu_emul" not in src and "build_emul" not in src and "tests/emul" not in src and "tests.emul" not in src, f


def test_script_lowering(vb):
    s = vb.water_system(4)
    props = vb.compile_script("r = rdf(element('O'), element('O'), 10.0);\nv = sdf(residue(1:10), element('O'), 5.0);\n"
                              "dz = density_z(name('HW*')); d = distance(1,10); rr = rdf(element('O'), not element('O'), 2.0:8.0);", s)
    r, v, dz, d, rr = props
    assert r.op == vb.OP_RDF and np.array_equal(r.idx[0], np.arange(0, 192, 3)) and r.cutoff_max == 10.0 and r.cutoff_min == 0.0
    assert v.op == vb.OP_SDF and v.num_structures == 10 and v.structure_size == 3 and np.array_equal(v.idx[0], np.arange(30))
    assert dz.op == vb.OP_DENSITY_Z and len(dz.idx[0]) == 128
    assert d.op == vb.OP_DISTANCE and d.idx[0][0] == 0 and d.idx[1][0] == 9          # 1-based script indices
    assert rr.cutoff_min == 2.0 and rr.cutoff_max == 8.0 and len(rr.idx[1]) == 128
    with pytest.raises(vb.ScriptError):
        vb.compile_script("x = shape_weights(all);", s)
    dpg = vb.compile_script("dpg = distance_pair(residue(1:3), residue(5));", s)[0]   # array of selections x one selection
    assert dpg.op == vb.OP_DISTANCE_PAIR and dpg.num_structures == 3 and list(dpg.structure_offsets) == [0, 3, 6, 9] and dpg.structure_offsets_b is None and list(dpg.idx[1]) == [12, 13, 14]
    dp = vb.compile_script("dp = distance_pair(residue(1), element('O'));", s)[0]
    assert dp.op == vb.OP_DISTANCE_PAIR and list(dp.idx[0]) == [0, 1, 2] and len(dp.idx[1]) == 64
    c, ci, pl = vb.compile_script("c = com(residue(2)); ci = com(7); pl = plane(atom(1:9));", s)
    assert c.op == vb.OP_COM and c.com_args == 1 and list(c.idx[0]) == [3, 4, 5] and ci.com_args == 0 and list(ci.idx[0]) == [6]
    assert pl.op == vb.OP_PLANE and list(pl.idx[0]) == list(range(9))
    plg = vb.compile_script("x = plane(residue(1:3));", s)[0]   # the plane through the three residues' centres of mass
    assert plg.op == vb.OP_PLANE and plg.num_structures == 3 and list(plg.structure_offsets) == [0, 3, 6, 9]
    cxg, dmg = vb.compile_script("cxg = coord_x(residue(1:3)); dmg = distance_min(residue(1:2), element('O'));", s)   # one centre of mass per selection of the array
    assert cxg.num_structures == 3 and list(cxg.structure_offsets) == [0, 3, 6, 9] and list(cxg.idx[0]) == list(range(9))
    assert dmg.op == vb.OP_DISTANCE_MIN and dmg.num_structures == 2 and list(dmg.structure_offsets) == [0, 3, 6] and dmg.structure_offsets_b is None and len(dmg.idx[1]) == 64
    # an ARRAY of selections as one position argument: the centre of the selections' centres for com / angle / dihedral (arg_offsets),
    # the union for distance (FLAG_FLATTEN, md_script_functions.inl:680) — a com(...) inside distance included
    ca, an, df, dcm = vb.compile_script("ca = com(residue(1:3)); an = angle(residue(1:2), 7, com(residue(3:5))); df = distance(residue(1:2), 5); dcm = distance(com(residue(1:3)), residue(4));", s)
    assert ca.op == vb.OP_COM and list(ca.idx[0]) == list(range(9)) and list(ca.arg_offsets[0]) == [0, 3, 6, 9] and ca.com_args == 1
    assert sorted(an.arg_offsets) == [0, 2] and list(an.arg_offsets[2]) == [0, 3, 6, 9] and list(an.idx[2]) == list(range(6, 15)) and list(an.idx[1]) == [6] and an.com_args == 5
    assert not df.arg_offsets and list(df.idx[0]) == list(range(6)) and df.com_args == 1
    assert not dcm.arg_offsets and list(dcm.idx[0]) == list(range(9)) and list(dcm.idx[1]) == [9, 10, 11] and dcm.com_args == 3
    rwp = vb.compile_script("rw = rdf(within(4.0, residue(2)), element('O'), 2.0:6.0);", s)[0]
    assert rwp.op == vb.OP_RDF and rwp.ref_within == 4.0 and list(rwp.idx[0]) == [3, 4, 5] and rwp.cutoff_min == 2.0 and rwp.num_structures == 0
    inc = vb.compile_script("x = dihedral(1,2,3,1) in residue(2:4);", s)[0]
    assert inc.op == vb.OP_DIHEDRAL and inc.num_structures == 3 and [list(i) for i in inc.idx] == [[3, 6, 9], [4, 7, 10], [5, 8, 11], [3, 6, 9]]
    with pytest.raises(vb.ScriptError):
        vb.compile_script("x = distance(residue(1), 2) in residue(2:4);", s)
    dcm = vb.compile_script("x = distance(com(atom(1:6)), 10);", s)[0]
    assert dcm.op == vb.OP_DISTANCE and dcm.com_args == 1 and list(dcm.idx[0]) == [0, 1, 2, 3, 4, 5] and list(dcm.idx[1]) == [9]
    cwo, rwo = vb.compile_script("cwo = count(element('O') and within(2.0:4.0, residue(2))); rwo = rdf(within(5.0, residue(1)) and not element('O'), element('O'), 6.0);", s)
    assert cwo.op == vb.OP_WITHIN_COUNT and cwo.com_args == 1 and (cwo.cutoff_min, cwo.cutoff_max) == (2.0, 4.0) and list(cwo.idx[0]) == [3, 4, 5] and np.array_equal(cwo.idx[2], np.arange(0, 192, 3))
    assert rwo.op == vb.OP_RDF and rwo.ref_within == 5.0 and rwo.com_args == 1 and len(rwo.idx[2]) == 128 and list(rwo.idx[0]) == [0, 1, 2]
    for src in ("x = count(element('O') or within(4.0, residue(1)));", "x = count(not within(4.0, residue(1)));"):
        with pytest.raises(vb.ScriptError):
            vb.compile_script(src, s)
    cw = vb.compile_script("cw = count(within(4.5, residue(2)));", s)[0]
    assert cw.op == vb.OP_WITHIN_COUNT and cw.cutoff_max == 4.5 and list(cw.idx[0]) == [3, 4, 5]
    with pytest.raises(vb.ScriptError):
        vb.compile_script("x = count(element('O'));", s)
    rm = vb.compile_script("rm = rmsd(residue(2:4));", s)[0]                           # array of selections -> their union
    assert rm.op == vb.OP_RMSD and list(rm.idx[0]) == list(range(3, 12))
    # array-of-selections reference -> centre-of-mass groups with offsets; selection arguments of the temporals -> com_args; pair minimum
    rc, dc, dm, dmin = vb.compile_script("rc = rdf(residue(2:5), element('O'), 5.0); dc = distance(residue(1), residue(3)); "
                                         "dm = distance(atom(1:6), 10); dmin = distance_min(residue(1), element('H'));", s)
    assert rc.op == vb.OP_RDF and rc.num_structures == 4 and list(rc.structure_offsets) == [0, 3, 6, 9, 12] and np.array_equal(rc.idx[0], np.arange(3, 15))
    assert dc.op == vb.OP_DISTANCE and dc.com_args == 3 and list(dc.idx[0]) == [0, 1, 2] and list(dc.idx[1]) == [6, 7, 8]
    assert dm.com_args == 1 and list(dm.idx[0]) == [0, 1, 2, 3, 4, 5] and list(dm.idx[1]) == [9]
    assert dmin.op == vb.OP_DISTANCE_MIN and list(dmin.idx[0]) == [0, 1, 2] and len(dmin.idx[1]) == 128
    single = vb.compile_script("r1 = rdf(residue(2), element('O'), 5.0);", s)[0]       # one selection: plain atom references, no groups
    assert single.num_structures == 0 and list(single.idx[0]) == [3, 4, 5]


def test_synth_determinism_and_tool_agreement(vb, tmp_path):
    base, L = vb.synth_water_base(5, 99)
    fr = vb.synth_water_frames_host(5, 99, base, 3, 2)
    fr2 = vb.synth_water_frames_host(5, 99, base, 0, 5)
    assert np.array_equal(fr, fr2[3:5]) and fr.min() >= 0 and fr.max() < L
    # per-molecule rigid displacement is bounded by 510 * 2^-9 A wherever no wrap happened
    d = fr2[1] - base
    inside = np.abs(d) < 2.0
    assert inside.mean() > 0.9 and np.abs(d[inside]).max() <= 510 / 512 + 1e-4
    tool = os.path.join(ROOT, "oracle", "build", "synth_tool")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    raw = str(tmp_path / "w.raw")
    subprocess.check_call([tool, "water-raw", "5", "99", "5", raw])
    import refio
    frames, cells, flags = refio.read_raw_traj(raw)
    assert np.array_equal(frames, fr2) and np.allclose(cells[0][[0, 3, 5]], L) and flags[0] == 29


def test_frame_shard_partition():
    from viamd_b200.dist import frame_shard
    for F in (1, 7, 1000, 10 ** 6):
        for G in (1, 2, 4, 8):
            blocks = [frame_shard(F, G, g) for g in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == F
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from viamd_b200.dist import frame_shard, allreduce_counts_
    import oracle_lib as O
    from helpers import load_golden, golden_system, sel_element, cell_from_row
    g = load_golden("water6.npz"); s = golden_system(g); o = sel_element(s, 8)
    F = g["frames"].shape[0]
    beg, end = frame_shard(F, world, rank)
    local = np.zeros(1024, np.int64)
    for f in range(beg, end):   # stand-in for the GPU evaluation of this rank's frame block (same integer bins)
        bins, _, _ = O.rdf_frame(*g["frames"][f], o, o, cell_from_row(g["cells"][f], g["cell_flags"][f]), 0.0, 6.0)
        local += bins.astype(np.int64)
    t = torch.from_numpy(local)
    allreduce_counts_(t)
    q.put((rank, t.numpy().copy()))
    dist.destroy_process_group()


def test_world_size_2_shard_and_allreduce_gloo():
    import torch.multiprocessing as mp
    from helpers import load_golden
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(60)
    g = load_golden("water6.npz")
    want = g["r__pf"][:, :1024].astype(np.int64).sum(axis=0)
    assert np.array_equal(res[0], want) and np.array_equal(res[1], want)
    # mean after the reduce is independent of the number of ranks and matches the reference's 4-frame average
    np.testing.assert_allclose((want / 4).astype(np.float32), g["r__full"][:1024], rtol=1e-5, atol=1e-6)


def test_rdf_weights_python_matches_reference():
    from viamd_b200.dist import rdf_weights
    from helpers import load_golden
    g = load_golden("water6.npz")
    for key, lo, hi in (("r", 0.0, 6.0), ("rh", 1.5, 6.0)):
        tot = int(g[f"{key}__pf"][3, :1024].sum())
        assert np.array_equal(rdf_weights(tot, lo, hi), g[f"{key}__pf"][3, 1024:])


def test_xtc_frame_offsets_host_side(vb):
    """mdgpu_xtc_frame_offsets (md_xtc_read_frame_offsets_and_times md_xtc.c:436-570) runs on the host: the product's scan of the golden XTC
    byte streams agrees with the oracle's and ends at the file size; a truncated file yields the complete frames only; garbage is refused."""
    import numpy as np
    import oracle_lib as O
    from helpers import load_golden
    g = load_golden("xtc_cases.npz")
    for case in ("water6", "tric6", "water16", "small5", "wide12", "huge12", "lowprec"):
        blob = g[case + "__xtc"]
        offs, na = vb.xtc_frame_offsets(blob)
        assert na == int(g[case + "__na"]) and int(offs[-1]) == blob.size
        assert np.array_equal(offs.astype(np.int64), O.xtc_frame_offsets(blob))
        assert len(offs) - 1 == len(g[case + "__cells"])
    blob = g["water6__xtc"]; full, _ = vb.xtc_frame_offsets(blob)
    cut, _ = vb.xtc_frame_offsets(blob[: int(full[2]) + 100])          # third frame incomplete
    assert list(cut) == list(full[:3])
    with pytest.raises(vb.MdgpuError):
        vb.xtc_frame_offsets(np.zeros(200, np.uint8))


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` prints exactly one JSON line on stdout with the contract's keys (runs the reference CPU harness when
    oracle/_ref is present, the oracle port otherwise): a bounded sample, so this stays a few seconds."""
    import json, subprocess, sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s" and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_frame_geometry_matches_the_oracle(vb):
    """The product's per-frame cell-grid geometry (compute_frame_geom: the code its device kernels run, evaluated on the host through
    mdgpu_debug_frame_geom) against the oracle's md_spatial_acc_init restatement on random cells: orthorhombic, anisotropic, one or more
    axes non-periodic (grid fitted to the points' bounding box), triclinic. Grid dimensions, neighbour reach, metric and r2 bit for bit."""
    import numpy as np
    import oracle_lib as O
    L = vb.lib()
    L.mdgpu_debug_frame_geom.argtypes = [C.POINTER(vb.UnitCell), C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    for case in range(300):
        ext = rng.uniform(8.0, 120.0, 3); kind = rng.integers(0, 4)
        flags = vb.CELL_ORTHO | vb.CELL_PBC_ALL; xy = xz = yz = 0.0
        if kind == 1: flags &= ~int(rng.choice([vb.CELL_PBC_X, vb.CELL_PBC_Y, vb.CELL_PBC_Z, vb.CELL_PBC_X | vb.CELL_PBC_Z, vb.CELL_PBC_ALL]))
        if kind == 2: flags = vb.CELL_TRICLINIC | vb.CELL_PBC_ALL; xy, xz, yz = rng.uniform(-0.3, 0.3, 3) * ext.min()
        cutoff = float(rng.uniform(2.0, 0.7 * ext.min())); cell_ext = cutoff if rng.random() < 0.7 else float(rng.uniform(3.0, 12.0))
        n = 200
        pts = (rng.random((3, n)) * ext[:, None] * rng.uniform(0.5, 1.0) + rng.uniform(-5, 5, 3)[:, None]).astype(np.float32)
        cell = vb.UnitCell(float(ext[0]), float(xy), float(xz), float(ext[1]), float(yz), float(ext[2]), int(flags))
        ocell = O.UnitCell.from_params(ext[0], xy, xz, ext[1], yz, ext[2], int(flags))
        # the reference starts its bounding box from {0} (md_spatial_acc.c:204): the origin is always inside; k_aabb does the same on the device
        aabb = np.concatenate([np.minimum(pts.min(axis=1), 0.0), np.maximum(pts.max(axis=1), 0.0)]).astype(np.float32)
        gi = np.zeros(13, np.int32); gf = np.zeros(7, np.float32)
        assert L.mdgpu_debug_frame_geom(C.byref(cell), cell_ext, cutoff, aabb.ctypes.data, gi.ctypes.data, gf.ctypes.data) == 0
        oi = np.zeros(7, np.int32); of = np.zeros(10, np.float32)
        x, y, z = (np.ascontiguousarray(p) for p in pts)
        O.lib().mdo_debug_geom(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.byref(ocell),
                               C.c_double(cell_ext), C.c_double(cutoff), oi.ctypes.data_as(C.c_void_p), of.ctypes.data_as(C.c_void_p))
        tag = (case, int(flags), cutoff, cell_ext, list(ext))
        assert list(gi[:3]) == list(oi[:3]), ("cdim", tag, gi[:3], oi[:3])
        assert list(gi[3:6]) == list(oi[3:6]), ("ncell", tag, gi[3:6], oi[3:6])
        assert np.array_equal(gf[:7], of[:7]), ("metric/r2", tag, gf, of[:7])
        assert (gi[12] > 0) == all(2 * int(v) + 1 <= 5 for v in oi[3:6]), ("valid", tag, gi[12], oi[3:6])


def test_host_fold_of_multi_valued_temporals_matches_the_reference(vb):
    """mdgpu_plan_sync's per-frame fold (min / max / mean / population variance, two float passes) on the reference's distance_pair() rows:
    equal to the aggregates the reference stored (pairs6.npz)."""
    from helpers import load_golden
    L = vb.lib(); L.mdgpu_debug_aggregate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    p = load_golden("pairs6.npz")
    for k in ("w_dp", "w_dpo", "t_dp", "t_dpo"):
        F, n = (int(v) for v in p[k + "__dim"][:2]); vals = np.ascontiguousarray(p[k + "__full"], np.float32)
        for f in range(F):
            row = np.ascontiguousarray(vals[f * n:(f + 1) * n]); out = np.zeros(4, np.float32)
            assert L.mdgpu_debug_aggregate(row.ctypes.data, n, out.ctypes.data) == 0
            assert out[0] == p[k + "__ext"][f, 0] and out[1] == p[k + "__ext"][f, 1] and out[2] == p[k + "__mean"][f] and out[3] == p[k + "__var"][f], (k, f)
