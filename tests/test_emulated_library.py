"""The whole library on the CPU: tests/emul/build/libmdgpu_emul.so is libmdgpu's own sources (plan.cu, every kernel file) compiled by g++ —
kernel launches turned into emul_launch, CUDA runtime calls served by tests/emul/fake_cudart.cpp — so the C ABI, the plan's host logic
(batching, slots, scratch sizing, result folds) and the kernels run together without a GPU. Here: the tests of the device paths written after
the GPU budget was spent (tests/test_zz_gpu_new_ops.py, all of them) and a few of the GPU-validated parity tests as a check of the emulation.
The complete GPU suite passes this way too (28 tests, ~25 min): `python tests/emul/run_under_emulation.py tests/test_gpu_parity.py -m gpu`.

This is evidence about source logic, not a substitute for the GPU run: launch limits, memory spaces and timing are not modelled."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))


@pytest.fixture(scope="module")
def emulated_library():
    import build_emul
    import viamd_b200.api as api
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH = build_emul.build_library(); api._lib = None
    yield api
    api.LIB_PATH, api._lib = saved


def _run_all(mod, names):
    for n in names:
        getattr(mod, n)()


def test_new_device_paths_through_the_c_abi(emulated_library):
    """rmsd, distance_pair + aggregates, com, plane, count(within()), their error paths: every test of tests/test_zz_gpu_new_ops.py."""
    import test_zz_gpu_new_ops as P
    import inspect
    names = [n for n in dir(P) if n.startswith("test_") and not inspect.signature(getattr(P, n)).parameters]   # the shim test (tmp_path) drives a binary linked to the real library
    for slow in ("test_within_min_max_form", "test_static_selection_and_within"):   # ~40 s each under emulation; they pass (run_under_emulation.py) and
        names.remove(slow)                                                              # their forms are in the shim test of this suite, against the reference itself
    assert len(names) >= 7
    _run_all(P, names)


def test_validated_paths_agree_under_emulation(emulated_library):
    """A slice of tests/test_gpu_parity.py (all of which have passed on a B200): rdf per-frame bins incl. batching and stream slots,
    centre-of-mass references, density + every temporal, error paths."""
    import test_gpu_parity as G
    _run_all(G, ["test_golden_water_rdf_per_frame_bitexact", "test_golden_water_rdf_com_references_bitexact", "test_golden_water_density_and_temporals",
                 "test_empty_and_error_paths"])


def _shard_worker(rank, world, port, q):
    """one rank of the frame-sharded evaluation: global plan, its shard at global frame offsets, the one exchange step over gloo"""
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(here, "emul")):
        sys.path.insert(0, p)
    import build_emul
    import viamd_b200.api as api
    api.LIB_PATH = build_emul.build_library(); api._lib = None
    import viamd_b200 as vb
    from viamd_b200 import dist as vdist
    from helpers import load_golden, golden_system, vb_system, vb_cell
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("water6.npz"); s = golden_system(g); F = g["frames"].shape[0]
    props = vb.compile_script("r = rdf(element('O'), element('O'), 6.0); d = distance(1,10); dp = distance_pair(atom(1:5), atom(20:30)); cw = count(within(4.0, residue(1)));", vb_system(s))
    plan = vb.Plan(vb_system(s), props, F)
    cells = [vb_cell(g["cells"][f], g["cell_flags"][f]) for f in range(F)]
    plan.set_initial_frame(*g["frames"][0], cells[0])
    beg, end = vdist.frame_shard(F, world, rank)
    plan.eval_host_frames(g["frames"][beg:end], cells[beg:end], beg)

    def host_view(ptr, n, typestr):   # the emulated library's "device" memory is host memory: wrap it in place
        dt = np.dtype(typestr); buf = (ctypes.c_char * (n * dt.itemsize)).from_address(ptr)
        return torch.from_numpy(np.frombuffer(buf, dtype=dt))
    vdist.allreduce_plan(plan, F, view=host_view)
    out = {k: plan.property_data(k).values.copy() for k in ("r", "d", "dp", "cw")}
    out["dp_mean"] = plan.aggregate("dp")["mean"]; out["d_minmax"] = np.array([plan.property_data("d").min_value, plan.property_data("d").max_value], np.float32)
    out["mask"] = plan.frame_mask()
    plan.close()
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_frame_shards_merge_through_the_exchange_step():
    """SURVEY 8(e) with world_size 2 over gloo, on the emulated library: every rank holds a plan over the GLOBAL frame range, evaluates its
    contiguous shard, and viamd_b200.dist.allreduce_plan — the function bench.py calls over NCCL — merges integer bins and the disjoint float
    rows of the temporals. Both ranks end with the reference's 4-frame results: rdf mean, distance, the distance_pair matrix with its
    per-frame aggregates, count(within()), min/max over all frames, a full frame mask."""
    import numpy as np
    import torch.multiprocessing as mp
    from helpers import load_golden
    ctx = mp.get_context("spawn"); q = ctx.Queue(); port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs: p.join(120)
    g = load_golden("water6.npz"); pz = load_golden("pairs6.npz")
    for r in (0, 1):
        o = res[r]
        np.testing.assert_allclose(o["r"][:1024], g["r__full"][:1024], rtol=1e-5, atol=1e-6)
        assert np.array_equal(o["d"], g["d__full"]) and np.array_equal(o["dp"], pz["w_dp__full"]) and np.array_equal(o["cw"], g["cw__full"])
        assert np.array_equal(o["dp_mean"], pz["w_dp__mean"]) and o["d_minmax"][0] == g["d__meta"][0] and o["d_minmax"][1] == g["d__meta"][1]
        assert o["mask"].all()
    assert np.array_equal(res[0]["r"], res[1]["r"])
