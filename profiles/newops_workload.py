"""The procedures added after the last profiled commit (rmsd, distance_pair, com, plane, coord rows, `in` contexts, count(within()),
rdf(within())) on the bench's water box (n=32: 98 304 atoms), device-resident frames, ONE script per run so that a launch list attributes
every kernel. Prints one JSON line per script: frames/s of the whole plan (CUDA events around the evaluation) and the launch count.

They have parity on the B200 (profiles/r03a_newops_gpu_tests.log) and no timing yet; this is the workload for the first GPU call of round 2:

    python profiles/newops_workload.py                                   # timings, all scripts
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/newops_launches.csv \
        python profiles/newops_workload.py --reps 1 --warm 1             # launch list (per-kernel share; never a bench value)
    ncu --set full --clock-control none --import-source on -k regex:k_within_mark -c 1 -o gpurun_out/k_within_mark \
        python profiles/newops_workload.py --only cw --reps 1 --warm 0

`--n 6` is small enough for tests/emul (python tests/emul/run_under_emulation.py is not needed: pass --emulated)."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import viamd_b200 as vb

SCRIPTS = {   # name -> script; selections sized like VIAMD's typical use (one molecule / a residue range / a solvation shell)
    "rmsd": "p = rmsd(residue(1:1000));",
    "dpair": "p = distance_pair(residue(1:40), residue(41:80));",
    "com_plane": "c = com(atom(1:3000)); q = plane(atom(1:3000));",
    "coord": "z = coord_z(element('O'));",
    "ctx": "d = distance(1, 2) in residue(1:10000);",
    "cw": "n = count(within(5.0, residue(1:100)));",
    "cw_and": "n = count(element('O') and within(3.5, residue(1:100)));",
    "rdf_within": "r = rdf(within(5.0, residue(1:100)), element('O'), 10.0);",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32); ap.add_argument("--frames", type=int, default=148); ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warm", type=int, default=2); ap.add_argument("--only", default=""); ap.add_argument("--emulated", action="store_true")
    a = ap.parse_args()
    if a.emulated:   # CPU check of this script itself (test infrastructure, swapped in inside this process only)
        sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
        import build_emul, viamd_b200.api as api
        api.LIB_PATH = build_emul.build_library(); api._lib = None
    n, seed, F = a.n, 1234, a.frames
    base, L = vb.synth_water_base(n, seed); na = 3 * n ** 3
    sysm = vb.water_system(n); cell = vb.UnitCell.from_basis(L, L, L)
    d_base = vb.device_alloc(0, base.nbytes); vb.memcpy_h2d(0, d_base, base.ctypes.data, base.nbytes)
    d_fr = vb.device_alloc(0, F * 3 * na * 4)
    vb.synth_water_frames_device(0, n, seed, d_base, 0, F, d_fr, 3 * na, na)
    first = vb.synth_water_frames_host(n, seed, base, 0, 1)[0]
    nres = n ** 3
    for name, src in SCRIPTS.items():
        if a.only and name not in a.only.split(","): continue
        if nres < 10000: src = src.replace("1:10000", f"1:{nres}").replace("1:1000", f"1:{min(1000, nres)}").replace("1:3000", f"1:{min(3000, 3 * nres)}")
        props = vb.compile_script(src, sysm)
        plan = vb.Plan(sysm, props, F, batch_frames=F)
        plan.set_initial_frame(*first, cell)
        for _ in range(a.warm): plan.eval_device_frames(d_fr, 3 * na, na, cell, 0, F)
        plan.sync(); plan.clear(); vb.launch_count(reset=True)
        plan.timer_begin()
        for _ in range(a.reps): plan.eval_device_frames(d_fr, 3 * na, na, cell, 0, F)
        ms = plan.timer_end()
        vals = plan.property_data(props[0].name).values
        print(json.dumps({"script": src, "atoms": na, "frames": F * a.reps, "ms_per_batch": ms / a.reps, "frames_per_s": (F * a.reps / ms * 1e3) if ms > 0 else None,
                          "launches": vb.launch_count(), "checksum": float(np.nansum(np.asarray(vals, np.float64)))}), flush=True)
        plan.close()
    vb.device_free(0, d_fr); vb.device_free(0, d_base)


if __name__ == "__main__":
    main()
