"""Config 4 (synthetic membrane, 994 656 atoms) on one B200: device-resident frames, lipid-tail rdf + density_z over all atoms.
Prints one JSON line with the per-kernel CUDA-event times (plan kernel timing) and the achieved HBM rate of k_density.
Run on the GPU box:  python profiles/config4_membrane.py [--frames 24]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viamd_b200 as vb

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=24); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--script", default="rt = rdf(name('C2*'), name('C2*'), 12.0); dz = density_z(all);")
a = ap.parse_args()
nl, nwxy, nwz, seed, F = 38, 100, 48, 4321, a.frames
base, whole, mol, L3 = vb.synth_membrane_base(nl, nwxy, nwz, seed); na = base.shape[1]
sysm = vb.membrane_system(nl, nwxy, nwz)
props = vb.compile_script(a.script, sysm)
d_base = vb.device_alloc(0, base.nbytes); vb.memcpy_h2d(0, d_base, base.ctypes.data, base.nbytes)
d_mol = vb.device_alloc(0, mol.nbytes); vb.memcpy_h2d(0, d_mol, mol.ctypes.data, mol.nbytes)
d_fr = vb.device_alloc(0, F * 3 * na * 4)
vb.synth_membrane_frames_device(0, nl, nwxy, nwz, seed, d_base, d_mol, 0, F, d_fr, 3 * na, na)
cell = vb.UnitCell.from_basis(*L3)
plan = vb.Plan(sysm, props, F, batch_frames=F)
plan.set_initial_frame(*vb.synth_membrane_frames_host(nl, nwxy, nwz, seed, base, mol, 0, 1)[0], cell)
for _ in range(3): plan.eval_device_frames(d_fr, 3 * na, na, cell, 0, F)
plan.sync(); plan.clear(); plan.enable_kernel_timing(True)
plan.timer_begin()
for _ in range(a.reps): plan.eval_device_frames(d_fr, 3 * na, na, cell, 0, F)
ms = plan.timer_end()
out = {"workload": "membrane 994656 atoms", "script": a.script, "frames": F * a.reps, "ms_total": ms, "frames_per_s": F * a.reps / ms * 1e3}
for k in ("k_rdf_pairs", "k_sdf", "k_density"):
    t, n = plan.kernel_time_ms(k)
    if n: out[k] = {"ms_per_launch": t / n, "launches": n}
if "k_density" in out:
    dens = [p for p in props if vb.OP_DENSITY_X <= p.op <= vb.OP_DENSITY_Z]
    alg = np.mean([len(p.idx[0]) for p in dens]) * 12.0 * F     # coordinate + index + mass, 4 B each, per selected atom per frame, per launch
    out["k_density"]["algorithmic_GBps"] = alg / (out["k_density"]["ms_per_launch"] * 1e-3) / 1e9
print(json.dumps(out))
