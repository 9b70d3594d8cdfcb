"""ONE process, several GPUs, through the C ABI (mdgpu_plan_options_t.num_devices; SURVEY.md 8(e) inside libmdgpu, the shape VIAMD needs):
pinned host frames of the bench workload -> mdgpu_eval_host_frames on a multi-device plan (contiguous frame blocks per device, one host thread
per device) -> mdgpu_plan_sync (NCCL reduce onto device 0) -> results. Checks that the integer accumulators equal the single-device plan's and
prints frames/s end to end for 1..N devices plus the time of the exchange step.
    gpurun --gpus 2 -- python profiles/multi_device_e2e.py --devices 2"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import viamd_b200 as vb

ap = argparse.ArgumentParser(); ap.add_argument("--devices", type=int, default=2); ap.add_argument("--frames", type=int, default=2368); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
n, seed = 32, 1234
base, L = vb.synth_water_base(n, seed); na = base.shape[1]; fstride = 3 * na
sysm = vb.water_system(n); cell = vb.UnitCell.from_basis(L, L, L)
script = "r = rdf(element('O'), element('O'), 10.0); v = sdf(residue(1:1000), element('O'), 10.0); d = distance(1, 10);"
F = a.frames
d_base = vb.device_alloc(0, base.nbytes); vb.memcpy_h2d(0, d_base, base.ctypes.data, base.nbytes)
d_fr = vb.device_alloc(0, F * fstride * 4); vb.synth_water_frames_device(0, n, seed, d_base, 0, F, d_fr, fstride, na)
h = vb.host_alloc_pinned(F * fstride * 4); vb.memcpy_d2h(0, h, d_fr, F * fstride * 4); vb.device_free(0, d_fr)
f0 = vb.synth_water_frames_host(n, seed, base, 0, 1)
out = {"workload": f"water n={n}, {F} frames per pass, script: {script}", "runs": []}
ref = None
for nd in sorted({1, a.devices}):
    plan = vb.Plan(sysm, vb.compile_script(script, sysm), F, devices=list(range(nd)))
    plan.set_initial_frame(*f0[0], cell)
    plan.eval_host_ptr(h, fstride, na, cell, 0, F); plan.sync()          # warm-up (allocations, NCCL communicator)
    best = 1e9
    for _ in range(a.reps):
        plan.clear(); t0 = time.perf_counter()
        plan.eval_host_ptr(h, fstride, na, cell, 0, F); plan.sync(); best = min(best, time.perf_counter() - t0)
    res = (plan.counts("r"), plan.counts("v"), plan.property_data("d").values.copy(), plan.frame_mask().copy(), plan.property_data("r").weights.copy())
    ex_ms, ex_n = plan.exchange_stats()
    if ref is None: ref = res
    same = all(np.array_equal(x, y) for x, y in zip(ref, res))
    out["runs"].append({"devices": nd, "frames_per_s_e2e": F / best, "s_per_pass": best, "exchange_ms": ex_ms, "exchanges": ex_n, "equal_to_single_device": bool(same),
                        "ingest_atoms_per_frame": plan.ingest_info()[0]})
    plan.close()
print(json.dumps(out))
sys.exit(0 if all(r["equal_to_single_device"] for r in out["runs"]) else 1)
