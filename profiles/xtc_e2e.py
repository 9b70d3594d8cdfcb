"""End-to-end rate of the RDF+SDF script fed with XTC-compressed host frames (decode on the device) next to raw float host frames.
Input: oracle/_ref/xtc_water32_16.xtc (profiles/make_xtc_sample.py), tiled to --frames frames in pinned host memory.
Prints one JSON line. Run on the GPU box: python profiles/xtc_e2e.py"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viamd_b200 as vb

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=2368); ap.add_argument("--steps", type=int, default=4); ap.add_argument("--streams", type=int, default=0); ap.add_argument("--script", default="")
a = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
blob = np.fromfile(os.path.join(ROOT, "oracle", "_ref", "xtc_water32_16.xtc"), np.uint8)
offs, na = vb.xtc_frame_offsets(blob); nf = len(offs) - 1
reps = (a.frames + nf - 1) // nf; F = reps * nf
xyz, cells, _, _ = vb.xtc_decode_frames(blob, offs, na)
# tiled copies in pinned memory
hb = vb.host_alloc_pinned(blob.size * reps)
for r in range(reps): C.memmove(hb + r * blob.size, blob.ctypes.data, blob.size)
toffs = np.concatenate([offs[:-1] + np.uint64(r * blob.size) for r in range(reps)] + [np.array([reps * blob.size], np.uint64)])
hr = vb.host_alloc_pinned(F * 3 * na * 4)
for r in range(reps): C.memmove(hr + r * xyz.nbytes, xyz.ctypes.data, xyz.nbytes)
sysm = vb.water_system(32)
SCRIPT = a.script or "r = rdf(element('O'), element('O'), 10.0); v = sdf(residue(1:1000), element('O'), 10.0);"
out = {"frames_per_step": F, "xtc_bytes_per_frame": blob.size / nf, "raw_bytes_per_frame": 12 * na}
for mode in ("xtc", "raw"):
    plan = vb.Plan(sysm, vb.compile_script(SCRIPT, sysm), F * (a.steps + 2), num_streams=a.streams)
    plan.set_initial_frame(*xyz[0], cells[0])
    def step(i):
        if mode == "xtc": plan.eval_xtc_ptr(hb, toffs, i * F)
        else: plan.eval_host_ptr(hr, 3 * na, na, cells[0], i * F, F)
        return sum(float(plan.property_data(pp.name).values[0]) for pp in plan.properties)
    for i in range(2): step(i)
    t0 = time.perf_counter()
    for i in range(2, 2 + a.steps): step(i)
    dt = time.perf_counter() - t0
    out[mode + "_frames_per_s"] = a.steps * F / dt
    out[mode + "_checks"] = [float(np.float64(plan.property_data(pp.name).values).sum()) for pp in plan.properties]
    plan.close()
print(json.dumps(out))
