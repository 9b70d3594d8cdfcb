"""BASELINE configs[4] ("config 5") at N = 1: the rdf + sdf script over 10^6 device-generated frames of the 98 304-atom water box on ONE GPU — the
single-GPU point of the frame-sharded job (with G GPUs each rank evaluates 10^6 / G of these frames and one exchange step follows: bench.py / dist.py).
The trajectory (1.18 TB of floats) never exists as a whole: chunks of frames are generated on the device (k_synth_frames, the bench's generator, frame
index = global index) into a buffer in HBM and evaluated from there. Prints one JSON line.  Run on the GPU box: python profiles/config5_million_frames.py"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import viamd_b200 as vb
import bench as B

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=1_000_000); ap.add_argument("--chunk", type=int, default=148 * 128)
a = ap.parse_args()
dev = 0; cfg = B.CONFIGS["bench"]
chunk = a.chunk; nchunks = (a.frames + chunk - 1) // chunk; F = nchunks * chunk
wl = B.Workload(vb, cfg, dev, 0, chunk)                       # allocates one chunk of frames (and fills it with frames 0 .. chunk-1)
bufs = [wl.d_frames, vb.device_alloc(dev, chunk * wl.fstride * 4)]   # double buffer: chunk k + 1 is generated while chunk k is evaluated
plan = wl.plan(F, batch_frames=148)
gen_s = 0.0
torch.cuda.synchronize(dev); t0 = time.perf_counter()
for k in range(nchunks):
    g0 = time.perf_counter()
    vb.synth_water_frames_device(dev, B.WATER_N, B.WATER_SEED, wl.d_base, k * chunk, chunk, bufs[k & 1], wl.fstride, wl.na)   # synchronous: the generator's own stream is synchronised
    gen_s += time.perf_counter() - g0
    plan.eval_device_frames(bufs[k & 1], wl.fstride, wl.na, wl.cell, k * chunk, chunk)
    plan.sync()       # a sync per chunk keeps the example simple; cost: one pipeline drain per 18 944 frames
plan.sync(); torch.cuda.synchronize(dev); dt = time.perf_counter() - t0
r = plan.property_data("r"); v = plan.property_data("v")
print(json.dumps({"workload": B.workload_string(cfg), "frames": F, "n_gpus": 1, "wall_s": dt, "frames_per_s_incl_generation": F / dt, "generation_s": gen_s,
                  "frames_per_s_evaluation": F / (dt - gen_s), "frames_accumulated": int(r.frames_accumulated),
                  "checks": {"r_sum_per_frame": float(np.float64(r.values[:1024]).sum()), "v_sum_per_frame": float(np.float64(v.values).sum())}}))
