#!/usr/bin/env python
"""bench.py — frames/s of the md_script per-frame hot path (RDF + SDF evaluation) on N B200s, next to the reference CPU path.

Metric / config (BASELINE.json): "frames/sec RDF+SDF eval, 100k-atom synthetic traj". Workload = synthetic water box
n=32 (98 304 atoms, 32 768 O, L = 99.328 A, viamd_b200/csrc/synth.h), script
    r = rdf(element('O'), element('O'), 10.0);  v = sdf(residue(1:1000), element('O'), 10.0);
i.e. BASELINE configs[1] and configs[2] evaluated together on every frame, as VIAMD evaluates all properties of a script per frame.

A "step" = one pass of the hot path over one batch of `--frames-per-step` frames. Every step reads different frames (no frame is
reused inside the timed region), and one step's input (frames_per_step x 1.18 MB) is far larger than the 126 MB L2.

  value : whole-job frames/s with the frames already resident in HBM when the timed region starts (device stopwatch: CUDA events
          on the library's streams, max over ranks).
  e2e   : same metric through the public host API with HOST (pinned) buffers: every step copies its frames host->device and reads the
          step's results (RDF bins/weights 8 KB + SDF volume 8 MB) back to the host inside the timed region.
  N > 1 : frames are sharded contiguously per rank (weak scaling: every rank processes frames_per_step frames per step); one NCCL
          all-reduce of the RDF bins and SDF voxel grid at the end, inside the timed region.

`--impl reference` times the reference's own CPU md_script_eval_frame_range (oracle/_ref/ref_harness_fast: the unmodified mdlib
sources compiled with their shipped flags) on all host cores for the same script/workload, on a bounded sample of frames per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec RDF+SDF eval, 100k-atom synthetic traj"
UNIT = "frames/s"
SCRIPT = "r = rdf(element('O'), element('O'), 10.0); v = sdf(residue(1:1000), element('O'), 10.0);"
WATER_N, WATER_SEED = 32, 1234
WORKLOAD = (f"BASELINE configs[1]+[2] on one trajectory: synthetic water n={WATER_N} ({3 * WATER_N ** 3} atoms, {WATER_N ** 3} O, "
            f"L={WATER_N * 3.104:.3f} A), script: {SCRIPT}")


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). In-process NVML (pynvml) when available:
    spawning nvidia-smi several times a second on an 8-GPU box perturbs the driver and slowed the sampled rank; nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.stop_ev, self.th = gpu_index, [], threading.Event(), None
        self.nv = None; self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = self.idx
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try: phys = int(vis.split(",")[self.idx])
                except Exception: phys = self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys); self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample_nvml(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM); mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        try: r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception: r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        def bit(name_new, name_old):
            m = getattr(nv, name_new, None) or getattr(nv, name_old, 0)
            return "Active" if (r & m) else "Not Active"
        self.rows.append([str(self.idx), str(sm), str(mx), "0",
                          bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                          bit("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                          bit("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                          bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")])

    def _run(self):
        while not self.stop_ev.is_set():
            try:
                if self.nv is not None: self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.05 if self.nv is not None else 0.5)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True); self.th.start()

    def stop(self):
        self.stop_ev.set()
        if self.th: self.th.join(3)
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(frames_per_launch: float):
    """dram__bytes_read+write of the dominant kernel per launch, from the committed ncu --set full capture (scaled to this run's
    frames per launch); None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "rdf_traffic.json")
    try:
        t = json.load(open(p))
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) * (frames_per_launch / t["frames_per_launch"])
    except Exception:
        return None


def harness_path(kind="fast"):
    return os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{kind}")


def ensure_water_gro(tmpdir):
    tool = os.path.join(ROOT, "oracle", "build", "synth_tool")
    if not os.path.exists(tool):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    gro = os.path.join(tmpdir, f"water{WATER_N}_{WATER_SEED}.gro")
    if not os.path.exists(gro):
        subprocess.check_call([tool, "water-gro", str(WATER_N), str(WATER_SEED), gro])
    return gro


def run_reference_sample(frames: int, threads: int, repeat: int = 1, total_frames: int = 1 << 20):
    """Reference CPU md_script_eval_frame_range on `frames` frames with `threads` threads. Returns dict or None."""
    h = harness_path("fast")
    if not os.path.exists(h):
        return None
    tmp = os.environ.get("TMPDIR", "/tmp")
    gro = ensure_water_gro(tmp)
    out = subprocess.run([h, "time", "--sys", gro, "--traj", f"synthwater:{WATER_N}:{WATER_SEED}:{total_frames}", "--script", SCRIPT,
                          "--frames", f"0:{frames}", "--threads", str(threads), "--repeat", str(repeat)], capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    sys.stderr.write(out.stderr[-2000:])
    return None


def oracle_port_sample(frames: int):
    """Fallback CPU baseline: the plain-C oracle port on one core (kind 'port')."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as O
    import viamd_b200 as vb
    base, L = vb.synth_water_base(WATER_N, WATER_SEED)
    fr = vb.synth_water_frames_host(WATER_N, WATER_SEED, base, 0, frames)
    s = vb.water_system(WATER_N); o = np.arange(0, s.num_atoms, 3, dtype=np.int32)
    structs = np.arange(3000, dtype=np.int32).reshape(1000, 3); oc = O.UnitCell.ortho(L, L, L)
    vol = np.zeros(128 ** 3, np.float32)
    t0 = time.perf_counter()
    for f in range(frames):
        O.rdf_frame(*fr[f], o, o, oc, 0.0, 10.0)
        O.sdf_frame(*fr[f], fr[0], s.mass, structs, o, s.conn_offset, s.conn_idx, oc, 10.0, vol=vol)
    dt = time.perf_counter() - t0
    return {"frames": frames, "threads": 1, "best_s": dt, "frames_per_s": frames / dt}


def impl_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    threads = cores
    # bounded sample per step: ~2 frames per thread, so K+W steps finish within minutes even on few cores
    sample = max(2 * threads, 16)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": sample, "threads": threads}}
    if os.path.exists(harness_path("fast")):
        for _ in range(args.warmup):
            run_reference_sample(sample, threads)
        t_total, frames_total = 0.0, 0
        for _ in range(args.steps):
            r = run_reference_sample(sample, threads)
            t_total += r["best_s"]; frames_total += r["frames"]
        kind = "reference"
    else:
        t_total, frames_total = 0.0, 0
        for _ in range(max(1, min(args.steps, 3))):
            r = oracle_port_sample(4); t_total += r["best_s"]; frames_total += r["frames"]
        kind, threads, sample = "port", 1, 4
    v = frames_total / t_total
    line.update({"value": v, "ms_per_step": 1e3 * t_total / max(1, args.steps),
                 "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                  "sample": f"{sample} frames per step x {args.steps} steps of the same workload, in-memory trajectory"},
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)
    return 0


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the process's real stdout; everything else written to fd 1 meanwhile (NCCL's version banner, library
    chatter) was redirected to stderr in main()"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None: sys.stdout.write(data.decode()); sys.stdout.flush()
    else: os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush(); _REAL_STDOUT = os.dup(1); os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=4736)   # 32 x 148 frames
    ap.add_argument("--batch-frames", type=int, default=0)
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--rdf-variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return impl_reference(args)

    import numpy as np
    import torch
    import viamd_b200 as vb
    from viamd_b200 import dist as vdist

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as tdist
        torch.cuda.set_device(local_rank)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    if vb.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")

    K, W, FPS = args.steps, max(args.warmup, 3), args.frames_per_step
    n, seed = WATER_N, WATER_SEED
    base, L = vb.synth_water_base(n, seed); na = base.shape[1]
    sysm = vb.water_system(n)
    props = vb.compile_script(SCRIPT, sysm)
    total_steps = W + K
    # all (warm-up + timed) steps read distinct device-resident frames; keep that shard under ~60 GB of the 180 GB HBM
    max_fps = int(60e9 // (total_steps * 3 * na * 4)) // 148 * 148
    FPS = max(148, min(FPS, max_fps))
    frames_local = total_steps * FPS
    plan = vb.Plan(sysm, props, frames_local, device=dev, batch_frames=args.batch_frames, num_streams=args.streams, rdf_variant=args.rdf_variant)
    cell = vb.UnitCell.from_basis(L, L, L)
    f0 = vb.synth_water_frames_host(n, seed, base, 0, 1)
    plan.set_initial_frame(*f0[0], cell)   # frame 0 of the trajectory is the initial configuration on every rank

    # ---- device-resident frames of this rank's shard: global frame index = rank * frames_local + i
    fstride = 3 * na
    d_base = vb.device_alloc(dev, base.nbytes); vb.memcpy_h2d(dev, d_base, base.ctypes.data, base.nbytes)
    d_frames = vb.device_alloc(dev, frames_local * fstride * 4)
    vb.synth_water_frames_device(dev, n, seed, d_base, rank * frames_local, frames_local, d_frames, fstride, na)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(i):
        plan.eval_device_frames(d_frames + i * FPS * fstride * 4, fstride, na, cell, i * FPS, FPS)

    # ---- warm-up (untimed)
    for i in range(W):
        step_device(i)
    plan.sync()
    plan.enable_kernel_timing(True)
    vb.launch_count(reset=True)
    sampler = ClockSampler(dev); sampler.start()
    barrier()
    plan.timer_begin()
    t_wall0 = time.perf_counter()
    for i in range(W, W + K):
        step_device(i)
    if world > 1:
        plan.sync()
        vdist.allreduce_plan(plan, total_frames=world * (W + K) * FPS)   # the one exchange step: bins + voxels over NCCL
    ms_dev = plan.timer_end()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = vb.launch_count()
    plan.enable_kernel_timing(False)
    k_ms, k_n = plan.kernel_time_ms("k_rdf_pairs")
    if world > 1:
        t = torch.tensor([ms_dev], dtype=torch.float64, device=f"cuda:{dev}")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX); ms_dev = float(t.item())
        t = torch.tensor([float(launches)], dtype=torch.float64, device=f"cuda:{dev}")
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM); launches = int(t.item())
    value = world * K * FPS / (ms_dev * 1e-3)
    d_r = plan.property_data("r"); d_v = plan.property_data("v")
    checks = {"rdf_pairs_per_frame": float(np.float64(d_r.values[:1024]).sum()), "sdf_counts_per_frame": float(np.float64(d_v.values).sum())}

    # ---- end-to-end through the host API: pinned host frames -> H2D -> kernels -> D2H of the step's results
    e2e = None
    if not args.no_e2e:
        plan.clear(); plan.set_initial_frame(*f0[0], cell)
        nbuf = 2
        h_ptrs = [vb.host_alloc_pinned(FPS * fstride * 4) for _ in range(nbuf)]
        for b, hp in enumerate(h_ptrs):   # fill the pinned staging buffers from the device-generated frames (exact same data)
            vb.memcpy_d2h(dev, hp, d_frames + b * FPS * fstride * 4, FPS * fstride * 4)
        Ke = max(2, min(K, 4))

        def step_host(i):
            plan.eval_host_ptr(h_ptrs[i % nbuf], fstride, na, cell, i * FPS, FPS)
            dr = plan.property_data("r"); dv = plan.property_data("v")   # sync + D2H of bins and volume
            return dr.values[0] + dv.values[0]

        for i in range(2):
            step_host(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(2, 2 + Ke):
            step_host(i)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX); dt = float(t.item())
        e2e = {"value": world * Ke * FPS / dt, "unit": UNIT, "h2d_bytes_per_step": FPS * fstride * 4,
               "d2h_bytes_per_step": 2 * 1024 * 8 + 128 ** 3 * 4 + 2 * frames_local * 16, "steps": Ke}
        for hp in h_ptrs:
            vb.host_free_pinned(hp)

    if rank == 0:
        peak, peak_src = measured_peaks()
        B = args.batch_frames or 148
        algo_bytes_per_frame = 12 * 32768    # xyz of the selected (O) atoms read once; bins stay on chip
        avg_launch_s = (k_ms / max(k_n, 1)) * 1e-3
        frames_per_launch = (K * FPS) / max(k_n, 1)
        achieved = algo_bytes_per_frame * frames_per_launch / max(avg_launch_s, 1e-12) / 1e9
        pair_tests_per_frame = 32768 * 27 * (32768 / 729.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_step": FPS, "batch_frames": B, "parallelism": f"frame-sharded x{world}",
                       "l2_policy": "every step reads fresh frames; one step's input (%.2f GB) exceeds the 126 MB L2" % (FPS * fstride * 4 / 1e9)},
            "gpu_launches": launches,
            "clocks": clocks,
            "wall_s": t_wall,
            "checks": checks,
            "roofline": {"bound": "hbm", "kernel": "k_rdf_pairs", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(frames_per_launch), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes_per_frame * frames_per_launch,
                         "note": "k_rdf_pairs is FP32-ALU/shared-atomic bound, not DRAM bound (SURVEY.md §7): see alu",
                         "kernel_share_of_step": (k_ms / max(ms_dev, 1e-9)),
                         "alu": {"pair_tests_definition": "the reference's enumeration (refs x 27 neighbour cells x mean cell population); tests avoided by the symmetric mode and the exact cull count as done",
                                 "pair_tests_per_s": pair_tests_per_frame * frames_per_launch / max(avg_launch_s, 1e-12),
                                 "avg_launch_ms": avg_launch_s * 1e3, "launches_timed": k_n}},
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = max(2 * cores, 32)
            r = run_reference_sample(sample, cores)
            if r:
                line["cpu_baseline"] = {"value": r["frames_per_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{sample} frames of the same workload through md_script_eval_frame_range (oracle/_ref/ref_harness_fast, shipped flags), {cores} threads, in-memory trajectory"}
            else:
                r = oracle_port_sample(4)
                line["cpu_baseline"] = {"value": r["frames_per_s"], "unit": UNIT, "cores": 1, "kind": "port", "sample": "4 frames, oracle/md_oracle.c scalar port"}
        emit(line)
    vb.device_free(dev, d_base); vb.device_free(dev, d_frames)
    plan.close()
    if world > 1:
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
