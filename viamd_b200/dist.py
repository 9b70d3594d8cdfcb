"""Frame sharding across GPUs + the final exchange step.

Frames are independent units: rank g evaluates the contiguous block [F*g/G, F*(g+1)/G) into its own integer accumulators
(no data-path collective), then ONE exchange at the end: all-reduce(sum) of the RDF bins (u64[1024]) and the SDF voxel grid
(u32[128^3]) over NCCL/NVLink, all-reduce(min/max) of the per-frame bin extrema, and the pair total of the globally last frame
(for the RDF weights). Dividing by the global frame count happens after the reduce, so the result is independent of G.
torch.distributed is plumbing only; on CPU (gloo) the same code path runs on host copies of the accumulators.
"""
from __future__ import annotations

import numpy as np

from . import api


def frame_shard(num_frames: int, world: int, rank: int):
    """Contiguous frame block of `rank` (the partition VIAMD's range task uses per thread, src/task_system.cpp:73-87)."""
    return (num_frames * rank) // world, (num_frames * (rank + 1)) // world


class _CudaView:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def allreduce_counts_(counts, group=None):
    """In-place sum of an integer tensor across ranks (works for CPU tensors with gloo and CUDA tensors with nccl)."""
    import torch.distributed as dist
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def rdf_weights(total_pairs: int, cutoff_min: float, cutoff_max: float) -> np.ndarray:
    """compute_rdf's normalisation (md_script_functions.inl:5323-5337), same double arithmetic as csrc/plan.cu."""
    mn = np.float32(max(np.float32(cutoff_min), np.float32(1e-3))); mx = np.float32(cutoff_max)
    sv = lambda r: (4.0 / 3.0) * 3.1415926535897932 * (r * r * r)
    total_vol = sv(float(mx)) - sv(float(mn))
    rho = float(total_pairs) / total_vol
    dr = float(np.float32(np.float32(mx - mn) / np.float32(1024)))
    w = np.empty(1024, np.float32); prev = 0.0
    for i in range(1024):
        s = sv(float(mn) + (i + 0.5) * dr); w[i] = np.float32(rho * (s - prev)); prev = s
    return w


def allreduce_plan(plan: api.Plan, total_frames: int, group=None, device=None, view=None):
    """The single exchange step for a frame-sharded evaluation. Reduces every accumulator of `plan` (integer bins / voxels / mass sums; float rows
    of temporals, which are disjoint between ranks)
    in place on the device (NCCL) and tells the plan the global frame count. Returns {name: merged extras}.
    `view(ptr, count, typestr)` turns an accumulator into the tensor handed to all_reduce; the default wraps device memory."""
    import torch
    import torch.distributed as dist
    plan.sync()
    extras = {}
    dev = torch.device("cuda", plan.device if device is None else device) if view is None else None
    if view is None:
        view = lambda ptr, n, typestr: torch.as_tensor(_CudaView(ptr, n, typestr), device=dev)
    integer_ops = (api.OP_RDF, api.OP_SDF, api.OP_DENSITY_X, api.OP_DENSITY_Y, api.OP_DENSITY_Z)
    have_temporal = any(p.op not in integer_ops for p in plan.properties)
    if have_temporal and plan.num_frames != int(total_frames):
        raise ValueError("temporal properties are exchanged by global frame index: create every rank's plan with the global frame count "
                         "and evaluate its shard at its global offsets (frame_shard)")
    for p in plan.properties:
        ptr, nbytes, eb = plan.accum_ptr(p.name)
        if p.op in integer_ops:   # bins / voxels / fixed-point sums: exact integer sums, then the mean over the global frame count
            t = view(ptr, nbytes // eb, "<i8" if eb == 8 else "<i4")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            plan.set_frames_accumulated(p.name, int(total_frames))
            if plan.num_frames == int(total_frames):   # plans over the global frame range: the per-frame rows (totals -> rdf weights, frame min / max) are disjoint too
                for which in (0, 1, 2):
                    rptr, rbytes, reb = plan.frame_rows(p.name, which)
                    if rptr: dist.all_reduce(view(rptr, rbytes // reb, "<i8" if reb == 8 else "<i4"), op=dist.ReduceOp.SUM, group=group)
        else:                     # temporals: disjoint rows, zero elsewhere -> x + 0 merges them exactly
            t = view(ptr, nbytes // 4, "<f4")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    # plans over the global frame range: every rank now holds every frame's rows, so min / max, ranges, aggregates and the rdf weights (pair
    # total of the globally last frame) must cover all of them — also when the plan has no temporal property
    if plan.num_frames == int(total_frames): plan.mark_frames_done(0, int(total_frames))
    if dev is not None: torch.cuda.synchronize(dev)
    return extras
