"""Fuzz of the device path against the oracle (no reference needed; run on a B200 box):  python tests/golden/fuzz_gpu.py [cases] [seed]
Without a GPU:  python tests/golden/fuzz_gpu.py [cases] [seed] --emulated   (the library's sources executed on the CPU, ~1 min per case)
Random water boxes with anisotropic per-frame cells (one axis non-periodic or a sheared triclinic cell in part of the cases), random cutoffs
and selections; per-frame rdf bins (plain, min:max, centre-of-mass references), sdf voxels, density sums, temporals have to equal the oracle's
(bit-exact for integers and distances, 1e-5 for angles). Not collected by pytest: it is a search tool, the fixed cases live in tests/."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import viamd_b200 as vb

if "--emulated" in sys.argv:   # no GPU: the library's own sources compiled for the CPU (tests/emul) — test infrastructure, swapped in inside this process only
    sys.argv.remove("--emulated"); sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_emul, viamd_b200.api as _api
    _api.LIB_PATH = build_emul.build_library(); _api._lib = None


def main(cases=40, seed=1):
    rng = np.random.default_rng(seed); bad = 0; checks = 0
    for c in range(cases):
        n = int(rng.choice([4, 5, 6, 8])); sd = int(rng.integers(1, 10000)); F = 4
        base, L = vb.synth_water_base(n, sd); fr = vb.synth_water_frames_host(n, sd, base, 0, F).astype(np.float64)
        sysm = vb.water_system(n); na = 3 * n ** 3
        sc = 1.0 + 0.08 * (rng.random((F, 3)) - 0.5)
        fr *= sc[:, :, None]
        flags = vb.CELL_ORTHO | vb.CELL_PBC_ALL; tri = rng.random() < 0.25
        cellp = np.zeros((F, 6)); cellp[:, 0] = L * sc[:, 0]; cellp[:, 3] = L * sc[:, 1]; cellp[:, 5] = L * sc[:, 2]
        if tri:
            flags = vb.CELL_TRICLINIC | vb.CELL_PBC_ALL
            for f in range(F):
                xy, xz, yz = rng.uniform(-0.25, 0.25, 3) * L
                X, Y, Z = fr[f].copy(); fr[f, 0] = X + (xy / cellp[f, 3]) * Y + (xz / cellp[f, 5]) * Z; fr[f, 1] = Y + (yz / cellp[f, 5]) * Z
                cellp[f, 1], cellp[f, 2], cellp[f, 4] = xy, xz, yz
        elif rng.random() < 0.25: flags &= ~int(rng.choice([vb.CELL_PBC_X, vb.CELL_PBC_Y, vb.CELL_PBC_Z]))
        fr = fr.astype(np.float32)
        cut = float(np.round(rng.uniform(2.5, 0.62 * L), 2)); cmin = float(np.round(rng.uniform(0.5, 2.0), 2)); nres = int(rng.integers(2, 12))
        a1 = int(rng.integers(0, na - 40)); a2 = a1 + int(rng.integers(3, 30))
        o = np.arange(0, na, 3, dtype=np.int32); h = np.setdiff1d(np.arange(na, dtype=np.int32), o)
        groups = [np.arange(3 * r, 3 * r + 3, dtype=np.int32) for r in range(nres)]
        full_pbc = (flags & vb.CELL_PBC_ALL) == vb.CELL_PBC_ALL
        scut = float(np.round(min(cut, 0.45 * L), 2))
        props = [vb.rdf("r", o, o, cut), vb.rdf("rh", h, o, cut, cmin), vb.rdf_com("rc", groups, h, cut), vb.density("dz", 2, o),
                 vb.distance("d", a1, a2), vb.distance("dg", np.arange(a1, a2 + 1), groups[1]), vb.angle("an", a1, a1 + 1, a2), vb.distance_min("dmn", np.arange(a1, a2 + 1), groups[0])]
        sel = np.arange(a1, a2 + 1, dtype=np.int32); wr = float(np.round(min(cut, rng.uniform(2.0, 7.0)), 2)); allg = np.concatenate(groups)
        props += [vb.rmsd("rm", allg), vb.distance_pair("dp", sel[:5], groups[0]), vb.com("cm", sel), vb.plane("pl", sel), vb.count_within("cw", wr, groups[0]),
                  vb.rdf_within("rw", wr, groups[0], o, cut, cmin)]
        if full_pbc: props.append(vb.sdf("v", np.stack(groups), o, scut))
        cells = [vb.UnitCell(*cellp[f], flags) for f in range(F)]; ocells = [O.UnitCell.from_params(*cellp[f], flags) for f in range(F)]
        print("case", c, dict(n=n, seed=sd, cut=cut, cmin=cmin, flags=flags, nres=nres, wr=wr, tri=bool(tri)), flush=True)
        plan = vb.Plan(sysm, props, F, keep_frame_results=True, batch_frames=int(rng.choice([1, 3, 4])))
        plan.set_initial_frame(*fr[0], cells[0])
        def chk(ok, what, f):
            nonlocal bad, checks
            checks += 1
            if not ok: bad += 1; print("MISMATCH", what, "case", c, "frame", f, dict(n=n, seed=sd, cut=cut, cmin=cmin, flags=flags, nres=nres))
        for f in range(F):
            plan.clear(); plan.eval_host_frames(fr[f:f + 1], [cells[f]], f)
            x, y, z = fr[f]; oc = ocells[f]
            for key, ref, trg, lo in (("r", o, o, 0.0), ("rh", h, o, cmin)):
                b, w, t = O.rdf_frame(x, y, z, ref, trg, oc, lo, cut); gb, gt = plan.frame_counts(key, f)
                chk(gt == t and np.array_equal(gb.astype(np.float32), b), key, f)
            pos, off, idx = O.group_com(x, y, z, sysm.mass, groups)
            b, w, t = O.rdf_frame(x, y, z, None, h, oc, 0.0, cut, ref_pos=pos, excl_off=off, excl_idx=idx); gb, gt = plan.frame_counts("rc", f)
            chk(gt == t and np.array_equal(gb.astype(np.float32), b), "rc", f)
            chk(plan.property_data("d").values[f] == np.float32(O.distance(x, y, z, a1, a2, oc)), "d", f)
            chk(plan.property_data("dg").values[f] == O.distance_args(x, y, z, sysm.mass, np.arange(a1, a2 + 1), groups[1], oc), "dg", f)
            chk(abs(plan.property_data("an").values[f] - O.angle(x, y, z, a1, a1 + 1, a2)) <= 1e-5 * 3.2, "an", f)
            chk(plan.property_data("dmn").values[f] == O.min_distance(x, y, z, np.arange(a1, a2 + 1), groups[0], oc), "dmn", f)
            co, ci = np.asarray(sysm.conn_offset, np.uint32), np.asarray(sysm.conn_idx, np.int32)
            chk(plan.property_data("rm").values[f] == O.rmsd_frame(x, y, z, fr[0], sysm.mass, allg, co, ci, oc), "rm", f)
            npp = 3 * len(sel[:5])   # sel holds 4..30 atoms
            chk(np.array_equal(plan.property_data("dp").values[npp * f:npp * f + npp], O.distance_pair(x, y, z, sel[:5], groups[0], oc)), "dp", f)
            chk(np.array_equal(plan.property_data("cm").values[3 * f:3 * f + 3], O.arg_position(x, y, z, sysm.mass, sel, oc)), "cm", f)
            chk(np.array_equal(plan.property_data("pl").values[4 * f:4 * f + 4], O.plane_frame(x, y, z, sel, co, ci, oc)), "pl", f)
            wref = O.within(x, y, z, groups[0], wr, oc)
            chk(plan.property_data("cw").values[f] == len(wref), "cw", f)
            if len(wref):
                b, w, t = O.rdf_frame(x, y, z, wref, o, oc, cmin, cut); gb, gt = plan.frame_counts("rw", f)
                chk(gt == t and np.array_equal(gb.astype(np.float32), b), "rw", f)
            if full_pbc:
                vol, nn = O.sdf_frame(x, y, z, fr[0], sysm.mass, np.stack(groups), o, sysm.conn_offset, sysm.conn_idx, oc, scut)
                chk(np.array_equal(plan.counts("v").astype(np.float32), vol), "v", f)
        plan.close()
    print(f"{checks} checks in {cases} cases, {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(*(int(a) for a in sys.argv[1:3])) else 0)
