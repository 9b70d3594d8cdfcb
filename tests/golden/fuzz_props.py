"""Fuzz of the oracle's property evaluators against the UNMODIFIED reference (needs /root/reference: oracle/_ref/ref_harness_strict).
Water boxes of random size whose cell is rescaled anisotropically per frame (NPT-like), random cutoffs and selections; every per-frame
rdf bin / weight, sdf voxel, density bin, temporal value has to agree bit for bit. Run here:  python tests/golden/fuzz_props.py [cases] [seed]"""
import os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refio, oracle_lib as O
from make_golden import HARNESS, SYNTH, run
from helpers import cell_from_row, dense_from_sparse


def main(cases=40, seed=3):
    rng = np.random.default_rng(seed); bad = 0; checks = 0
    with tempfile.TemporaryDirectory() as tmp:
        for c in range(cases):
            n = int(rng.choice([4, 5, 6, 7])); sd = int(rng.integers(1, 10000)); F = 3
            gro, raw0, raw, out, si = [os.path.join(tmp, f"c{c}.{e}") for e in ("gro", "raw0", "raw", "out", "sys")]
            run(SYNTH, "water-gro", str(n), str(sd), gro); run(SYNTH, "water-raw", str(n), str(sd), str(F), raw0)
            fr, cells, flags = refio.read_raw_traj(raw0)
            L = cells[0][0]
            # anisotropic per-frame rescale (coordinates follow the cell), optionally one non-periodic axis
            sc = 1.0 + 0.08 * (rng.random((F, 3)) - 0.5); sc[0] = 1.0 + 0.05 * (rng.random(3) - 0.5)
            fr2 = (fr.astype(np.float64) * sc[:, :, None]).astype(np.float32)
            cells2 = np.zeros((F, 6)); cells2[:, 0] = L * sc[:, 0]; cells2[:, 3] = L * sc[:, 1]; cells2[:, 5] = L * sc[:, 2]
            fl = 1 | 4 | 8 | 16
            tri = rng.random() < 0.25
            if tri:   # shear into a triclinic cell that changes every frame (a = (x,0,0), b = (xy,y,0), c = (xz,yz,z))
                fl = 2 | 4 | 8 | 16
                for f in range(F):
                    xy, xz, yz = rng.uniform(-0.25, 0.25, 3) * L
                    X, Y, Z = fr2[f].astype(np.float64); Lx, Ly, Lz = cells2[f][0], cells2[f][3], cells2[f][5]
                    fr2[f, 0] = (X + (xy / Ly) * Y + (xz / Lz) * Z).astype(np.float32); fr2[f, 1] = (Y + (yz / Lz) * Z).astype(np.float32)
                    cells2[f][1] = xy; cells2[f][2] = xz; cells2[f][4] = yz
            elif rng.random() < 0.2: fl &= ~int(rng.choice([4, 8, 16]))
            refio.write_raw_traj(raw, fr2, cells2, np.full(F, fl, np.uint32))
            cut = float(np.round(rng.uniform(2.5, 0.62 * L), 2)); cmin = float(np.round(rng.uniform(0.5, 2.0), 2))
            nres = int(rng.integers(2, 12)); a1 = int(rng.integers(1, 3 * n ** 3 - 40)); a2 = a1 + int(rng.integers(3, 30))
            sdf_ok = (fl & 28) == 28
            script = (f"r = rdf(element('O'), element('O'), {cut}); rh = rdf(element('H'), element('O'), {cmin}:{cut}); rc = rdf(residue(1:{nres}), element('H'), {cut}); "
                      f"dz = density_z(element('O')); dy = density_y(element('H')); d = distance({a1},{a2}); dg = distance(atom({a1}:{a2}), residue(2)); "
                      f"an = angle({a1},{a1 + 1},{a2}); dmn = distance_min(atom({a1}:{a2}), residue(1)); " + f"rm = rmsd(residue(1:{nres})); " + f"cw = count(within({min(cut, 6.0)}, residue(1))); "
                      f"dp = distance_pair(atom({a1}:{a1 + 4}), residue(1)); cm = com(atom({a1}:{a2})); pl = plane(atom({a1}:{a2})); cwr = count(within({cmin}:{min(cut, 6.0)}, residue(1))); "
                      f"cwo = count(element('O') and within({min(cut, 6.0)}, residue(2))); anc = angle(2,1,3) in residue(1:{nres}); dpg = distance_pair(residue(1:{nres}), residue({nres + 1}:{nres + 3}));"
                      + (f" v = sdf(residue(1:{nres}), element('O'), {min(cut, 0.45 * L):.2f});" if sdf_ok else ""))
            p = subprocess.run([HARNESS, "eval", "--sys", gro, "--traj", f"raw:{raw}", "--script", script, "--out", out, "--perframe", f"0:{F}", "--full", f"0:{F}"], capture_output=True, text=True)
            if p.returncode != 0: print("reference failed on case", c, script); bad += 1; continue
            run(HARNESS, "sysinfo", "--sys", gro, "--out", si)
            s = refio.read_sysinfo(si); R = refio.read_refout(out)
            z = s["z"].astype(int); o = np.nonzero(z == 8)[0].astype(np.int32); h = np.nonzero(z == 1)[0].astype(np.int32)
            co = s["comp_off"]; groups = [np.arange(co[r], co[r + 1], dtype=np.int32) for r in range(nres)]
            A = lambda a, b: np.arange(a - 1, b, dtype=np.int32)
            c0 = cell_from_row(cells2[0], fl)
            def chk(ok, what, f):
                nonlocal bad, checks
                checks += 1
                if not ok: bad += 1; print("MISMATCH", what, "case", c, "frame", f, dict(n=n, cut=cut, cmin=cmin, flags=fl, nres=nres))
            for f in range(F):
                x, y, zz = fr2[f]; cell = cell_from_row(cells2[f], fl)
                b, w, _ = O.rdf_frame(x, y, zz, o, o, cell, 0.0, cut); chk(np.array_equal(b, R["r"].perframe[f][:1024]) and np.array_equal(w, R["r"].perframe[f][1024:]), "r", f)
                b, w, _ = O.rdf_frame(x, y, zz, h, o, cell, cmin, cut); chk(np.array_equal(b, R["rh"].perframe[f][:1024]) and np.array_equal(w, R["rh"].perframe[f][1024:]), "rh", f)
                pos, off, idx = O.group_com(x, y, zz, s["mass"], groups)
                b, w, _ = O.rdf_frame(x, y, zz, None, h, cell, 0.0, cut, ref_pos=pos, excl_off=off, excl_idx=idx); chk(np.array_equal(b, R["rc"].perframe[f][:1024]) and np.array_equal(w, R["rc"].perframe[f][1024:]), "rc", f)
                b, _ = O.density_frame(x, y, zz, s["mass"], o, c0, 2); chk(np.array_equal(b, R["dz"].perframe[f][:1024]), "dz", f)
                b, _ = O.density_frame(x, y, zz, s["mass"], h, c0, 1); chk(np.array_equal(b, R["dy"].perframe[f][:1024]), "dy", f)
                chk(np.float32(O.distance(x, y, zz, a1 - 1, a2 - 1, cell)) == R["d"].full[f], "d", f)
                chk(O.distance_args(x, y, zz, s["mass"], A(a1, a2), groups[1], cell) == R["dg"].full[f], "dg", f)
                chk(np.float32(O.angle(x, y, zz, a1 - 1, a1, a2 - 1)) == R["an"].full[f], "an", f)
                chk(O.min_distance(x, y, zz, A(a1, a2), groups[0], cell) == R["dmn"].full[f], "dmn", f)
                chk(O.rmsd_frame(x, y, zz, fr2[0], s["mass"], np.concatenate(groups), s["conn_off"], s["conn_idx"], cell) == R["rm"].full[f], "rm", f)
                chk(len(O.within(x, y, zz, groups[0], min(cut, 6.0), cell)) == int(R["cw"].full[f]), "cw", f)
                chk(np.array_equal(O.distance_pair(x, y, zz, A(a1, a1 + 4), groups[0], cell), R["dp"].full[15 * f:15 * f + 15]), "dp", f)
                chk(np.array_equal(O.arg_position(x, y, zz, s["mass"], A(a1, a2), cell), R["cm"].full[3 * f:3 * f + 3]), "cm", f)
                chk(np.array_equal(O.plane_frame(x, y, zz, A(a1, a2), s["conn_off"], s["conn_idx"], cell), R["pl"].full[4 * f:4 * f + 4]), "pl", f)
                chk(len(O.within(x, y, zz, groups[0], min(cut, 6.0), cell, rmin=cmin)) == int(R["cwr"].full[f]), "cwr", f)
                chk(len(np.intersect1d(O.within(x, y, zz, np.arange(co[1], co[2], dtype=np.int32), min(cut, 6.0), cell), o)) == int(R["cwo"].full[f]), "cwo", f)
                chk(np.array_equal(np.array([O.angle(x, y, zz, co[r] + 1, co[r], co[r] + 2) for r in range(nres)], np.float32), R["anc"].full[nres * f:nres * f + nres]), "anc", f)
                g2 = [np.arange(co[r], co[r + 1], dtype=np.int32) for r in range(nres, nres + 3)]
                chk(np.array_equal(O.distance_pair_args(x, y, zz, s["mass"], groups, g2, cell), R["dpg"].full[3 * nres * f:3 * nres * (f + 1)]), "dpg", f)
                if sdf_ok:
                    vol, nn = O.sdf_frame(x, y, zz, fr2[0], s["mass"], np.stack(groups), o, s["conn_off"], s["conn_idx"], cell, float(f"{min(cut, 0.45 * L):.2f}"))
                    chk(np.array_equal(vol, R["v"].perframe[f]), "v", f)
            for pth in (gro, raw0, raw, out, si): os.remove(pth)
    print(f"{checks} checks in {cases} cases, {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(*(int(a) for a in sys.argv[1:3])) else 0)
