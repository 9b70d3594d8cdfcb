"""Fuzz of the oracle's XTC decoder against the UNMODIFIED reference (needs /root/reference: oracle/_ref/ref_harness_strict).
Random systems (atom count, box, clustering, precision) are written by the reference's xdrfile writer, decoded by the reference's md_xtc
reader and by oracle/md_oracle.c; every coordinate has to agree bit for bit. Run here:  python tests/golden/fuzz_xtc.py [cases] [seed]"""
import os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refio, oracle_lib as O
from make_golden import HARNESS


def _write_gro(path, n, L):   # atoms on a 0.5 nm grid: no accidental bonds for the reference's topology post-processing
    with open(path, "w") as f:
        f.write("synthetic\n%d\n" % n)
        for i in range(n):
            f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (i % 99999 + 1, "ARG", "AR", i % 99999 + 1, 0.5 * (i % 20), 0.5 * ((i // 20) % 20), 0.5 * (i // 400)))
        f.write("%10.5f%10.5f%10.5f\n" % (L, L, L))

def main(cases=120, seed=7):
    rng = np.random.default_rng(seed); bad = 0; total_frames = 0; ref_crash = 0; ref_wrong = 0
    with tempfile.TemporaryDirectory() as tmp:
        for c in range(cases):
            n = int(rng.choice([10, 11, 12, 50, 333, 1000, 3000])); F = 3
            span = float(rng.choice([2.0, 20.0, 80.0, 400.0, 3000.0])); prec = float(rng.choice([10, 100, 1000, 10000]))
            if span / 10.0 * prec >= 2 ** 21: prec = 100.0   # stay inside the <= 64-bit packed branch: beyond it the reference reader is wrong (and can abort)
            mode = rng.integers(0, 3)
            if mode == 0: fr = rng.random((F, 3, n)) * span                                              # uniform
            elif mode == 1:                                                                                 # molecules: clusters of 3 within 1 A
                cen = rng.random((F, 3, (n + 2) // 3)) * span; fr = np.repeat(cen, 3, axis=2)[:, :, :n] + rng.normal(0, 0.5, (F, 3, n))
            else: fr = np.cumsum(rng.normal(0, 0.3, (F, 3, n)), axis=2) + span / 2                         # chain: small steps between consecutive atoms
            fr = fr.astype(np.float32)
            gro, raw, xtc, dec = [os.path.join(tmp, f"c{c}.{e}") for e in ("gro", "raw", "xtc", "dec")]
            _write_gro(gro, n, max(span / 10 + 1, 12.0))
            refio.write_raw_traj(raw, fr, np.tile([span + 1, 0, 0, span + 1, 0, span + 1], (F, 1)), np.full(F, 29, np.uint32))
            subprocess.check_call([HARNESS, "xtcwrite", "--sys", gro, "--traj", f"raw:{raw}", "--out", xtc, "--precision", str(prec)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            rc = subprocess.call([HARNESS, "dumptraj", "--sys", gro, "--traj", f"xtc:{xtc}", "--out", dec], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            blob = np.fromfile(xtc, np.uint8); offs = O.xtc_frame_offsets(blob)
            assert len(offs) == F + 1
            if rc != 0:   # the reference reader aborts when an axis spans a single integer (libdivide "divider must be != 1", md_xtc.c:826): check the round trip instead
                ref_crash += 1
                for f in range(F):
                    ok, xyz, cell, step, tm = O.xtc_decode_frame(blob, offs[f], offs[f + 1], n); total_frames += 1
                    if not ok or np.abs(xyz.astype(np.float64) - fr[f]).max() > 5.0 / prec + 1e-3 * span / 100: bad += 1; print("ROUNDTRIP MISMATCH case", c, f)
            else:
                ref, cells, flags = refio.read_raw_traj(dec)
                for f in range(F):
                    ok, xyz, cell, step, tm = O.xtc_decode_frame(blob, offs[f], offs[f + 1], n)
                    total_frames += 1
                    if ok and np.array_equal(xyz, ref[f]) and [cell.x, cell.xy, cell.xz, cell.y, cell.yz, cell.z] == list(cells[f]): continue
                    tol = 5.0 / prec + 1e-3 * span / 100
                    if ok and np.abs(xyz.astype(np.float64) - fr[f]).max() <= tol and np.abs(ref[f].astype(np.float64) - fr[f]).max() > tol:
                        ref_wrong += 1   # the oracle reproduces what was written, the reference reader does not (small-integer fields wider than 57 bits: md_xtc.c:850 reads them with extract_bits_be_raw_57)
                    else:
                        bad += 1; print("MISMATCH case", c, "frame", f, dict(n=n, span=span, prec=prec, mode=int(mode)))
            for p in (gro, raw, xtc, dec):
                if os.path.exists(p): os.remove(p)
    print(f"{total_frames} frames in {cases} cases, {bad} mismatches; the reference reader aborted on {ref_crash} cases (single-integer axis) "
          f"and mis-decoded {ref_wrong} frames that the oracle round-trips (small fields wider than 57 bits)")
    return bad

if __name__ == "__main__":
    sys.exit(1 if main(*(int(a) for a in sys.argv[1:3])) else 0)
