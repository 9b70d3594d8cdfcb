"""Writes oracle/_ref/xtc_water32_16.xtc: 16 frames of the bench workload (synthetic water n=32, seed 1234) through the reference's xdrfile
writer (needs /root/reference: oracle/_ref/ref_harness_strict). The file is git-ignored but travels to the GPU box; profiles/xtc_e2e.py reads it."""
import os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle", "_ref", "ref_harness_strict"); T = os.path.join(ROOT, "oracle", "build", "synth_tool")
with tempfile.TemporaryDirectory() as tmp:
    gro = os.path.join(tmp, "w.gro")
    subprocess.check_call([T, "water-gro", "32", "1234", gro])
    subprocess.check_call([H, "xtcwrite", "--sys", gro, "--traj", "synthwater:32:1234:16", "--out", os.path.join(ROOT, "oracle", "_ref", "xtc_water32_16.xtc")], stdout=subprocess.DEVNULL)
print(os.path.getsize(os.path.join(ROOT, "oracle", "_ref", "xtc_water32_16.xtc")))
