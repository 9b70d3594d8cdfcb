import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, viamd_b200 as vb
from helpers import load_golden, golden_system, vb_system, vb_cell
import torch
g = load_golden("water6.npz"); sysm = vb_system(golden_system(g)); F = g["frames"].shape[0]
cells = [vb_cell(g["cells"][f], g["cell_flags"][f]) for f in range(F)]
script = ("r = rdf(element('O'), element('O'), 6.0); v = sdf(residue(1:20), element('O'), 5.0); dz = density_z(element('O')); d = distance(1,10); "
          "rw = rdf(within(4.0, residue(1)), element('O'), 6.0); cc = contact_count(residue(1:5), residue(10:40), 4.0); aa = angle(residue(1:2), residue(5:7), 30); "
          "rt = rdf(element('O'), residue(10:60), 6.0); dm = distance_min(residue(1:4), residue(100:130)); pl = plane(residue(1:10)); rm = rmsd(residue(1:10));")
free0 = None
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    plan = vb.Plan(sysm, vb.compile_script(script, sysm), F, keep_frame_results=True)
    plan.set_initial_frame(*g["frames"][0], cells[0]); plan.eval_host_frames(g["frames"], cells, 0); plan.property_data("v"); plan.close()
    free, total = torch.cuda.mem_get_info(0)
    if it == 2: free0 = free
    if it % 10 == 0 or it < 3: print(it, 'free MB', free // 2**20)
print('leak since iteration 2 (MB):', (free0 - free) / 2**20)
