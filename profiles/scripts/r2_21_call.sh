set -x
mkdir -p gpurun_out
oracle/build/synth_tool water-gro 16 2024 /tmp/w16.gro
S="r = rdf(element('O'), element('O'), 8.0); d = distance(1,10); dz = density_z(element('O')); dp = distance_pair(residue(1:4), residue(10:15)); v = sdf(residue(1:200), element('O'), 6.0); rw = rdf(within(4.0, residue(1:20)), element('O'), 6.0); aa = angle(residue(1:2), residue(5:7), 30); rt = rdf(element('O'), residue(10:60), 6.0); dm = distance_min(residue(1:4), residue(100:130)); cc = contact_count(residue(1:5), residue(10:400), 4.0);"
# 4000 frames: the reference's float moving average has drifted up to ~3e-5 from the exact mean by then -> --tol 5e-5; 1000 frames at the nominal 1e-5
for t in 16 3; do
timeout 1200 oracle/_ref/shim_harness dropin --sys /tmp/w16.gro --traj synthwater:16:2024:4000 --script "$S" --threads $t --interrupt-at 1500 --tol 5e-5 > gpurun_out/r2_21_soak_4000_t$t.json 2> gpurun_out/r2_21_soak_4000_t$t.err; echo rc=$?
done
timeout 1200 oracle/_ref/shim_harness dropin --sys /tmp/w16.gro --traj synthwater:16:2024:600 --script "$S" --threads 16 --interrupt-at 200 > gpurun_out/r2_21_soak_600_t16.json 2> gpurun_out/r2_21_soak_600_t16.err; echo rc=$?
for f in gpurun_out/r2_21_soak_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], 'frames',d['frames'],'threads',d['threads'],'cpu_s',d['cpu_s'],'dropin_s',d['dropin_s'],'mask views',d['partial_mask_views'],'bad',sum(p['out_of_tol'] for p in d['properties']+d['after_restart']),'minmax',all(p['min_max_equal'] for p in d['properties']),'max_rel',max(p['max_rel'] for p in d['properties']))
PY
done
