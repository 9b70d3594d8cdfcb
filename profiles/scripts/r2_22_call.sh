set -x
mkdir -p gpurun_out
oracle/build/synth_tool water-gro 16 2024 /tmp/w16.gro
S="r = rdf(element('O'), element('O'), 8.0); d = distance(1,10); dz = density_z(element('O')); v = sdf(residue(1:200), element('O'), 6.0);"
for cfg in "1 2" "1 16" "1 148" "16 2" "16 16"; do set -- $cfg
timeout 600 oracle/_ref/shim_harness dropin --sys /tmp/w16.gro --traj synthwater:16:2024:1184 --script "$S" --threads $1 --chunk $2 --tol 5e-5 > gpurun_out/r2_22_t$1_c$2.json 2>/dev/null
python - gpurun_out/r2_22_t$1_c$2.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); print('threads',d['threads'],'chunk',d['chunk'],'cpu_s',d['cpu_s'],'dropin_s',d['dropin_s'],'bad',sum(p['out_of_tol'] for p in d['properties']))
PY
done
nsys --version 2>/dev/null | head -1
