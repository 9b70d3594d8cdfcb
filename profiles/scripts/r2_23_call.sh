set -x
mkdir -p gpurun_out
timeout 600 python profiles/xtc_e2e.py > gpurun_out/r2_23_xtc_e2e_rdf_sdf.json 2>gpurun_out/r2_23_xtc.err; tail -1 gpurun_out/r2_23_xtc_e2e_rdf_sdf.json | cut -c1-600; tail -2 gpurun_out/r2_23_xtc.err
timeout 600 python profiles/xtc_e2e.py --script "d = distance(1,10); dz = density_z(element('O'));" > gpurun_out/r2_23_xtc_e2e_light_script.json 2>>gpurun_out/r2_23_xtc.err; tail -1 gpurun_out/r2_23_xtc_e2e_light_script.json | cut -c1-600
timeout 600 python profiles/xtc_e2e.py --script "r = rdf(element('O'), element('O'), 10.0);" > gpurun_out/r2_23_xtc_e2e_rdf_only.json 2>>gpurun_out/r2_23_xtc.err; tail -1 gpurun_out/r2_23_xtc_e2e_rdf_only.json | cut -c1-600
