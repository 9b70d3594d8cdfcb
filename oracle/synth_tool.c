/* synth_tool.c — TEST INFRASTRUCTURE: writes the synthetic workloads of viamd_b200/csrc/synth.h to files
 * the reference harness can load (.gro topology, raw trajectory).
 *   synth_tool water-gro <n> <seed> <out.gro>
 *   synth_tool water-raw <n> <seed> <nframes> <out.raw>      (MDRAWTRJ container, see ref_harness.c)
 */
#include "../viamd_b200/csrc/synth.h"
#include <stdlib.h>
#include <string.h>

int main(int argc, char** argv) {
    if (argc >= 5 && strcmp(argv[1], "water-gro") == 0) {
        mdsynth_water_t w = mdsynth_water_desc((uint32_t)atoi(argv[2]), (uint32_t)atoi(argv[3]));
        float* p = malloc((size_t)w.num_atoms * 4 * 3);
        mdsynth_water_base(&w, NULL, NULL, NULL, p, p + w.num_atoms, p + 2 * w.num_atoms);
        return mdsynth_water_write_gro(&w, argv[4], p, p + w.num_atoms, p + 2 * w.num_atoms) ? 2 : 0;
    }
    if (argc >= 6 && strcmp(argv[1], "water-raw") == 0) {
        mdsynth_water_t w = mdsynth_water_desc((uint32_t)atoi(argv[2]), (uint32_t)atoi(argv[3]));
        const uint64_t nf = (uint64_t)atoll(argv[4]), na = w.num_atoms;
        float* b = malloc(na * 12), *x = malloc(na * 12);
        mdsynth_water_base(&w, b, b + na, b + 2 * na, NULL, NULL, NULL);
        FILE* f = fopen(argv[5], "wb"); if (!f) return 2;
        fwrite("MDRAWTRJ", 1, 8, f); fwrite(&nf, 8, 1, f); fwrite(&na, 8, 1, f);
        for (uint64_t fr = 0; fr < nf; ++fr) {
            double cell[6] = { w.L, 0, 0, w.L, 0, w.L }; uint32_t fl[2] = { 1 | 4 | 8 | 16, 0 };
            mdsynth_water_frame(&w, (uint32_t)fr, b, b + na, b + 2 * na, x, x + na, x + 2 * na);
            fwrite(cell, 8, 6, f); fwrite(fl, 4, 2, f); fwrite(x, 4, na * 3, f);
        }
        fclose(f); return 0;
    }
    if (argc >= 7 && strcmp(argv[1], "membrane-gro") == 0) {
        mdsynth_membrane_t m = mdsynth_membrane_desc((uint32_t)atoi(argv[2]), (uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]));
        float* p = malloc((size_t)m.num_atoms * 12);
        mdsynth_membrane_base(&m, NULL, p, NULL);
        return mdsynth_membrane_write_gro(&m, argv[6], p) ? 2 : 0;
    }
    if (argc >= 8 && strcmp(argv[1], "membrane-raw") == 0) {
        mdsynth_membrane_t m = mdsynth_membrane_desc((uint32_t)atoi(argv[2]), (uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]));
        const uint64_t nf = (uint64_t)atoll(argv[6]), na = m.num_atoms;
        float* b = malloc(na * 12), *x = malloc(na * 12); uint32_t* mol = malloc(na * 4);
        mdsynth_membrane_base(&m, b, NULL, mol);
        FILE* f = fopen(argv[7], "wb"); if (!f) return 2;
        fwrite("MDRAWTRJ", 1, 8, f); fwrite(&nf, 8, 1, f); fwrite(&na, 8, 1, f);
        for (uint64_t fr = 0; fr < nf; ++fr) {
            double cell[6] = { m.Lx, 0, 0, m.Ly, 0, m.Lz }; uint32_t fl[2] = { 1 | 4 | 8 | 16, 0 };
            mdsynth_membrane_frame(&m, (uint32_t)fr, b, mol, x, x + na, x + 2 * na);
            fwrite(cell, 8, 6, f); fwrite(fl, 4, 2, f); fwrite(x, 4, na * 3, f);
        }
        fclose(f); return 0;
    }
    fprintf(stderr, "usage: synth_tool water-gro n seed out.gro | water-raw n seed nframes out.raw | membrane-gro nl nwxy nwz seed out.gro | membrane-raw nl nwxy nwz seed nframes out.raw\n");
    return 1;
}
