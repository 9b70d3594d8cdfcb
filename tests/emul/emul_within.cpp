// emul_within.cpp — TEST INFRASTRUCTURE: cells.cu (cell-list build, GPU-validated) + within.cu (new) compiled by g++ and run through
// emul_launch with the launch sequences of launch_cell_list / launch_within_count: count(within(radius, selection)) per frame on the CPU.
#include "cuda_emul.h"
#include "cells_nolaunch.cu"
#include "within_nolaunch.cu"
#include <math.h>
#include <vector>

namespace mdg { void note_launch(const char*, cudaStream_t) {} }

namespace {
struct HostCellList {
    std::vector<float4> sorted, scratch; std::vector<uint32_t> cell_of, rank, cnt; mdg::CellList cl{};
    HostCellList(uint32_t B, uint32_t max_points, uint32_t cap) : sorted((size_t)B * max_points + 1), scratch((size_t)B * max_points + 1), cell_of((size_t)B * max_points + 1),
        rank((size_t)B * max_points + 1), cnt((size_t)B * (cap + 1) + B, 0u) {
        cl.sorted = sorted.data(); cl.scratch = scratch.data(); cl.cell_of = cell_of.data(); cl.rank = rank.data(); cl.cell_cnt = cnt.data();
        cl.oob = cnt.data() + (size_t)B * (cap + 1); cl.max_points = max_points; cl.cap = cap;
    }
};

// launch_cell_list, with the scan block narrowed from 1024 to 128 threads (k_scan_cells is written for any multiple of 32)
void cell_list(int mode, const mdg::BatchFrames& fr, const int32_t* idx, uint32_t n, const mdg::FrameGeom* geom, const mdg::CellList& cl) {
    const dim3 grid((n + 255u) / 256u, fr.count);
    if (n) { if (mode == 0) emul_launch(grid, dim3(256), [&]() { mdg::k_bin_points<0>(fr, idx, nullptr, n, geom, cl, 0, mdg::DynSel{}); });
             else           emul_launch(grid, dim3(256), [&]() { mdg::k_bin_points<1>(fr, idx, nullptr, n, geom, cl, 0, mdg::DynSel{}); }); }
    if (mode == 0) emul_launch(dim3(fr.count), dim3(128), [&]() { mdg::k_scan_cells<0>(geom, cl); });
    else           emul_launch(dim3(fr.count), dim3(128), [&]() { mdg::k_scan_cells<1>(geom, cl); });
    if (n) emul_launch(grid, dim3(256), [&]() { mdg::k_scatter_points(n, cl, nullptr); });
}
}  // namespace

extern "C" int emul_within_count(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                                 uint32_t num_atoms, const int32_t* sel, uint32_t n_sel, float radius, uint32_t cap, float* out) {
    mdg::BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    const double cell_ext = ceil((double)radius / 6.0) * 6.0;   // get_spatial_acc (md_script_functions.inl:734): as plan.cu passes it
    bool all_pbc = true, tri = (cells[0].flags & MDGPU_CELL_TRICLINIC) != 0;
    for (uint32_t f = 0; f < num_frames; ++f) all_pbc = all_pbc && ((cells[f].flags & MDGPU_CELL_PBC_ALL) == MDGPU_CELL_PBC_ALL);
    std::vector<float> aabb((size_t)6 * num_frames, 0.0f);
    if (!all_pbc) emul_launch(dim3(std::min((num_atoms + 255u) / 256u, 64u), num_frames), dim3(256), [&]() { mdg::k_aabb(fr, nullptr, num_atoms, aabb.data(), mdg::DynSel{}); });   // launch_aabb
    std::vector<mdg::FrameGeom> geom(num_frames); int err = 0;
    emul_launch(dim3((num_frames + 63) / 64), dim3(64), [&]() { mdg::k_frame_geom(cells, all_pbc ? nullptr : aabb.data(), geom.data(), cell_ext, (double)radius, cap, (int)num_frames, &err); });
    if (err) return err;
    HostCellList trg(num_frames, num_atoms, cap), ref(num_frames, n_sel ? n_sel : 1, cap);
    cell_list(0, fr, nullptr, num_atoms, geom.data(), trg.cl);
    cell_list(1, fr, sel, n_sel, geom.data(), ref.cl);
    std::vector<uint8_t> flags((size_t)num_frames * num_atoms, 0);
    mdg::WithinArgs a{}; a.geom = geom.data(); a.trg = trg.cl; a.ref = ref.cl; a.sel = sel; a.n_sel = n_sel; a.num_atoms = num_atoms; a.flags = flags.data(); a.out = out; a.frame0 = 0;
    if (n_sel) { if (tri) emul_launch(dim3(3, num_frames), dim3(mdg::WITHIN_WARPS * 32), [&]() { mdg::k_within_mark<true>(a); });
                 else     emul_launch(dim3(3, num_frames), dim3(mdg::WITHIN_WARPS * 32), [&]() { mdg::k_within_mark<false>(a); }); }
    emul_launch(dim3(num_frames), dim3(256), [&]() { mdg::k_within_count(a); });
    return 0;
}
