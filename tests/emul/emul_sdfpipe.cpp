// emul_sdfpipe.cpp — TEST INFRASTRUCTURE: the whole sdf() pipeline of the product (cells.cu cell list of the targets, then k_sdf_ref0,
// k_sdf_fit, k_sdf_scatter<TRI> of sdf.cu) compiled by g++ and run through emul_launch with the launch shapes of launch_cell_list / launch_sdf.
// All of these kernels have passed on the GPU; running them here makes the CPU test stage exercise the same source.
#include "cuda_emul.h"
#include "cells_nolaunch.cu"
#include "sdf_nolaunch.cu"
#include <vector>

namespace mdg { void note_launch(const char*, cudaStream_t) {} }

extern "C" int emul_sdf(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                        const float* init_xyz, size_t init_axis_stride, const float* mass, const int32_t* struct_idx, uint32_t n_struct, uint32_t struct_size,
                        const int32_t* trg_idx, uint32_t n_trg, const int32_t* pairs, uint32_t n_pairs, float cutoff, uint32_t cap,
                        uint32_t* vol /* [128^3], accumulated */, unsigned long long* frame_total /* [num_frames] */) {
    using namespace mdg;
    BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    const bool tri = (cells[0].flags & MDGPU_CELL_TRICLINIC) != 0;
    std::vector<FrameGeom> geom(num_frames); int err = 0;
    emul_launch(dim3((num_frames + 63) / 64), dim3(64), [&]() { k_frame_geom(cells, nullptr, geom.data(), (double)cutoff, (double)cutoff, cap, (int)num_frames, &err); });
    if (err) return err;
    const size_t np = (size_t)num_frames * n_trg + 1;
    std::vector<float4> sorted(np), scratch(np); std::vector<uint32_t> cell_of(np), rank(np), cnt((size_t)num_frames * (cap + 1) + num_frames, 0u);
    CellList cl{}; cl.sorted = sorted.data(); cl.scratch = scratch.data(); cl.cell_of = cell_of.data(); cl.rank = rank.data(); cl.cell_cnt = cnt.data();
    cl.oob = cnt.data() + (size_t)num_frames * (cap + 1); cl.max_points = n_trg; cl.cap = cap;
    const dim3 grid((n_trg + 255u) / 256u, num_frames);   // launch_cell_list(0, ...), scan block narrowed to 128 threads
    emul_launch(grid, dim3(256), [&]() { k_bin_points<0>(fr, trg_idx, nullptr, n_trg, geom.data(), cl, 0, mdg::DynSel{}); });
    emul_launch(dim3(num_frames), dim3(128), [&]() { k_scan_cells<0>(geom.data(), cl); });
    emul_launch(grid, dim3(256), [&]() { k_scatter_points(n_trg, cl, nullptr); });

    std::vector<float4> xyzw((size_t)num_frames * (n_struct + 1) * struct_size);
    std::vector<float> ref0((size_t)num_frames * 20), mats((size_t)num_frames * n_struct * 32);
    SdfArgs a{};
    a.geom = geom.data(); a.trg = cl; a.frames = fr; a.cells = cells; a.init_xyz = init_xyz; a.init_axis_stride = init_axis_stride; a.mass = mass;
    a.struct_idx = struct_idx; a.n_struct = n_struct; a.struct_size = struct_size; a.unwrap_pairs = (const int2*)pairs; a.n_unwrap = n_pairs; a.cutoff = cutoff;
    a.scratch_xyzw = xyzw.data(); a.ref0 = ref0.data(); a.matrices = mats.data(); a.vol = vol; a.frame_total = frame_total; a.frame0 = 0;
    const int B = (int)num_frames;   // launch_sdf
    emul_launch(dim3((B + 31) / 32), dim3(32), [&]() { k_sdf_ref0(a, B); });
    emul_launch(dim3((n_struct + 63) / 64, B), dim3(64), [&]() { k_sdf_fit(a, B); });
    if (tri) emul_launch(dim3((n_struct + SDF_WARPS - 1) / SDF_WARPS, B), dim3(SDF_WARPS * 32), [&]() { k_sdf_scatter<true, 4>(a, B); });
    else     emul_launch(dim3((n_struct + SDF_WARPS - 1) / SDF_WARPS, B), dim3(SDF_WARPS * 32), [&]() { k_sdf_scatter<false, 4>(a, B); });
    return 0;
}
