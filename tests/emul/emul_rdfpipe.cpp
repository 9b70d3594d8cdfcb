// emul_rdfpipe.cpp — TEST INFRASTRUCTURE: the rdf() hot path of the product — cell lists (cells.cu), group centres (props.cu), k_rdf_cull,
// k_rdf_pairs_v2 (packed pair loop, hit queue, symmetric counting), the scalar k_rdf_pairs (exclusion path) and k_rdf_finalize (rdf.cu) —
// compiled by g++ and run through emul_launch in the order of plan.cu / launch_rdf. rdf.cu's PTX helpers are swapped for plain C++ by
// build_emul.py (RDF_PATCHES). Grids are narrowed (both kernels stride over the home cells), block shapes are the product's.
#include "cuda_emul.h"
#include "cells_nolaunch.cu"
#include "props_nolaunch.cu"
#include "rdf_nolaunch.cu"
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
void cell_list(int mode, const mdg::BatchFrames& fr, const int32_t* idx, const float* aos, uint32_t n, const mdg::FrameGeom* geom, const mdg::CellList& cl) {
    const dim3 grid((n + 255u) / 256u, fr.count);
    if (n) { if (mode == 0) emul_launch(grid, dim3(256), [&]() { mdg::k_bin_points<0>(fr, idx, aos, n, geom, cl, 0, mdg::DynSel{}); });
             else           emul_launch(grid, dim3(256), [&]() { mdg::k_bin_points<1>(fr, idx, aos, n, geom, cl, 0, mdg::DynSel{}); }); }
    if (mode == 0) emul_launch(dim3(fr.count), dim3(128), [&]() { mdg::k_scan_cells<0>(geom, cl); });
    else           emul_launch(dim3(fr.count), dim3(128), [&]() { mdg::k_scan_cells<1>(geom, cl); });
    if (n) emul_launch(grid, dim3(256), [&]() { mdg::k_scatter_points(n, cl, nullptr); });
}
}  // namespace

// ref_idx/n_ref: reference atoms; or, when n_groups > 0, the atoms of n_groups groups (CSR offsets group_off) whose centres of mass are the
// references and whose own atoms are excluded from their pairs. keep: [num_frames][1024] per-frame bins; totals: [num_frames].
extern "C" int emul_rdf(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells, const float* mass,
                        const int32_t* ref_idx, uint32_t n_ref, const uint32_t* group_off, uint32_t n_groups, const int32_t* trg_idx, uint32_t n_trg,
                        float cutoff_min, float cutoff_max, int symmetric, uint32_t cap, uint32_t* keep, unsigned long long* totals) {
    using namespace mdg;
    const int B = (int)num_frames;
    BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    const bool tri = (cells[0].flags & MDGPU_CELL_TRICLINIC) != 0;
    bool all_pbc = true; for (int f = 0; f < B; ++f) all_pbc = all_pbc && ((cells[f].flags & MDGPU_CELL_PBC_ALL) == MDGPU_CELL_PBC_ALL);
    std::vector<float> aabb((size_t)6 * B, 0.0f);
    if (!all_pbc) emul_launch(dim3(std::min((n_trg + 255u) / 256u, 64u), B), dim3(256), [&]() { k_aabb(fr, trg_idx, n_trg, aabb.data(), mdg::DynSel{}); });
    std::vector<FrameGeom> geom(B); int err = 0;
    emul_launch(dim3((B + 63) / 64), dim3(64), [&]() { k_frame_geom(cells, all_pbc ? nullptr : aabb.data(), geom.data(), (double)cutoff_max, (double)cutoff_max, cap, B, &err); });
    if (err) return err;
    HostCellList trg(B, n_trg, cap), ref(B, n_groups ? n_groups : n_ref, cap);
    cell_list(0, fr, trg_idx, nullptr, n_trg, geom.data(), trg.cl);
    std::vector<float> com((size_t)B * (n_groups ? n_groups : 1) * 3);
    if (n_groups) {
        emul_launch(dim3((n_groups + 127u) / 128u, B), dim3(128), [&]() { k_group_com(fr, ref_idx, group_off, n_groups, mass, com.data()); });
        cell_list(1, fr, nullptr, com.data(), n_groups, geom.data(), ref.cl);
    } else cell_list(1, fr, ref_idx, nullptr, n_ref, geom.data(), ref.cl);

    const size_t list_stride = (size_t)125 * n_trg + 1024;
    std::vector<uint32_t> frame_bins((size_t)B * (MDGPU_DIST_BINS + 1), 0u), pair_list((size_t)B * list_stride), cursor(B, 0u), fmin(B), fmax(B);
    std::vector<uint4> hdr((size_t)B * cap); std::vector<unsigned long long> acc(MDGPU_DIST_BINS, 0ull);
    RdfArgs a{};
    a.geom = geom.data(); a.trg = trg.cl; a.ref = ref.cl;
    a.inv_cutoff_range = 1.0f / (cutoff_max - cutoff_min); a.min_cutoff = cutoff_min > 1e-3f ? cutoff_min : 1e-3f; a.min_r2 = a.min_cutoff * a.min_cutoff;   // plan.cu
    a.frame_bins = frame_bins.data(); a.frame0 = 0;
    a.pair_list = pair_list.data(); a.list_hdr = hdr.data(); a.list_cursor = cursor.data(); a.list_stride = list_stride; a.hdr_stride = cap; a.err = &err;
    a.excl_off = n_groups ? group_off : nullptr; a.excl_idx = n_groups ? ref_idx : nullptr; a.symmetric = symmetric;
    a.acc = acc.data(); a.frame_total = totals; a.frame_min = fmin.data(); a.frame_max = fmax.data(); a.keep = keep;
    if (!n_groups) {   // launch_rdf, default variant
        if (tri) emul_launch(dim3(4, B), dim3(CULL_WARPS * 32), [&]() { k_rdf_cull<true>(a); }); else emul_launch(dim3(4, B), dim3(CULL_WARPS * 32), [&]() { k_rdf_cull<false>(a); });
        if (err) return err;
        if (tri) emul_launch(dim3(2, B), dim3(V2_THREADS), [&]() { k_rdf_pairs_v2<true, 0>(a); }); else emul_launch(dim3(2, B), dim3(V2_THREADS), [&]() { k_rdf_pairs_v2<false, 0>(a); });
    } else {           // exclusion path: scalar kernel
        if (tri) emul_launch(dim3(2, B), dim3(RDF_THREADS), [&]() { k_rdf_pairs<true, true, false>(a); }); else emul_launch(dim3(2, B), dim3(RDF_THREADS), [&]() { k_rdf_pairs<false, true, false>(a); });
    }
    emul_launch(dim3(B), dim3(MDGPU_DIST_BINS), [&]() { k_rdf_finalize(a); });
    return err;
}
