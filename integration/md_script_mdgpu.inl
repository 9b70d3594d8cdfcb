/* md_script_mdgpu.inl — the reference-side binding of libmdgpu (include/mdgpu.h).
 *
 * HOW A MAINTAINER USES IT: add `#include "md_script_mdgpu.inl"` at the end of mdlib/src/md_script.c (it needs
 * the translation unit's private types: md_script_ir_t, ast_node_t, data_t — the same trick the reference's own white-box tests use,
 * mdlib/unittest/test_script.c:20) and call md_script_gpu_eval_frame_range where VIAMD calls md_script_eval_frame_range
 * (src/main.cpp:993-997, 1029-1033). Nothing else in mdlib changes: the tokenizer, parser, static type check and static evaluation
 * of selections stay as they are; this file only LOWERS the compiled IR's property statements to flat descriptors.
 *
 * Lowering rule: a property statement `ident = proc(args);` is accepted when proc is on the GPU hot path (rdf, sdf, density_x/_y/_z,
 * distance, angle, dihedral) and every argument was evaluated statically at compile time (FLAG_CONSTANT, md_script.c:5492-5524), i.e.
 * selections are fixed atom sets. Bitfields become ascending atom index lists (md_bitfield_iter_extract_indices), 1-based script atom
 * indices become 0-based. Anything else is reported through MD_LOG_ERROR and the call returns false — there is no silent CPU fallback.
 *
 * This file is written against the reference's private API on purpose and contains no reference code.
 */
#include <mdgpu.h>
#include <stddef.h>

/* The C ABI passes the reference's own structs through casts (md_unitcell_t, the frame header, the trajectory interfaces): pin their
 * layouts against this mdlib at compile time, so a change upstream breaks the build here instead of corrupting frames at run time. */
_Static_assert(sizeof(mdgpu_unitcell_t) == sizeof(md_unitcell_t), "md_unitcell_t layout");
_Static_assert(offsetof(mdgpu_unitcell_t, x) == offsetof(md_unitcell_t, x) && offsetof(mdgpu_unitcell_t, xy) == offsetof(md_unitcell_t, xy) &&
               offsetof(mdgpu_unitcell_t, xz) == offsetof(md_unitcell_t, xz) && offsetof(mdgpu_unitcell_t, y) == offsetof(md_unitcell_t, y) &&
               offsetof(mdgpu_unitcell_t, yz) == offsetof(md_unitcell_t, yz) && offsetof(mdgpu_unitcell_t, z) == offsetof(md_unitcell_t, z) &&
               offsetof(mdgpu_unitcell_t, flags) == offsetof(md_unitcell_t, flags), "md_unitcell_t fields");
_Static_assert(MDGPU_CELL_ORTHO == MD_UNITCELL_ORTHO && MDGPU_CELL_TRICLINIC == MD_UNITCELL_TRICLINIC && MDGPU_CELL_PBC_X == MD_UNITCELL_PBC_X &&
               MDGPU_CELL_PBC_Y == MD_UNITCELL_PBC_Y && MDGPU_CELL_PBC_Z == MD_UNITCELL_PBC_Z, "md_unitcell_flags_t values");
_Static_assert(sizeof(mdgpu_frame_header_t) == sizeof(md_trajectory_frame_header_t) &&
               offsetof(mdgpu_frame_header_t, num_atoms) == offsetof(md_trajectory_frame_header_t, num_atoms) &&
               offsetof(mdgpu_frame_header_t, index) == offsetof(md_trajectory_frame_header_t, index) &&
               offsetof(mdgpu_frame_header_t, timestamp) == offsetof(md_trajectory_frame_header_t, timestamp) &&
               offsetof(mdgpu_frame_header_t, unitcell) == offsetof(md_trajectory_frame_header_t, unitcell), "md_trajectory_frame_header_t layout");
_Static_assert(sizeof(mdgpu_trajectory_reader_i) == sizeof(md_trajectory_reader_i) && offsetof(mdgpu_trajectory_reader_i, load_frame) == offsetof(md_trajectory_reader_i, load_frame) &&
               offsetof(mdgpu_trajectory_reader_i, free) == offsetof(md_trajectory_reader_i, free), "md_trajectory_reader_i layout");
_Static_assert(sizeof(mdgpu_trajectory_i) == sizeof(md_trajectory_i) && offsetof(mdgpu_trajectory_i, get_header) == offsetof(md_trajectory_i, get_header) &&
               offsetof(mdgpu_trajectory_i, init_reader) == offsetof(md_trajectory_i, init_reader), "md_trajectory_i layout");
_Static_assert(sizeof(mdgpu_trajectory_header_t) == sizeof(md_trajectory_header_t) && offsetof(mdgpu_trajectory_header_t, num_frames) == offsetof(md_trajectory_header_t, num_frames) &&
               offsetof(mdgpu_trajectory_header_t, num_atoms) == offsetof(md_trajectory_header_t, num_atoms) &&
               offsetof(mdgpu_trajectory_header_t, frame_times) == offsetof(md_trajectory_header_t, frame_times), "md_trajectory_header_t layout");

typedef struct md_script_gpu_lowered_t {
    size_t num_props;
    mdgpu_property_desc_t* props;   /* arena-allocated, as are the index lists they point to */
    char (*names)[64];
} md_script_gpu_lowered_t;

static const ast_node_t* mdgpu__rhs(const ast_node_t* node) {
    /* property nodes are the assignment `ident = expr` (extract_properties, md_script.c:6173-6193) */
    if (node->type == AST_ASSIGNMENT && node->children && md_array_size(node->children) == 2) return node->children[1];
    return node;
}

/* static argument -> ascending 0-based atom index list (concatenated over all bitfields of an array); returns count, -1 on error */
static int64_t mdgpu__arg_indices(int32_t** out, size_t* out_num_sets, size_t* out_set_size, const ast_node_t* arg, md_allocator_i* alloc) {
    if (!(arg->flags & FLAG_CONSTANT) || !arg->data.ptr) return -1;
    const data_t d = arg->data;
    if (d.type.base_type == TYPE_BITFIELD) {
        const md_bitfield_t* bf = (const md_bitfield_t*)d.ptr;
        const size_t n = element_count(d);
        size_t total = 0; size_t set_size = 0; bool uniform = true;
        for (size_t i = 0; i < n; ++i) { const size_t c = md_bitfield_popcount(&bf[i]); if (i == 0) set_size = c; else if (c != set_size) uniform = false; total += c; }
        int32_t* idx = (int32_t*)md_alloc(alloc, sizeof(int32_t) * (total ? total : 1));
        size_t off = 0;
        for (size_t i = 0; i < n; ++i) { const size_t c = md_bitfield_popcount(&bf[i]); md_bitfield_iter_extract_indices(idx + off, c, md_bitfield_iter_create(&bf[i])); off += c; }
        *out = idx; if (out_num_sets) *out_num_sets = n; if (out_set_size) *out_set_size = uniform ? set_size : 0;
        return (int64_t)total;
    }
    if (d.type.base_type == TYPE_INT) {
        const int32_t* v = (const int32_t*)d.ptr;
        const size_t n = element_count(d);
        int32_t* idx = (int32_t*)md_alloc(alloc, sizeof(int32_t) * (n ? n : 1));
        for (size_t i = 0; i < n; ++i) idx[i] = v[i] - 1;   /* remap_index_to_context with the whole system as context (md_script_functions.inl:1023) */
        *out = idx; if (out_num_sets) *out_num_sets = 1; if (out_set_size) *out_set_size = n;
        return (int64_t)n;
    }
    return -1;
}

/* CSR offsets of the bitfields of a static array argument inside the concatenated list mdgpu__arg_indices returns (n + 1 entries) */
static uint32_t* mdgpu__arg_part_offsets(const ast_node_t* arg, md_allocator_i* alloc) {
    const md_bitfield_t* bf = (const md_bitfield_t*)arg->data.ptr;
    const size_t n = element_count(arg->data);
    uint32_t* off = (uint32_t*)md_alloc(alloc, sizeof(uint32_t) * (n + 1));
    off[0] = 0;
    for (size_t i = 0; i < n; ++i) off[i + 1] = off[i] + (uint32_t)md_bitfield_popcount(&bf[i]);
    return off;
}

/* _internal_flatten_bf: union of an array of bitfields. The concatenated index lists of disjoint sets are already that union; sort + unique
 * covers overlapping ones. Returns the new count. */
static size_t mdgpu__flatten(int32_t* v, size_t n) {
    for (size_t i = 1; i < n; ++i) { int32_t t = v[i]; size_t j = i; while (j > 0 && v[j - 1] > t) { v[j] = v[j - 1]; --j; } v[j] = t; }
    size_t w = 0; for (size_t i = 0; i < n; ++i) if (w == 0 || v[i] != v[w - 1]) v[w++] = v[i];
    return w;
}

/* `within([min:]max, selection)` or `static_selection and within(...)` (either order; _and md_script_functions.inl:1975 flattens arrays): the
 * dynamic selection the device evaluates per frame. Returns the within() call and, through out_mask, the static side (NULL if none). */
static const ast_node_t* mdgpu__within_expr(const ast_node_t* node, const ast_node_t** out_mask) {
    *out_mask = NULL;
    if (node->type != AST_PROC_CALL || !node->proc || md_array_size(node->children) != 2) return NULL;
    if (str_eq(node->proc->name, STR_LIT("within"))) return node;
    if (str_eq(node->proc->name, STR_LIT("and"))) {
        for (int side = 0; side < 2; ++side) {
            const ast_node_t* w = node->children[side]; const ast_node_t* m = node->children[1 - side];
            if (w->type == AST_PROC_CALL && w->proc && str_eq(w->proc->name, STR_LIT("within")) && md_array_size(w->children) == 2 &&
                (m->flags & FLAG_CONSTANT) && m->data.type.base_type == TYPE_BITFIELD) { *out_mask = m; return w; }
        }
    }
    return NULL;
}

/* fills radius (and lower bound), the within() selection into idx[0] and the optional AND mask into idx[2] + com_args bit 0 */
static bool mdgpu__lower_within(mdgpu_property_desc_t* out, const ast_node_t* w, const ast_node_t* mask, float* rmin, float* rmax, md_allocator_i* alloc) {
    ast_node_t** c = w->children; size_t ns = 0; int64_t n;
    if (!(c[0]->flags & FLAG_CONSTANT) || c[1]->data.type.base_type != TYPE_BITFIELD) return false;
    if (c[0]->data.type.base_type == TYPE_FRANGE) { const frange_t r = *(const frange_t*)c[0]->data.ptr; *rmin = r.beg; *rmax = r.end; }   /* _within_expl_frng :2609 */
    else if (c[0]->data.type.base_type == TYPE_FLOAT) { *rmin = 0.0f; *rmax = *(const float*)c[0]->data.ptr; }
    else return false;
    if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], &ns, NULL, c[1], alloc)) < 0) return false;
    out->idx_count[0] = mdgpu__flatten((int32_t*)out->idx[0], (size_t)n);   /* within is FLAG_FLATTEN (:673): an array of selections is their union */
    if (mask) {
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[2], NULL, NULL, mask, alloc)) < 0) return false;
        out->idx_count[2] = mdgpu__flatten((int32_t*)out->idx[2], (size_t)n); out->com_args |= 1u;
    }
    return true;
}

/* argument k of a property: a static selection, or a dynamic one (within([min:]max, sel) [and static]) -> idx[k] (+ dyn[k]). Returns the number of
 * bitfields of a static array argument through out_sets (1 for dynamic ones), or -1 when the argument is neither. */
static int64_t mdgpu__lower_sel_arg(mdgpu_property_desc_t* out, int k, const ast_node_t* arg, size_t* out_sets, md_allocator_i* alloc) {
    const ast_node_t* wmask = NULL; const ast_node_t* wnode = mdgpu__within_expr(arg, &wmask);
    if (wnode) {
        ast_node_t** c = wnode->children; size_t ns = 0; int64_t n;
        mdgpu_dynamic_arg_t* dy = &out->dyn[k];
        if (!(c[0]->flags & FLAG_CONSTANT) || c[1]->data.type.base_type != TYPE_BITFIELD) return -1;
        if (c[0]->data.type.base_type == TYPE_FRANGE) { const frange_t r = *(const frange_t*)c[0]->data.ptr; dy->radius_min = r.beg; dy->radius_max = r.end; }   /* _within_expl_frng :2609 */
        else if (c[0]->data.type.base_type == TYPE_FLOAT) { dy->radius_min = 0.0f; dy->radius_max = *(const float*)c[0]->data.ptr; }
        else return -1;
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[k], &ns, NULL, c[1], alloc)) < 0) return -1;
        out->idx_count[k] = mdgpu__flatten((int32_t*)out->idx[k], (size_t)n);   /* within is FLAG_FLATTEN (:673) */
        if (wmask) {
            int32_t* m = NULL;
            if ((n = mdgpu__arg_indices(&m, NULL, NULL, wmask, alloc)) < 0) return -1;
            dy->and_idx = m; dy->and_count = mdgpu__flatten(m, (size_t)n); dy->has_and = 1u;
        }
        if (out_sets) *out_sets = 1;
        return (int64_t)out->idx_count[k];
    }
    {
        int64_t n = mdgpu__arg_indices((int32_t**)&out->idx[k], out_sets, NULL, arg, alloc);
        if (n >= 0) out->idx_count[k] = (size_t)n;
        return n;
    }
}

/* the script and system being lowered (md_script_gpu_lower_sys sets them): context-relative arguments are evaluated with the reference's own evaluator */
static _Thread_local const md_script_ir_t* mdgpu__lower_ir = NULL;
static _Thread_local const md_system_t* mdgpu__lower_mol = NULL;

static bool mdgpu__lower_property(mdgpu_property_desc_t* out, str_t ident, const ast_node_t* node, const md_system_t* mol, md_allocator_i* alloc) {
    const ast_node_t* rhs = mdgpu__rhs(node);
    const md_bitfield_t* ctx_bf = NULL; size_t n_ctx = 0;
    if (rhs->type == AST_CONTEXT && rhs->children && md_array_size(rhs->children) == 2) {   /* `expr in contexts` (evaluate_context md_script.c:3418) */
        const ast_node_t* cn = rhs->children[1];
        if (!cn->data.ptr || cn->data.type.base_type != TYPE_BITFIELD) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': the contexts of `in` are not a static selection array", STR_ARG(ident)); return false; }
        ctx_bf = (const md_bitfield_t*)cn->data.ptr; n_ctx = element_count(cn->data); rhs = rhs->children[0];
        if (!n_ctx) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': no contexts", STR_ARG(ident)); return false; }
    }
    if (rhs->type != AST_PROC_CALL || !rhs->proc) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "' is not a direct procedure call", STR_ARG(ident)); return false; }
    const str_t pname = rhs->proc->name;
    if (n_ctx && !(str_eq(pname, STR_LIT("distance")) || str_eq(pname, STR_LIT("angle")) || str_eq(pname, STR_LIT("dihedral")))) {
        MD_LOG_ERROR("mdgpu: property '" STR_FMT "': `in` is lowered for distance / angle / dihedral only", STR_ARG(ident)); return false;
    }
    const size_t nargs = md_array_size(rhs->children);
    ast_node_t** args = rhs->children;
    memset(out, 0, sizeof(*out));
    size_t nsets = 0, set_size = 0; int64_t n;
    if (str_eq(pname, STR_LIT("rdf")) && nargs == 3) {
        out->op = MDGPU_OP_RDF;
        const ast_node_t* wmask = NULL; const ast_node_t* wnode = mdgpu__within_expr(args[0], &wmask);
        if (wnode) {   /* dynamic reference set (_within_expl_flt :2485 / _frng :2609), optionally `and` a static selection: evaluated per frame on the device */
            if (!mdgpu__lower_within(out, wnode, wmask, &out->ref_within_min, &out->ref_within_radius, alloc)) goto dynamic;
        } else
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], &nsets, &set_size, args[0], alloc)) < 0) goto dynamic; else out->idx_count[0] = (size_t)n;
        if (args[0]->data.type.base_type == TYPE_BITFIELD && nsets > 1) {   /* array of bitfields: COM references + exclusion masks (:5275) */
            const md_bitfield_t* bf = (const md_bitfield_t*)args[0]->data.ptr;
            uint32_t* off = (uint32_t*)md_alloc(alloc, sizeof(uint32_t) * (nsets + 1));
            off[0] = 0; for (size_t i = 0; i < nsets; ++i) off[i + 1] = off[i] + (uint32_t)md_bitfield_popcount(&bf[i]);
            out->num_structures = nsets; out->structure_size = set_size; out->structure_offsets = off;
        }
        {
            size_t tsets = 0;
            if ((n = mdgpu__lower_sel_arg(out, 1, args[1], &tsets, alloc)) < 0) goto dynamic;
            if (!(out->dyn[1].radius_max > 0.0f) && args[1]->data.type.base_type == TYPE_BITFIELD && tsets > 1) {   /* one centre of mass per bitfield is the target point (coordinate_extract :1503, compute_rdf :5299) */
                if (wnode) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': a dynamic reference set with an array of selections as target is not lowered", STR_ARG(ident)); return false; }
                out->structure_offsets_b = mdgpu__arg_part_offsets(args[1], alloc); out->num_structures_b = tsets;
            }
        }
        if (!(args[2]->flags & FLAG_CONSTANT)) goto dynamic;
        if (args[2]->data.type.base_type == TYPE_FRANGE) { const frange_t r = *(const frange_t*)args[2]->data.ptr; out->cutoff_min = r.beg; out->cutoff_max = r.end; }
        else { out->cutoff_min = 0.0f; out->cutoff_max = *(const float*)args[2]->data.ptr; }
        return true;
    }
    if (str_eq(pname, STR_LIT("sdf")) && nargs == 3) {
        out->op = MDGPU_OP_SDF;
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], &nsets, &set_size, args[0], alloc)) < 0) goto dynamic; out->idx_count[0] = (size_t)n;
        out->num_structures = nsets; out->structure_size = set_size;
        if ((n = mdgpu__lower_sel_arg(out, 1, args[1], NULL, alloc)) < 0) goto dynamic;
        if (!(args[2]->flags & FLAG_CONSTANT)) goto dynamic;
        out->cutoff_max = *(const float*)args[2]->data.ptr;
        return true;
    }
    if ((str_eq(pname, STR_LIT("density_x")) || str_eq(pname, STR_LIT("density_y")) || str_eq(pname, STR_LIT("density_z"))) && nargs == 1) {
        out->op = MDGPU_OP_DENSITY_X + (uint32_t)(pname.ptr[8] - 'x');
        if ((n = mdgpu__lower_sel_arg(out, 0, args[0], NULL, alloc)) < 0) goto dynamic;
        out->idx_count[0] = mdgpu__flatten((int32_t*)out->idx[0], out->idx_count[0]);
        return true;
    }
    if ((str_eq(pname, STR_LIT("coord_x")) || str_eq(pname, STR_LIT("coord_y")) || str_eq(pname, STR_LIT("coord_z"))) && nargs == 1) {   /* _coordinate_x/_y/_z :5077 */
        size_t ns = 0;
        out->op = MDGPU_OP_COORD_X + (uint32_t)(pname.ptr[6] - 'x');
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], &ns, NULL, args[0], alloc)) < 0) goto dynamic; out->idx_count[0] = (size_t)n;
        if (args[0]->data.type.base_type == TYPE_BITFIELD && ns > 1) {   /* one value per selection: the coordinate of its centre of mass (coordinate_extract :1503) */
            out->structure_offsets = mdgpu__arg_part_offsets(args[0], alloc); out->num_structures = ns;
        }
        return true;
    }
    if ((str_eq(pname, STR_LIT("com")) || str_eq(pname, STR_LIT("plane"))) && nargs == 1) {   /* _com :4726, _plane :4755: [F,3] / [F,4] temporals */
        size_t ns = 0;
        const bool is_com = str_eq(pname, STR_LIT("com"));
        out->op = is_com ? MDGPU_OP_COM : MDGPU_OP_PLANE;
        if (is_com) { if ((n = mdgpu__lower_sel_arg(out, 0, args[0], &ns, alloc)) < 0) goto dynamic; if (out->dyn[0].radius_max > 0.0f) { out->com_args = 1u; return true; } }
        else { if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], &ns, NULL, args[0], alloc)) < 0) goto dynamic; out->idx_count[0] = (size_t)n; }
        if (args[0]->data.type.base_type == TYPE_BITFIELD) {
            if (ns > 1) {
                if (!is_com) { out->structure_offsets = mdgpu__arg_part_offsets(args[0], alloc); out->num_structures = ns; }   /* plane through the selections' centres of mass (coordinate_extract :1503) */
                else { out->arg_offsets[0] = mdgpu__arg_part_offsets(args[0], alloc); out->arg_parts[0] = (uint32_t)ns; }   /* the centre of the selections' centres (coordinate_extract_com :1826-1842) */
            }
            if (is_com) out->com_args = 1u;
        }
        return true;
    }
    if (str_eq(pname, STR_LIT("count")) && nargs == 1) {   /* count(<within expression>): _count :2868 on the per-frame selection */
        const ast_node_t* wmask = NULL; const ast_node_t* wnode = mdgpu__within_expr(args[0], &wmask);
        if (!wnode) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': count() is lowered for within(...) expressions only", STR_ARG(ident)); return false; }
        out->op = MDGPU_OP_WITHIN_COUNT;
        if (!mdgpu__lower_within(out, wnode, wmask, &out->cutoff_min, &out->cutoff_max, alloc)) goto dynamic;
        return true;
    }
    if (str_eq(pname, STR_LIT("contact_count")) && (nargs == 3 || nargs == 4)) {   /* _contact_count :2756-2866 */
        size_t na_sets = 0; int64_t nb;
        out->op = MDGPU_OP_CONTACT_COUNT;
        if (!mol) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': contact_count needs the system (bonds) to build its exclusion masks", STR_ARG(ident)); return false; }
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], &na_sets, NULL, args[0], alloc)) < 0) goto dynamic; out->idx_count[0] = (size_t)n;
        if ((nb = mdgpu__arg_indices((int32_t**)&out->idx[1], NULL, NULL, args[1], alloc)) < 0) goto dynamic;
        out->idx_count[1] = mdgpu__flatten((int32_t*)out->idx[1], (size_t)nb);                                    /* _internal_flatten_bf :2796 */
        if (!(args[2]->flags & FLAG_CONSTANT) || args[2]->data.type.base_type != TYPE_FLOAT) goto dynamic;
        out->cutoff_max = *(const float*)args[2]->data.ptr;
        int path_length = 4;
        if (nargs == 4) { if (!(args[3]->flags & FLAG_CONSTANT) || args[3]->data.type.base_type != TYPE_INT) goto dynamic; path_length = *(const int32_t*)args[3]->data.ptr; }
        {
            const md_bitfield_t* bf_a = (const md_bitfield_t*)args[0]->data.ptr;
            uint32_t* off = (uint32_t*)md_alloc(alloc, sizeof(uint32_t) * (na_sets + 1));
            uint32_t* eoff = (uint32_t*)md_alloc(alloc, sizeof(uint32_t) * (na_sets + 1));
            off[0] = 0; for (size_t i = 0; i < na_sets; ++i) off[i + 1] = off[i] + (uint32_t)md_bitfield_popcount(&bf_a[i]);
            /* the exclusion masks are static (topology only): built here with the reference's own function, once (md_script_functions.inl:2838-2840) */
            md_bitfield_t bf_b = md_bitfield_create(alloc); md_bitfield_t excl = md_bitfield_create(alloc);
            for (size_t j = 0; j < out->idx_count[1]; ++j) md_bitfield_set_bit(&bf_b, (uint64_t)out->idx[1][j]);
            size_t total = 0; int32_t* eidx = NULL; size_t ecap = 0;
            eoff[0] = 0;
            for (size_t i = 0; i < na_sets; ++i) {
                md_bitfield_clear(&excl); md_bitfield_and(&excl, &bf_a[i], &bf_b);
                md_util_mask_grow_by_bonds(&excl, mol, (size_t)path_length, NULL);
                const size_t c = md_bitfield_popcount(&excl);
                if (total + c > ecap) { const size_t ncap = (total + c) * 2 + 16; int32_t* ne = (int32_t*)md_alloc(alloc, sizeof(int32_t) * ncap); if (total) MEMCPY(ne, eidx, sizeof(int32_t) * total); eidx = ne; ecap = ncap; }
                if (c) md_bitfield_iter_extract_indices(eidx + total, c, md_bitfield_iter_create(&excl));
                total += c; eoff[i + 1] = (uint32_t)total;
            }
            out->num_structures = na_sets; out->structure_offsets = off;
            out->idx[2] = eidx; out->idx_count[2] = total; out->structure_offsets_b = eoff; out->num_structures_b = na_sets;
        }
        return true;
    }
    if (str_eq(pname, STR_LIT("rmsd")) && nargs == 1) {   /* _rmsd :4287: the (flattened) selection against the initial configuration */
        out->op = MDGPU_OP_RMSD;
        if ((n = mdgpu__arg_indices((int32_t**)&out->idx[0], NULL, NULL, args[0], alloc)) < 0) goto dynamic;
        out->idx_count[0] = mdgpu__flatten((int32_t*)out->idx[0], (size_t)n);
        return true;
    }
    if (str_eq(pname, STR_LIT("distance_pair")) && nargs == 2) {   /* _distance_pair :3972; an array of bitfields = one centre of mass per bitfield (coordinate_extract :1503 -> extract_com :857) */
        out->op = MDGPU_OP_DISTANCE_PAIR;
        for (size_t k = 0; k < 2; ++k) {
            size_t ns = 0;
            if ((n = mdgpu__arg_indices((int32_t**)&out->idx[k], &ns, NULL, args[k], alloc)) < 0) goto dynamic; out->idx_count[k] = (size_t)n;
            if (args[k]->data.type.base_type == TYPE_BITFIELD && ns > 1) {
                const md_bitfield_t* bf = (const md_bitfield_t*)args[k]->data.ptr;
                uint32_t* off = (uint32_t*)md_alloc(alloc, sizeof(uint32_t) * (ns + 1));
                off[0] = 0; for (size_t i = 0; i < ns; ++i) off[i + 1] = off[i] + (uint32_t)md_bitfield_popcount(&bf[i]);
                if (k == 0) { out->structure_offsets = off; out->num_structures = ns; } else { out->structure_offsets_b = off; out->num_structures_b = ns; }
            }
        }
        return true;
    }
    if ((str_eq(pname, STR_LIT("distance_min")) || str_eq(pname, STR_LIT("distance_max"))) && nargs == 2) {
        out->op = str_eq(pname, STR_LIT("distance_min")) ? MDGPU_OP_DISTANCE_MIN : MDGPU_OP_DISTANCE_MAX;
        for (size_t k = 0; k < 2; ++k) {
            size_t ns = 0;
            if ((n = mdgpu__lower_sel_arg(out, (int)k, args[k], &ns, alloc)) < 0) goto dynamic;
            if (!(out->dyn[k].radius_max > 0.0f) && args[k]->data.type.base_type == TYPE_BITFIELD && ns > 1) {   /* one centre of mass per selection (coordinate_extract :1503), as for distance_pair */
                uint32_t* off = mdgpu__arg_part_offsets(args[k], alloc);
                if (k == 0) { out->structure_offsets = off; out->num_structures = ns; } else { out->structure_offsets_b = off; out->num_structures_b = ns; }
            }
        }
        return true;
    }
    {
        const bool dist = str_eq(pname, STR_LIT("distance")), ang = str_eq(pname, STR_LIT("angle")), dih = str_eq(pname, STR_LIT("dihedral"));
        const size_t need = dist ? 2 : (ang ? 3 : (dih ? 4 : 0));
        if (need && nargs == need && n_ctx) {   /* integer arguments remapped into every context: first atom of the context + i - 1 (remap_index_to_context :1023) */
            out->op = dist ? MDGPU_OP_DISTANCE : (ang ? MDGPU_OP_ANGLE : MDGPU_OP_DIHEDRAL);
            out->num_structures = n_ctx;
            for (size_t k = 0; k < need; ++k) {
                const ast_node_t* an = args[k];
                if (an->type == AST_PROC_CALL && an->proc && str_eq(an->proc->name, STR_LIT("com")) && md_array_size(an->children) == 1) an = an->children[0];   /* com(x) contributes x's position (:4726) */
                if (an->data.type.base_type == TYPE_BITFIELD && !(an->flags & FLAG_DYNAMIC)) {
                    /* a selection inside the contexts: in context c its position is the centre of mass of (selection AND context) (coordinate_extract_com
                     * with ctx->mol_ctx, md_script_functions.inl:1812-1823). A constant node carries the whole-system selection; a context-relative one
                     * (atom(2:3) in residue(:)) is evaluated per context by the reference's own evaluator, as evaluate_context does (md_script.c:3463-3494). */
                    uint32_t* off = (uint32_t*)md_alloc(alloc, sizeof(uint32_t) * (n_ctx + 1)); off[0] = 0;
                    md_array(int32_t) all = 0;
                    md_bitfield_t tmp = {0}; md_bitfield_init(&tmp, alloc);
                    for (size_t c = 0; c < n_ctx; ++c) {
                        const md_bitfield_t* src = NULL; size_t nsrc = 0;
                        data_t dd = {0};
                        if ((an->flags & FLAG_CONSTANT) && an->data.ptr) { src = (const md_bitfield_t*)an->data.ptr; nsrc = element_count(an->data); }
                        else {
                            eval_context_t ectx = { .ir = (md_script_ir_t*)mdgpu__lower_ir, .mol = mdgpu__lower_mol, .temp_alloc = alloc, .alloc = alloc };
                            ectx.mol_ctx = &ctx_bf[c];
                            dd.type = an->data.type; allocate_data(&dd, dd.type, alloc);
                            if (!evaluate_node(&dd, an, &ectx)) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': argument %zu could not be evaluated in context %zu", STR_ARG(ident), k, c); return false; }
                            src = (const md_bitfield_t*)dd.ptr; nsrc = element_count(dd);
                        }
                        if (nsrc != 1) { MD_LOG_ERROR("mdgpu: property '" STR_FMT "': an array of selections inside a context expression is not lowered", STR_ARG(ident)); return false; }
                        md_bitfield_and(&tmp, src, &ctx_bf[c]);
                        const size_t cnt = md_bitfield_popcount(&tmp);
                        const size_t at = md_array_size(all);
                        md_array_resize(all, at + cnt, alloc);
                        if (cnt) md_bitfield_iter_extract_indices(all + at, cnt, md_bitfield_iter_create(&tmp));
                        off[c + 1] = off[c] + (uint32_t)cnt;
                    }
                    out->idx[k] = all; out->idx_count[k] = md_array_size(all); out->arg_offsets[k] = off; out->arg_parts[k] = (uint32_t)n_ctx; out->com_args |= 1u << k;
                    continue;
                }
                if (!(args[k]->flags & FLAG_CONSTANT) || args[k]->data.type.base_type != TYPE_INT || element_count(args[k]->data) != 1) {
                    MD_LOG_ERROR("mdgpu: property '" STR_FMT "': `in` is lowered for integer and selection arguments", STR_ARG(ident)); return false;
                }
                const int32_t v = *(const int32_t*)args[k]->data.ptr;
                int32_t* idx = (int32_t*)md_alloc(alloc, sizeof(int32_t) * n_ctx);
                for (size_t c = 0; c < n_ctx; ++c) {   /* remap_index_to_context rejects indices outside the context (md_script_functions.inl:1023-1040) */
                    if (v < 1 || (int64_t)ctx_bf[c].beg_bit + v - 1 >= (int64_t)ctx_bf[c].end_bit) {
                        MD_LOG_ERROR("mdgpu: property '" STR_FMT "': supplied index (%d) is not within the range of context %zu", STR_ARG(ident), v, c); return false;
                    }
                    idx[c] = (int32_t)ctx_bf[c].beg_bit + v - 1;
                }
                out->idx[k] = idx; out->idx_count[k] = n_ctx;
            }
            return true;
        }
        if (need && nargs == need) {
            out->op = dist ? MDGPU_OP_DISTANCE : (ang ? MDGPU_OP_ANGLE : MDGPU_OP_DIHEDRAL);
            for (size_t k = 0; k < need; ++k) {
                size_t ns = 0;
                const ast_node_t* a = args[k];
                /* com(x) as an argument is the position coordinate_extract_com yields for x (_com :4726), which is what the argument x itself
                 * contributes (:1717): distance(com(sel), 5) == distance(sel, 5) */
                if (a->type == AST_PROC_CALL && a->proc && str_eq(a->proc->name, STR_LIT("com")) && md_array_size(a->children) == 1) a = a->children[0];
                if ((n = mdgpu__lower_sel_arg(out, (int)k, a, &ns, alloc)) < 0) goto dynamic;
                if (out->dyn[k].radius_max > 0.0f) { out->com_args |= 1u << k; continue; }   /* the frame's dynamic selection: its centre of mass */
                if (a->data.type.base_type == TYPE_BITFIELD) {
                    /* an array of selections: the centre of the selections' centres (coordinate_extract_com :1826-1842). distance() is FLAG_FLATTEN
                     * (md_script_functions.inl:680), so the front-end has already merged its array arguments into one bitfield: ns == 1 there. */
                    if (ns > 1) { out->arg_offsets[k] = mdgpu__arg_part_offsets(a, alloc); out->arg_parts[k] = (uint32_t)ns; }
                    out->com_args |= 1u << k;   /* a selection goes through md_util_com_compute even with one atom (coordinate_extract_com :1812) */
                }
            }
            return true;
        }
    }
    MD_LOG_ERROR("mdgpu: procedure '" STR_FMT "' of property '" STR_FMT "' is outside the GPU hot-path scope", STR_ARG(pname), STR_ARG(ident));
    return false;
dynamic:
    MD_LOG_ERROR("mdgpu: property '" STR_FMT "' has a dynamic (per-frame) argument; only static selections are lowered", STR_ARG(ident));
    return false;
}

/* Lower every property of a compiled script. */
static bool md_script_gpu_lower_sys(md_script_gpu_lowered_t* out, const md_script_ir_t* ir, const md_system_t* mol, md_allocator_i* alloc) {
    mdgpu__lower_ir = ir; mdgpu__lower_mol = mol;
    const size_t np = md_array_size(ir->property_names);
    out->num_props = np;
    out->props = (mdgpu_property_desc_t*)md_alloc(alloc, sizeof(mdgpu_property_desc_t) * (np ? np : 1));
    out->names = (char(*)[64])md_alloc(alloc, 64 * (np ? np : 1));
    for (size_t i = 0; i < np; ++i) {
        if (!mdgpu__lower_property(&out->props[i], ir->property_names[i], ir->property_nodes[i], mol, alloc)) return false;
        const str_t nm = ir->property_names[i];
        memset(out->names[i], 0, 64); memcpy(out->names[i], nm.ptr, nm.len < 63 ? nm.len : 63);
        out->props[i].name = out->names[i];
    }
    return true;
}

static bool md_script_gpu_lower(md_script_gpu_lowered_t* out, const md_script_ir_t* ir, md_allocator_i* alloc) { return md_script_gpu_lower_sys(out, ir, NULL, alloc); }

/* md_script_eval_create counterpart for the device side: the plan holds what md_script_eval_t holds on the host. */
static mdgpu_plan* md_script_gpu_plan_create(const md_script_ir_t* ir, const md_system_t* mol, size_t num_frames, int device, md_allocator_i* alloc) {
    md_script_gpu_lowered_t low = {0};
    if (!md_script_gpu_lower_sys(&low, ir, mol, alloc)) return NULL;
    float* mass = (float*)md_alloc(alloc, sizeof(float) * mol->atom.count);
    md_atom_extract_masses(mass, 0, mol->atom.count, &mol->atom);                       /* as eval_properties does, md_script.c:5764 */
    mdgpu_system_desc_t sys = {0};
    sys.num_atoms = mol->atom.count; sys.atom_mass = mass;
    sys.bond_conn_offset = mol->bond.conn.offset; sys.bond_conn_atom_idx = mol->bond.conn.atom_idx; sys.bond_conn_offset_count = mol->bond.conn.offset_count;
    mdgpu_plan_options_t opt = {0}; opt.device = device;
    mdgpu_plan* plan = mdgpu_plan_create(&sys, low.props, low.num_props, num_frames, &opt);
    if (!plan) MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error());
    return plan;
}

/* md_script_eval_frame_range with the frame loop on the GPU: same arguments plus the plan; fills eval->property_data exactly where the
 * CPU path does (values, weights, min/max, ranges), sets the completed-frame bits, stamps a fresh fingerprint (md_script.c:6604-6609). */
static bool md_script_gpu_eval_frame_range(mdgpu_plan* plan, md_script_eval_t* eval, const md_script_ir_t* ir, const md_trajectory_i* traj,
                                           uint32_t frame_beg, uint32_t frame_end, uint32_t loader_threads) {
    (void)ir;
    if (mdgpu_eval_trajectory(plan, (const mdgpu_trajectory_i*)traj, frame_beg, frame_end, loader_threads) != 0 || mdgpu_plan_sync(plan) != 0) {
        MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error());
        return false;
    }
    const size_t np = md_array_size(eval->property_data);
    for (size_t i = 0; i < np; ++i) {
        mdgpu_property_data_t d;
        if (mdgpu_plan_property_data(plan, i, &d) != 0) { MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error()); return false; }
        md_script_property_data_t* p = &eval->property_data[i];
        if (d.num_values != p->num_values) { MD_LOG_ERROR("mdgpu: property %zu layout mismatch", i); return false; }
        MEMCPY(p->values, d.values, sizeof(float) * d.num_values);
        p->min_value = d.min_value; p->max_value = d.max_value;
        p->min_range[0] = d.min_range[0]; p->max_range[0] = d.max_range[0];
        if (p->aggregate && p->aggregate->num_values == eval->frame_count &&   /* several values per frame: per-frame mean / variance / extent (md_script.c:5886-5890) */
            mdgpu_plan_property_aggregate(plan, i, p->aggregate->population_mean, p->aggregate->population_var, (float*)p->aggregate->population_ext, eval->frame_count) != 0) {
            MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error()); return false;
        }
    }
    const size_t nwords = (eval->frame_count + 63) / 64;
    uint64_t* words = (uint64_t*)md_alloc(md_get_heap_allocator(), sizeof(uint64_t) * (nwords ? nwords : 1));
    if (mdgpu_plan_frame_mask(plan, words, nwords) == 0) {
        md_mutex_lock(&eval->frame_lock);
        for (size_t f = 0; f < eval->frame_count; ++f) if (words[f >> 6] >> (f & 63) & 1ull) md_bitfield_set_bit(&eval->frame_mask, f);
        md_mutex_unlock(&eval->frame_lock);
    }
    md_free(md_get_heap_allocator(), words, sizeof(uint64_t) * (nwords ? nwords : 1));
    const uint64_t fingerprint = generate_fingerprint();
    for (size_t i = 0; i < np; ++i) eval->property_data[i].fingerprint = fingerprint;
    return true;
}


/* =============================================================================================================================
 * The dispatcher: md_script_eval_frame_range itself runs on libmdgpu (zero edits in md_script.c and in VIAMD).
 * Active when md_script_mdgpu_pre.h was included before md_script.c (see md_script_mdgpu.c).
 *
 * One plan per md_script_eval_t, created lazily by the first md_script_eval_frame_range call on that eval (the eval does not know the system
 * or the trajectory before, md_script.c:6506), kept in a small table keyed by the eval pointer, destroyed by md_script_eval_free.
 *  - The plan writes straight into eval->property_data[i].values (mdgpu_plan_bind_property_storage): VIAMD keeps reading the arrays it
 *    always read (src/main.cpp:1286-1303, 1524; density_volume.cpp:149-151).
 *  - Re-entrant like the reference: N enkiTS threads call with disjoint ranges on one eval (src/task_system.cpp:73-87); each call creates its
 *    own readers (md_script.c:5754), takes stream slots as they free up, and returns when its frames are evaluated.
 *  - Every completed batch sets its bits in eval->frame_mask under eval->frame_lock (md_script.c:5962-5964) and refreshes min/max/ranges,
 *    so the UI's concurrent reads (src/main.cpp:1513-1524) see progress; distributions / volumes are refreshed at most every 100 ms.
 *  - md_script_eval_interrupt (:6663) also interrupts the plan (polled between batches); md_script_eval_clear_data (:6563) clears it.
 *  - A script with a statement outside the lowered set is an error (MD_LOG_ERROR, false): the library has no CPU fallback. A host that wants
 *    such scripts evaluated by the reference's own CPU code sets MDGPU_ALLOW_CPU_SCRIPTS=1; MDGPU_DISABLE=1 routes everything there.
 *  - MDGPU_DEVICES="0,1,2,3" makes the plan span several GPUs of the box from this one process (frame blocks per device, one NCCL reduce).
 * ============================================================================================================================= */
#ifdef MD_SCRIPT_MDGPU_DROPIN
#undef md_script_eval_frame_range
#undef md_script_eval_clear_data
#undef md_script_eval_free
#undef md_script_eval_interrupt

typedef struct mdgpu__entry_t {
    md_script_eval_t* eval;
    mdgpu_plan* plan;
    int state;              /* 0 = free slot, 1 = GPU plan, 2 = this eval's script is evaluated by the reference's CPU code */
} mdgpu__entry_t;

#define MDGPU__MAX_EVALS 64
static mdgpu__entry_t mdgpu__entries[MDGPU__MAX_EVALS];
static md_mutex_t mdgpu__table_lock;
static volatile int mdgpu__table_state = 0;   /* 0 = untouched, 1 = being initialised, 2 = ready */

static void mdgpu__table_init(void) {
    if (__sync_bool_compare_and_swap(&mdgpu__table_state, 0, 1)) { md_mutex_init(&mdgpu__table_lock); __sync_synchronize(); mdgpu__table_state = 2; }
    while (mdgpu__table_state != 2) { /* another thread is initialising the lock */ }
}

static mdgpu__entry_t* mdgpu__find(const md_script_eval_t* eval) {   /* table lock held */
    for (int i = 0; i < MDGPU__MAX_EVALS; ++i) if (mdgpu__entries[i].state && mdgpu__entries[i].eval == eval) return &mdgpu__entries[i];
    return NULL;
}

static void mdgpu__copy_scalars(md_script_eval_t* eval, mdgpu_plan* plan) {
    const size_t np = md_array_size(eval->property_data);
    for (size_t i = 0; i < np; ++i) {
        mdgpu_property_data_t d;
        if (mdgpu_plan_property_peek(plan, i, &d) != 0) continue;
        md_script_property_data_t* p = &eval->property_data[i];
        p->min_value = d.min_value; p->max_value = d.max_value;
        p->min_range[0] = d.min_range[0]; p->max_range[0] = d.max_range[0];
    }
}

/* mdgpu_progress_fn: a batch has completed and its results are in eval->property_data */
static void mdgpu__on_batch(void* user, uint32_t frame_beg, uint32_t frame_count) {
    mdgpu__entry_t* e = (mdgpu__entry_t*)user;
    md_script_eval_t* eval = e->eval;
    mdgpu__copy_scalars(eval, e->plan);
    md_mutex_lock(&eval->frame_lock);
    for (uint32_t f = frame_beg; f < frame_beg + frame_count && f < eval->frame_count; ++f) md_bitfield_set_bit(&eval->frame_mask, f);
    md_mutex_unlock(&eval->frame_lock);
}

static size_t mdgpu__device_list(int32_t* out, size_t cap) {
    const char* env = getenv("MDGPU_DEVICES"); size_t n = 0;
    if (!env) return 0;
    while (*env && n < cap) { char* end = NULL; const long v = strtol(env, &end, 10); if (end == env) break; out[n++] = (int32_t)v; env = (*end == ',') ? end + 1 : end; }
    return n;
}

/* the plan of this eval, created on first use; NULL with *cpu = true when the reference's own code evaluates this eval */
static mdgpu_plan* mdgpu__plan_for(md_script_eval_t* eval, const md_script_ir_t* ir, const md_system_t* mol, bool* cpu, bool* failed) {
    *cpu = false; *failed = false;
    mdgpu__table_init();
    md_mutex_lock(&mdgpu__table_lock);
    mdgpu__entry_t* e = mdgpu__find(eval);
    if (!e) {
        for (int i = 0; i < MDGPU__MAX_EVALS && !e; ++i) if (!mdgpu__entries[i].state) e = &mdgpu__entries[i];
        if (!e) { md_mutex_unlock(&mdgpu__table_lock); MD_LOG_ERROR("mdgpu: more than %d live md_script_eval_t objects", MDGPU__MAX_EVALS); *failed = true; return NULL; }
        e->eval = eval; e->plan = NULL; e->state = 2;
        const char* off = getenv("MDGPU_DISABLE");
        if (!(off && off[0] == '1')) {
            md_allocator_i* tmp = md_arena_allocator_create(md_get_heap_allocator(), MEGABYTES(1));
            md_script_gpu_lowered_t low = {0};
            mdgpu_plan* plan = NULL;
            if (md_script_gpu_lower_sys(&low, ir, mol, tmp)) {
                float* mass = (float*)md_alloc(tmp, sizeof(float) * (mol->atom.count ? mol->atom.count : 1));
                md_atom_extract_masses(mass, 0, mol->atom.count, &mol->atom);                       /* as eval_properties does, md_script.c:5764 */
                mdgpu_system_desc_t sd = {0};
                sd.num_atoms = mol->atom.count; sd.atom_mass = mass;
                sd.bond_conn_offset = mol->bond.conn.offset; sd.bond_conn_atom_idx = mol->bond.conn.atom_idx; sd.bond_conn_offset_count = mol->bond.conn.offset_count;
                mdgpu_plan_options_t opt = {0};
                opt.num_streams = 6;   /* concurrent callers each hold a slot while they decode their frames */
                const size_t nd = mdgpu__device_list(opt.devices, 16);
                if (nd > 1) opt.num_devices = (uint32_t)nd; else if (nd == 1) opt.device = opt.devices[0];
                plan = mdgpu_plan_create(&sd, low.props, low.num_props, eval->frame_count, &opt);
                if (!plan) MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error());
                const size_t np = md_array_size(eval->property_data);
                for (size_t i = 0; plan && i < np; ++i) {   /* results go where VIAMD reads them */
                    md_script_property_data_t* p = &eval->property_data[i];
                    const bool agg = p->aggregate && p->aggregate->num_values == eval->frame_count;
                    if (mdgpu_plan_bind_property_storage(plan, i, p->values, p->num_values, agg ? p->aggregate->population_mean : NULL,
                                                         agg ? p->aggregate->population_var : NULL, agg ? (float*)p->aggregate->population_ext : NULL) != 0) {
                        MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error()); mdgpu_plan_destroy(plan); plan = NULL;
                    }
                }
            }
            md_arena_allocator_destroy(tmp);
            if (plan) { e->plan = plan; e->state = 1; mdgpu_plan_set_progress_callback(plan, mdgpu__on_batch, e); }
            else {
                const char* allow = getenv("MDGPU_ALLOW_CPU_SCRIPTS");
                if (!(allow && allow[0] == '1')) { e->state = 0; md_mutex_unlock(&mdgpu__table_lock); *failed = true; return NULL; }   /* no CPU fallback */
                MD_LOG_INFO("mdgpu: this script is evaluated by the reference's CPU code (MDGPU_ALLOW_CPU_SCRIPTS=1)");
            }
        }
    }
    mdgpu_plan* plan = e->plan; *cpu = (e->state == 2);
    md_mutex_unlock(&mdgpu__table_lock);
    return plan;
}

bool md_script_eval_frame_range(md_script_eval_t* eval, const struct md_script_ir_t* ir, const struct md_system_t* mol, const struct md_trajectory_i* traj, uint32_t frame_beg, uint32_t frame_end) {
    ASSERT(eval);
    /* argument checks and messages of the reference (md_script.c:6576-6602) */
    if (!ir)   { MD_LOG_ERROR("Script eval: Immediate representation was null"); return false; }
    if (!mol)  { MD_LOG_ERROR("Script eval: Molecule was null"); return false; }
    if (!traj) { MD_LOG_ERROR("Script eval: Trajectory was null"); return false; }
    const uint32_t num_frames = (uint32_t)md_trajectory_num_frames(traj);
    if (num_frames == 0) { MD_LOG_ERROR("Script eval: Trajectory was empty"); return false; }
    if (frame_beg > frame_end || frame_end > num_frames) { MD_LOG_ERROR("Script eval: Invalid frame range"); return false; }
    if (md_array_size(eval->property_data) == 0) { MD_LOG_INFO("Script eval: No properties present, nothing to evaluate"); return false; }

    bool cpu = false, failed = false;
    mdgpu_plan* plan = mdgpu__plan_for(eval, ir, mol, &cpu, &failed);
    if (failed) return false;
    if (cpu) return md_script_eval_frame_range__cpu(eval, ir, mol, traj, frame_beg, frame_end);

    bool result = true;
    if (!eval->interrupt && frame_end > frame_beg) {
        const uint32_t span = frame_end - frame_beg;
        const uint32_t readers = span >= 256 ? 8 : (span >= 64 ? 2 : 1);   /* a call over a short range is one of many concurrent ones: it is its own loader */
        int rc = mdgpu_eval_trajectory(plan, (const mdgpu_trajectory_i*)traj, frame_beg, frame_end, readers);
        if (rc == 0) rc = mdgpu_plan_sync(plan);                            /* this call's frames are evaluated and folded when it returns */
        if (rc != 0 && rc != MDGPU_ERR_INTERRUPTED) { MD_LOG_ERROR("mdgpu: %s", mdgpu_last_error()); result = false; }
        mdgpu__copy_scalars(eval, plan);
        /* the frame-mask bits were set batch by batch (mdgpu__on_batch) */
    }
    const uint64_t fingerprint = generate_fingerprint();                    /* md_script.c:6604-6609 */
    for (size_t i = 0; i < md_array_size(eval->property_data); ++i) eval->property_data[i].fingerprint = fingerprint;
    return result;
}

void md_script_eval_clear_data(md_script_eval_t* eval) {
    ASSERT(eval);
    mdgpu__table_init();
    md_mutex_lock(&mdgpu__table_lock);
    mdgpu__entry_t* e = mdgpu__find(eval); mdgpu_plan* plan = e ? e->plan : NULL;
    md_mutex_unlock(&mdgpu__table_lock);
    if (plan) mdgpu_plan_clear(plan);          /* zeroes the device accumulators, the bound arrays and the plan's interrupt flag */
    md_script_eval_clear_data__cpu(eval);      /* the reference's own clear has the last word on the host-side state */
}

void md_script_eval_interrupt(md_script_eval_t* eval) {
    md_script_eval_interrupt__cpu(eval);
    mdgpu__table_init();
    md_mutex_lock(&mdgpu__table_lock);
    mdgpu__entry_t* e = mdgpu__find(eval); mdgpu_plan* plan = e ? e->plan : NULL;
    md_mutex_unlock(&mdgpu__table_lock);
    if (plan) mdgpu_plan_interrupt(plan);
}

void md_script_eval_free(md_script_eval_t* eval) {
    mdgpu__table_init();
    md_mutex_lock(&mdgpu__table_lock);
    mdgpu__entry_t* e = mdgpu__find(eval); mdgpu_plan* plan = NULL;
    if (e) { plan = e->plan; e->plan = NULL; e->eval = NULL; e->state = 0; }
    md_mutex_unlock(&mdgpu__table_lock);
    if (plan) mdgpu_plan_destroy(plan);        /* before the arena that holds the bound arrays goes away */
    md_script_eval_free__cpu(eval);
}
#endif /* MD_SCRIPT_MDGPU_DROPIN */
