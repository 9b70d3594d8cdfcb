/* harness_common.h — TEST INFRASTRUCTURE shared by ref_harness.c and shim_harness.c: in-memory md_trajectory_i implementations,
 * system loading through the reference's own loaders, tiny argv helpers. Included after the reference headers. */
#ifndef HARNESS_COMMON_H
#define HARNESS_COMMON_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../viamd_b200/csrc/synth.h"

/* ---------------------------------------------------------------- in-memory trajectories */

typedef enum { TRAJ_RAW, TRAJ_SYNTHWATER, TRAJ_SYNTHMEMBRANE } traj_kind_t;

typedef struct mem_traj_t {
    traj_kind_t kind;
    size_t num_frames, num_atoms;
    double* frame_times;
    /* raw */
    const unsigned char* raw; size_t raw_frame_bytes;
    /* synthwater */
    mdsynth_water_t water; float *bx, *by, *bz;
    /* synthmembrane */
    mdsynth_membrane_t memb; float* mbase; uint32_t* mmol;
    /* frames [cache_beg, cache_end) of a synthetic trajectory generated ahead of time (timing runs: load_frame is then a memcpy, as for a
     * trajectory held in memory) */
    float* cache; size_t cache_beg, cache_end;
} mem_traj_t;

typedef struct raw_frame_hdr_t { double cell[6]; uint32_t flags; uint32_t pad; } raw_frame_hdr_t;

static bool mt_get_header(struct md_trajectory_o* inst, md_trajectory_header_t* h) {
    mem_traj_t* t = (mem_traj_t*)inst;
    memset(h, 0, sizeof(*h));
    h->num_frames = t->num_frames; h->num_atoms = t->num_atoms; h->frame_times = t->frame_times;
    return true;
}

static bool mt_load_frame(struct md_trajectory_reader_o* inst, int64_t idx, md_trajectory_frame_header_t* hdr, float* x, float* y, float* z) {
    mem_traj_t* t = (mem_traj_t*)inst;
    if (idx < 0 || (size_t)idx >= t->num_frames) return false;
    md_unitcell_t cell = {0};
    if (t->cache && x && (size_t)idx >= t->cache_beg && (size_t)idx < t->cache_end && t->kind != TRAJ_RAW) {
        const float* p = t->cache + ((size_t)idx - t->cache_beg) * 3 * t->num_atoms;
        memcpy(x, p, t->num_atoms * 4); memcpy(y, p + t->num_atoms, t->num_atoms * 4); memcpy(z, p + 2 * t->num_atoms, t->num_atoms * 4);
        if (t->kind == TRAJ_SYNTHMEMBRANE) cell = md_unitcell_from_extent((double)t->memb.Lx, (double)t->memb.Ly, (double)t->memb.Lz);
        else cell = md_unitcell_from_extent((double)t->water.L, (double)t->water.L, (double)t->water.L);
    } else
    if (t->kind == TRAJ_RAW) {
        const unsigned char* p = t->raw + (size_t)idx * t->raw_frame_bytes;
        raw_frame_hdr_t fh; memcpy(&fh, p, sizeof(fh)); p += sizeof(fh);
        cell.x = fh.cell[0]; cell.xy = fh.cell[1]; cell.xz = fh.cell[2];
        cell.y = fh.cell[3]; cell.yz = fh.cell[4]; cell.z = fh.cell[5];
        cell.flags = (md_unitcell_flags_t)fh.flags;
        if (x) { memcpy(x, p, t->num_atoms * 4); memcpy(y, p + t->num_atoms * 4, t->num_atoms * 4); memcpy(z, p + t->num_atoms * 8, t->num_atoms * 4); }
    } else if (t->kind == TRAJ_SYNTHMEMBRANE) {
        cell = md_unitcell_from_extent((double)t->memb.Lx, (double)t->memb.Ly, (double)t->memb.Lz);
        if (x) mdsynth_membrane_frame(&t->memb, (uint32_t)idx, t->mbase, t->mmol, x, y, z);
    } else {
        cell = md_unitcell_from_extent((double)t->water.L, (double)t->water.L, (double)t->water.L);
        if (x) mdsynth_water_frame(&t->water, (uint32_t)idx, t->bx, t->by, t->bz, x, y, z);
    }
    if (hdr) {
        hdr->num_atoms = t->num_atoms; hdr->index = idx; hdr->timestamp = (double)idx; hdr->unitcell = cell;
    }
    return true;
}

static void mt_reader_free(struct md_trajectory_reader_i* r) { (void)r; }
static bool mt_init_reader(md_trajectory_reader_i* r, struct md_trajectory_o* inst) {
    r->inst = (struct md_trajectory_reader_o*)inst; r->free = mt_reader_free; r->load_frame = mt_load_frame;
    return true;
}
static void mt_free(struct md_trajectory_i* t) { (void)t; }

static void* read_file(const char* path, size_t* out_size) {
    FILE* f = fopen(path, "rb"); if (!f) return NULL;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    void* p = malloc((size_t)n); if (fread(p, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(p); return NULL; }
    fclose(f); *out_size = (size_t)n; return p;
}

static bool make_traj(md_trajectory_i* out, mem_traj_t* mt, const char* spec, md_system_t* sys) {
    memset(mt, 0, sizeof(*mt));
    if (strncmp(spec, "raw:", 4) == 0) {
        size_t sz = 0; unsigned char* buf = read_file(spec + 4, &sz);
        if (!buf || sz < 24 || memcmp(buf, "MDRAWTRJ", 8) != 0) { fprintf(stderr, "bad raw traj %s\n", spec); return false; }
        uint64_t nf, na; memcpy(&nf, buf + 8, 8); memcpy(&na, buf + 16, 8);
        mt->kind = TRAJ_RAW; mt->num_frames = nf; mt->num_atoms = na; mt->raw = buf + 24;
        mt->raw_frame_bytes = sizeof(raw_frame_hdr_t) + na * 12;
        if (sz < 24 + nf * mt->raw_frame_bytes) { fprintf(stderr, "raw traj truncated\n"); return false; }
    } else if (strncmp(spec, "synthwater:", 11) == 0) {
        unsigned n, seed, nf;
        if (sscanf(spec + 11, "%u:%u:%u", &n, &seed, &nf) != 3) return false;
        mt->kind = TRAJ_SYNTHWATER; mt->water = mdsynth_water_desc(n, seed);
        mt->num_frames = nf; mt->num_atoms = mt->water.num_atoms;
        mt->bx = malloc(mt->num_atoms * 4); mt->by = malloc(mt->num_atoms * 4); mt->bz = malloc(mt->num_atoms * 4);
        mdsynth_water_base(&mt->water, mt->bx, mt->by, mt->bz, NULL, NULL, NULL);
    } else if (strncmp(spec, "synthmembrane:", 14) == 0) {
        unsigned nl, nwxy, nwz, seed, nf;
        if (sscanf(spec + 14, "%u:%u:%u:%u:%u", &nl, &nwxy, &nwz, &seed, &nf) != 5) return false;
        mt->kind = TRAJ_SYNTHMEMBRANE; mt->memb = mdsynth_membrane_desc(nl, nwxy, nwz, seed);
        mt->num_frames = nf; mt->num_atoms = mt->memb.num_atoms;
        mt->mbase = malloc(mt->num_atoms * 12); mt->mmol = malloc(mt->num_atoms * 4);
        mdsynth_membrane_base(&mt->memb, mt->mbase, NULL, mt->mmol);
    } else if (strncmp(spec, "xtc:", 4) == 0) {   /* the reference's own XTC reader (md_xtc.c) */
        md_trajectory_i* t = md_xtc_trajectory_create((str_t){ spec + 4, strlen(spec + 4) }, md_get_heap_allocator(), MD_TRAJECTORY_FLAG_DISABLE_CACHE_WRITE);
        if (!t) { fprintf(stderr, "failed to open xtc %s\n", spec + 4); return false; }
        *out = *t; return true;
    } else if (strcmp(spec, "sys") == 0) {
        if (!sys->trajectory) { fprintf(stderr, "system has no attached trajectory\n"); return false; }
        *out = *sys->trajectory; return true;
    } else { fprintf(stderr, "unknown traj spec %s\n", spec); return false; }
    mt->frame_times = malloc(mt->num_frames * sizeof(double));
    for (size_t i = 0; i < mt->num_frames; ++i) mt->frame_times[i] = (double)i;
    out->inst = (struct md_trajectory_o*)mt; out->free = mt_free; out->get_header = mt_get_header; out->init_reader = mt_init_reader;
    return true;
}

/* generate frames [beg, end) of a synthetic trajectory once; load_frame then copies them (no-op for other kinds) */
static void mt_materialize(mem_traj_t* t, size_t beg, size_t end) {
    if (t->kind == TRAJ_RAW || end <= beg) return;
    const size_t n = t->num_atoms; float* c = malloc((end - beg) * 3 * n * sizeof(float)); if (!c) return;
    for (size_t f = beg; f < end; ++f) {
        float* p = c + (f - beg) * 3 * n;
        if (t->kind == TRAJ_SYNTHMEMBRANE) mdsynth_membrane_frame(&t->memb, (uint32_t)f, t->mbase, t->mmol, p, p + n, p + 2 * n);
        else mdsynth_water_frame(&t->water, (uint32_t)f, t->bx, t->by, t->bz, p, p + n, p + 2 * n);
    }
    t->cache = c; t->cache_beg = beg; t->cache_end = end;
}

/* ---------------------------------------------------------------- helpers */

static bool ends_with(const char* s, const char* suf) { size_t a = strlen(s), b = strlen(suf); return a >= b && strcmp(s + a - b, suf) == 0; }

static bool load_system(md_system_t* sys, const char* path, md_allocator_i* alloc) {
    memset(sys, 0, sizeof(*sys)); sys->alloc = alloc;
    str_t p = { path, strlen(path) };
    bool ok = false;
    if (ends_with(path, ".gro")) ok = md_gro_system_init_from_file(sys, p);
    else if (ends_with(path, ".pdb")) ok = md_pdb_system_init_from_file(sys, p, MD_PDB_OPTION_DISABLE_CACHE_FILE_WRITE);
    if (!ok) { fprintf(stderr, "failed to load system %s\n", path); return false; }
    md_util_system_postprocess(sys, MD_UTIL_POSTPROCESS_ALL);
    return true;
}

static const char* arg_val(int argc, char** argv, const char* key, const char* def) {
    for (int i = 2; i + 1 < argc; ++i) if (strcmp(argv[i], key) == 0) return argv[i + 1];
    return def;
}
static bool parse_range(const char* s, long* b, long* e) { return s && sscanf(s, "%ld:%ld", b, e) == 2; }
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }


#endif
