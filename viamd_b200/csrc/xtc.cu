// xtc.cu — XTC frame decode on the device: the step right before the per-frame hot path. Compressed frames cross PCIe
// (about a third of the raw floats) and are expanded in HBM with the arithmetic of the reference's reader
// (md_xtc_decode_frame_data_soa_scaled md_xtc.c:747-931 as xtc_reader_load_frame :947-993 calls it: scale 10, nm -> Angstrom).
//
// The bit stream of a frame is sequential (field widths adapt as the stream goes), so the work is split in two kernels:
//   k_xtc_scan    one warp per frame (lane 0 walks, all lanes stage the stream through shared memory) walks the stream WITHOUT decoding: per group (one full-width coordinate + its run of small
//                 differences) it only needs the flag/run bits to know where the next group starts. It records (bit offset, first atom,
//                 small-integer index, run length) per group — a short dependent chain per group, all frames of the batch in parallel.
//   k_xtc_decode  one thread per (frame, group) expands its group from that record: unpack the mixed-radix integers, add the bias,
//                 int -> float, scale. Fully parallel, writes the SoA frame layout the property kernels read.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace mdg {

__constant__ uint32_t c_magicints[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096,
    5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287,
    660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216 };
constexpr int XTC_FIRSTIDX = 9, XTC_LASTIDX = 73;

MDG_D uint32_t be32(const uint8_t* p) { return __byte_perm(*(const uint32_t*)p, 0, 0x0123); }   // frames start 4-byte aligned (XDR)

// up to 64 bits starting at bit `pos` of a big-endian bit stream whose first byte is 4-byte aligned (three aligned word loads)
MDG_D unsigned long long peek_bits(const uint8_t* stream, unsigned long long pos, unsigned n) {
    const uint32_t* w = (const uint32_t*)stream + (pos >> 5);
    const unsigned s = (unsigned)(pos & 31);
    const unsigned long long hi = ((unsigned long long)__byte_perm(w[0], 0, 0x0123) << 32) | __byte_perm(w[1], 0, 0x0123);
    const uint32_t lo = __byte_perm(w[2], 0, 0x0123);
    const unsigned long long x = s ? ((hi << s) | (lo >> (32 - s))) : hi;
    return x >> (64 - n);
}

// A packed field is transmitted as whole bytes, least significant first, then the remaining high bits (xdrfile.c sendints; md_xtc.c:304-309
// undoes it with a byte swap). n <= 64.
MDG_D unsigned long long field_value(unsigned long long w, unsigned n) {
    const unsigned k = n >> 3, part = n & 7;
    const unsigned long long bytes_be = part ? (w >> part) : w;                                   // k bytes, first transmitted byte on top
    unsigned long long v = 0;
    if (k) {
        const uint32_t hi = (uint32_t)(bytes_be >> 32), lo = (uint32_t)bytes_be;
        const unsigned long long rev = ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);   // bswap64
        v = rev >> (64 - 8 * k);
    }
    if (part) v |= (w & ((1ull << part) - 1ull)) << (8 * k);
    return v;
}

// three integers from an n-bit packed field with radices (size_y, size_z) (unpack_coord64 md_xtc.c:304-315, unpack_coord128 :317)
MDG_D void unpack3(const uint8_t* stream, unsigned long long pos, unsigned n, uint32_t size_y, uint32_t size_z, int out[3]) {
    if (n <= 64) {
        const unsigned long long v = field_value(peek_bits(stream, pos, n), n);
        const unsigned long long zy = (unsigned long long)size_z * size_y;
        const uint32_t x = (uint32_t)(v / zy);
        const unsigned long long yz = v / size_z;
        out[0] = (int)x; out[1] = (int)(uint32_t)(yz - (unsigned long long)x * size_y); out[2] = (int)(uint32_t)(v - yz * size_z);
    } else {   // wider than 64 bits (boxes beyond ~2.6 um): byte by byte into 128 bits, following the file format (xdrfile.c receiveints)
        unsigned __int128 v = 0; unsigned shift = 0, left = n; unsigned long long p = pos;
        while (left >= 8) { v |= (unsigned __int128)peek_bits(stream, p, 8) << shift; shift += 8; left -= 8; p += 8; }
        if (left) v |= (unsigned __int128)peek_bits(stream, p, left) << shift;
        const unsigned __int128 zy = (unsigned __int128)size_z * size_y;
        const uint32_t x = (uint32_t)(v / zy);
        const unsigned long long yz = (unsigned long long)(v / size_z);
        out[0] = (int)x; out[1] = (int)(uint32_t)(yz - (unsigned long long)x * size_y); out[2] = (int)(uint32_t)((unsigned long long)v - yz * size_z);
    }
}

MDG_D int sizeofint_dev(uint32_t size) {   // md_xtc.c:157-166
    uint32_t num = 1; int nb = 0;
    while ((int)size >= (int)num && nb < 32) { nb++; num *= 2; }
    return nb;
}
MDG_D int sizeofints_dev(const uint32_t sizes[3]) {   // md_xtc.c:168-193: bit length of the product of the three sizes
    unsigned __int128 t = (unsigned __int128)sizes[0] * sizes[1]; t *= sizes[2];
    int nbytes = 0; while (t > 0xff) { t >>= 8; nbytes++; }
    uint32_t top = (uint32_t)t, num = 1; int nb = 0;
    while (top >= num) { nb++; num *= 2; }
    return nb + nbytes * 8;
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per frame. Walking the stream is a serial dependent chain (where the next group starts depends on this group's flag and run
// bits), ~33 000 links for a 100k-atom water frame. Two things keep it short:
//  * the stream is staged through shared memory (all lanes copy, coalesced, byte-swapped once), so a link costs a ~25-cycle shared load
//    instead of an uncoalesced global one;
//  * SPECULATION: solvent-dominated frames repeat the same group shape (same run length, same small-integer width) thousands of times.
//    Lane j parses the group that would start j predicted lengths ahead; a ballot finds how many leading predictions held, those groups
//    are committed at once (up to 32 per round), and the first lane whose group differs is still a correctly placed group — it is
//    committed as well and updates the state and the prediction. Irregular streams degrade to one group per round, never to a wrong result.
constexpr uint32_t SCAN_CHUNK = 1024;   // bytes staged per round (small: the CTA has to fit beside the shared-memory-heavy pair kernel of another stream)
constexpr uint32_t SCAN_SLACK = 128;

__global__ void __launch_bounds__(32) k_xtc_scan(const uint8_t* __restrict__ blob, const unsigned long long* __restrict__ frame_off, uint32_t num_atoms, int B,
                           XtcFrameInfo* __restrict__ info, uint2* __restrict__ rec, uint16_t* __restrict__ rec_state, size_t rec_stride) {
    const int f = blockIdx.x, lane = threadIdx.x;
    constexpr uint32_t BUF_WORDS = (SCAN_CHUNK + SCAN_SLACK) / 4;
    __shared__ uint32_t s_buf[BUF_WORDS + 2];
    __shared__ XtcFrameInfo s_fi;
    __shared__ int s_state, s_smallidx;   // state: 0 walking, 1 finished ok, 2 error
    __shared__ uint32_t s_walk[8];        // walk state handed back by the serial burst
    const uint8_t* fr = blob + frame_off[f];
    const unsigned long long nbytes = frame_off[f + 1] - frame_off[f];
    uint2* r = rec + (size_t)f * rec_stride; uint16_t* rs = rec_state + (size_t)f * rec_stride;
    if (lane == 0) {
        XtcFrameInfo fi{}; fi.status = 1; int state = 2;
        do {
            if (nbytes < 56 || be32(fr) != 1995u) break;                                      // decode_header md_xtc.c:419-434
            if ((uint32_t)be32(fr + 4) != num_atoms || (uint32_t)be32(fr + 52) != num_atoms) break;   // :776-781
            if (num_atoms <= 9) { if (nbytes < 56 + 12ull * num_atoms) break; fi.status = 0; fi.ngroups = 0; fi.data_off = 56; state = 1; break; }
            if (nbytes < 92) break;
            fi.precision = __uint_as_float(be32(fr + 56));
            for (int k = 0; k < 3; ++k) { fi.minint[k] = (int)be32(fr + 60 + 4 * k); const int mx = (int)be32(fr + 72 + 4 * k); fi.sizeint[k] = (uint32_t)(mx - fi.minint[k] + 1); }
            const int smallidx = (int)be32(fr + 84);
            if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) break;
            const unsigned long long data_bytes = be32(fr + 88);
            if (nbytes < 92 + data_bytes) break;
            fi.data_off = 92;
            if ((fi.sizeint[0] | fi.sizeint[1] | fi.sizeint[2]) > 0xffffffu) { fi.bitsize = 0; for (int k = 0; k < 3; ++k) fi.bitsizeint[k] = (uint32_t)sizeofint_dev(fi.sizeint[k]); }
            else fi.bitsize = (uint32_t)sizeofints_dev(fi.sizeint);
            s_smallidx = smallidx; state = 0;
        } while (false);
        s_fi = fi; s_state = state;
    }
    __syncwarp();
    int state = s_state; uint32_t g = 0;
    if (state == 0) {
        const uint32_t step_bits = s_fi.bitsize ? s_fi.bitsize : s_fi.bitsizeint[0] + s_fi.bitsizeint[1] + s_fi.bitsizeint[2];
        const unsigned long long total_bits = (unsigned long long)be32(fr + 88) * 8ull;
        const uint32_t* stream32 = (const uint32_t*)(fr + 92);
        const uint32_t stream_words = (uint32_t)((total_bits / 8 + 3) / 4) + 4;    // + guard words (the blob is padded by 32 bytes)
        // warp-uniform walk state
        uint32_t bit = 0, atom = 0, lpred = 0; int smallidx = s_smallidx, run = 0, run_count = 0;
        uint32_t base_bit = 0, staged_bits = 0, n_rounds = 0, n_restage = 0;
        while (state == 0) {
            ++n_rounds;
            if (atom >= num_atoms) { state = 1; break; }
            if (bit < base_bit || bit + step_bits + 6 + 64 > base_bit + staged_bits) {   // (re)stage from the word that holds `bit`
                const uint32_t w0 = bit >> 5;
                const uint32_t nw = min(BUF_WORDS + 2, stream_words - min(stream_words, w0));
                __syncwarp();
                for (uint32_t i = lane; i < nw; i += 32) s_buf[i] = __byte_perm(stream32[w0 + i], 0, 0x0123);
                __syncwarp();
                base_bit = w0 << 5; staged_bits = nw * 32u; ++n_restage;
                if (bit + step_bits + 6 + 64 > base_bit + staged_bits) { state = 2; break; }   // stream ends inside a group
            }
            // lane j: the group that starts j predicted lengths ahead, parsed under the current state (md_xtc.c:850-929, positions only)
            const uint32_t per = (run > 0 ? (uint32_t)run_count : 0u) + 1u;
            const uint32_t p = bit + (uint32_t)lane * lpred, a = atom + (uint32_t)lane * per;
            const bool inwin = p + step_bits + 6 + 64 <= base_bit + staged_bits;
            uint32_t data = 0;
            if (inwin) { const uint32_t q = p + step_bits - base_bit; data = __funnelshift_l(s_buf[(q >> 5) + 1], s_buf[q >> 5], q & 31u) >> 26; }
            const uint32_t flag = data & 32u;
            int r_ = run, rc_ = run_count, ism = 0;
            if (flag) { r_ = (int)(data & 31u); rc_ = r_ / 3; ism = r_ % 3; r_ -= ism; ism--; }
            const uint32_t rcu = r_ > 0 ? (uint32_t)rc_ : 0u;
            const uint32_t len = step_bits + (flag ? 6u : 1u) + rcu * (uint32_t)smallidx;
            const bool bad = ((unsigned long long)p + step_bits + 1 > total_bits) || (a + (uint32_t)rc_ + 1u > num_atoms);   // :875 "Buffer overrun during decompression"
            const bool ok = inwin && a < num_atoms && !bad && len == lpred && ism == 0 && r_ == run && rc_ == run_count;
            const uint32_t okm = __ballot_sync(0xffffffffu, ok);
            const uint32_t nok = okm == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~okm) - 1u;   // leading predictions that held
            if ((uint32_t)lane < nok) { r[g + lane] = make_uint2(p, a); rs[g + lane] = (uint16_t)((uint32_t)smallidx | (rcu << 8)); }
            if (nok == 32u) { bit += 32u * lpred; atom += 32u * per; g += 32u; continue; }
            // the first lane whose group differs from the prediction: its start and incoming state are right, so it is a real group
            const uint32_t p_s = __shfl_sync(0xffffffffu, p, nok), a_s = __shfl_sync(0xffffffffu, a, nok), len_s = __shfl_sync(0xffffffffu, len, nok), rcu_s = __shfl_sync(0xffffffffu, rcu, nok);
            const int r_s = __shfl_sync(0xffffffffu, r_, nok), rc_s = __shfl_sync(0xffffffffu, rc_, nok), ism_s = __shfl_sync(0xffffffffu, ism, nok);
            const bool inwin_s = __shfl_sync(0xffffffffu, (int)inwin, nok) != 0, bad_s = __shfl_sync(0xffffffffu, (int)bad, nok) != 0;
            if (a_s >= num_atoms) { g += nok; atom = a_s; state = 1; break; }
            if (!inwin_s) { bit = p_s; atom = a_s; g += nok; continue; }                    // beyond the staged bytes: commit the prefix, restage
            if (bad_s) { state = 2; break; }
            if (lane == 0) { r[g + nok] = make_uint2(p_s, a_s); rs[g + nok] = (uint16_t)((uint32_t)smallidx | (rcu_s << 8)); }
            g += nok + 1u; bit = p_s + len_s; atom = a_s + rcu_s + 1u; run = r_s; run_count = rc_s; smallidx += ism_s; lpred = len_s;
            if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) { state = 2; break; }    // :914
            if (nok < 3u) {
                // The prediction keeps failing (group shapes alternate): a burst of plain serial steps by lane 0 inside the staged window —
                // one shared load + a dozen dependent integer instructions per group — before speculation is tried again.
                if (lane == 0) {
                    uint32_t b_ = bit, a_ = atom, g_ = g, lp = lpred, si = (uint32_t)smallidx, rn = (uint32_t)run, rcn = (uint32_t)run_count, err = 0;
                    const uint32_t win_end = base_bit + staged_bits - (step_bits + 6 + 64);
                    const uint32_t bits_end = (uint32_t)min(total_bits, 0xffffffffull);
                    // one loop-carried branch per group; everything else is selects (a lone warp pays ~20 cycles per branch)
                    for (int it = 0; it < 128 && a_ < num_atoms && b_ <= win_end; ++it) {
                        const uint32_t q = b_ + step_bits - base_bit;
                        const uint32_t d6 = __funnelshift_l(s_buf[(q >> 5) + 1], s_buf[q >> 5], q & 31u) >> 26;
                        const bool fl = (d6 & 32u) != 0u;
                        const uint32_t rnew = d6 & 31u, rcnew = (rnew * 86u) >> 8, ismnew = rnew - 3u * rcnew;   // rnew / 3, rnew % 3 (rnew < 32)
                        rn = fl ? rnew - ismnew : rn; rcn = fl ? rcnew : rcn;
                        const uint32_t ism1 = fl ? ismnew : 1u;                                                  // is_smaller + 1
                        err |= (uint32_t)(a_ + rcn + 1u > num_atoms) | (uint32_t)(b_ + step_bits + 1u > bits_end);
                        const uint32_t rcu2 = rn ? rcn : 0u;
                        r[g_] = make_uint2(b_, a_); rs[g_] = (uint16_t)(si | (rcu2 << 8));
                        lp = step_bits + (fl ? 6u : 1u) + rcu2 * si;
                        b_ += lp; a_ += rcu2 + 1u; ++g_; si += ism1 - 1u;
                        err |= (uint32_t)(si - (uint32_t)XTC_FIRSTIDX >= (uint32_t)(XTC_LASTIDX - XTC_FIRSTIDX));
                        if (err) break;
                    }
                    const int st = err ? 2 : 0;
                    s_walk[0] = b_; s_walk[1] = a_; s_walk[2] = g_; s_walk[3] = lp; s_walk[4] = si; s_walk[5] = rn; s_walk[6] = rcn; s_walk[7] = (uint32_t)st;
                }
                __syncwarp();
                bit = s_walk[0]; atom = s_walk[1]; g = s_walk[2]; lpred = s_walk[3]; smallidx = (int)s_walk[4]; run = (int)s_walk[5]; run_count = (int)s_walk[6];
                if (s_walk[7]) { state = 2; break; }
                __syncwarp();
            }
        }
        if (state == 1 && (unsigned long long)bit > total_bits + 7) state = 2;
        if (lane == 0) { s_fi.rounds = n_rounds; s_fi.restages = n_restage; }
    }
    if (lane == 0) {
        XtcFrameInfo fi = s_fi;
        if (state == 1) { fi.status = 0; if (num_atoms > 9) fi.ngroups = g; } else fi.status = 1;
        info[f] = fi;
    }
}

__global__ void k_xtc_decode(const uint8_t* __restrict__ blob, const unsigned long long* __restrict__ frame_off, uint32_t num_atoms,
                             const XtcFrameInfo* __restrict__ info, const uint2* __restrict__ rec, const uint16_t* __restrict__ rec_state, size_t rec_stride,
                             float* __restrict__ out, size_t frame_stride, size_t axis_stride, int* __restrict__ err) {
    const int f = blockIdx.y;
    const XtcFrameInfo fi = info[f];
    if (fi.status != 0) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(err, MDGPU_ERR_FRAME_SOURCE); return; }
    const uint8_t* fr = blob + frame_off[f];
    float* x = out + (size_t)f * frame_stride; float* y = x + axis_stride; float* z = y + axis_stride;
    const float scale = 10.0f;
    if (num_atoms <= 9) {   // stored as plain floats (:783-792)
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < num_atoms) { const uint8_t* p = fr + 56 + 12 * i; x[i] = __fmul_rn(__uint_as_float(be32(p)), scale); y[i] = __fmul_rn(__uint_as_float(be32(p + 4)), scale); z[i] = __fmul_rn(__uint_as_float(be32(p + 8)), scale); }
        return;
    }
    const float cs = __fdiv_rn(scale, fi.precision);                                      // coord_scale (:841)
    const uint8_t* stream = fr + fi.data_off;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < fi.ngroups; g += gridDim.x * blockDim.x) {
        const uint2 r = rec[(size_t)f * rec_stride + g]; const uint32_t st = rec_state[(size_t)f * rec_stride + g];
        const int smallidx = (int)(st & 0xffu), rc = (int)(st >> 8);
        unsigned long long bit = r.x; uint32_t atom = r.y;
        int c[3];
        if (fi.bitsize == 0) {   // one plain big-endian integer per axis, as the format's writer emits them (xdrfile.c sendbits/receivebits). The reference's
            // reader mis-decodes this branch (unpack_uint32 md_xtc.c:295) and the > 64-bit packed field (:317); DESIGN.md section 3
            for (int k = 0; k < 3; ++k) { c[k] = (int)(uint32_t)peek_bits(stream, bit, fi.bitsizeint[k]); bit += fi.bitsizeint[k]; }
        } else { unpack3(stream, bit, fi.bitsize, fi.sizeint[1], fi.sizeint[2], c); bit += fi.bitsize; }
        c[0] += fi.minint[0]; c[1] += fi.minint[1]; c[2] += fi.minint[2];
        const uint32_t flag = (uint32_t)peek_bits(stream, bit, 1);
        bit += flag ? 6 : 1;
        if (rc > 0) {
            const uint32_t ss = c_magicints[smallidx]; const int smallnum = (int)(ss / 2u);
            const int p0 = c[0], p1 = c[1], p2 = c[2];
            int d[3];
            unpack3(stream, bit, (unsigned)smallidx, ss, ss, d); bit += (unsigned)smallidx;
            c[0] = d[0] + (c[0] - smallnum); c[1] = d[1] + (c[1] - smallnum); c[2] = d[2] + (c[2] - smallnum);
            x[atom] = __fmul_rn((float)c[0], cs); y[atom] = __fmul_rn((float)c[1], cs); z[atom] = __fmul_rn((float)c[2], cs); ++atom;     // the first two are stored swapped (:855-856)
            x[atom] = __fmul_rn((float)p0, cs); y[atom] = __fmul_rn((float)p1, cs); z[atom] = __fmul_rn((float)p2, cs); ++atom;
            for (int i = 1; i < rc; ++i) {
                unpack3(stream, bit, (unsigned)smallidx, ss, ss, d); bit += (unsigned)smallidx;
                c[0] = d[0] + (c[0] - smallnum); c[1] = d[1] + (c[1] - smallnum); c[2] = d[2] + (c[2] - smallnum);
                x[atom] = __fmul_rn((float)c[0], cs); y[atom] = __fmul_rn((float)c[1], cs); z[atom] = __fmul_rn((float)c[2], cs); ++atom;
            }
        } else {
            x[atom] = __fmul_rn((float)c[0], cs); y[atom] = __fmul_rn((float)c[1], cs); z[atom] = __fmul_rn((float)c[2], cs);
        }
    }
}

// scan and expand as separate launches (the plan scans several batches at once on its own stream, then expands per batch)
void launch_xtc_scan(const uint8_t* d_blob, const unsigned long long* d_frame_off, uint32_t num_atoms, int nframes, XtcFrameInfo* d_info,
                     uint2* d_rec, uint16_t* d_rec_state, size_t rec_stride, cudaStream_t s) {
    if (nframes <= 0) return;
    k_xtc_scan<<<nframes, 32, 0, s>>>(d_blob, d_frame_off, num_atoms, nframes, d_info, d_rec, d_rec_state, rec_stride);
    note_launch("k_xtc_scan", s);
}
void launch_xtc_expand(const uint8_t* d_blob, const unsigned long long* d_frame_off, uint32_t num_atoms, int nframes, const XtcFrameInfo* d_info,
                       const uint2* d_rec, const uint16_t* d_rec_state, size_t rec_stride, float* d_out, size_t frame_stride, size_t axis_stride, int* d_err, cudaStream_t s) {
    if (nframes <= 0) return;
    const uint32_t per_frame = num_atoms <= 9 ? 1u : min((num_atoms + 255u) / 256u, 64u);
    dim3 grid(per_frame, nframes);
    k_xtc_decode<<<grid, 256, 0, s>>>(d_blob, d_frame_off, num_atoms, d_info, d_rec, d_rec_state, rec_stride, d_out, frame_stride, axis_stride, d_err);
    note_launch("k_xtc_decode", s);
}

void launch_xtc_decode(const uint8_t* d_blob, const unsigned long long* d_frame_off, uint32_t num_atoms, int B, XtcFrameInfo* d_info,
                       uint2* d_rec, uint16_t* d_rec_state, size_t rec_stride, float* d_out, size_t frame_stride, size_t axis_stride, int* d_err, cudaStream_t s) {
    if (B <= 0) return;
    static const bool timing = getenv("MDGPU_XTC_TIMING") != nullptr;   // diagnostics: per-kernel CUDA-event times on stderr
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventRecord(e0, s); }
    k_xtc_scan<<<B, 32, 0, s>>>(d_blob, d_frame_off, num_atoms, B, d_info, d_rec, d_rec_state, rec_stride);
    note_launch("k_xtc_scan", s);
    if (timing) cudaEventRecord(e1, s);
    const uint32_t per_frame = num_atoms <= 9 ? 1u : min((num_atoms + 255u) / 256u, 64u);
    dim3 grid(per_frame, B);
    k_xtc_decode<<<grid, 256, 0, s>>>(d_blob, d_frame_off, num_atoms, d_info, d_rec, d_rec_state, rec_stride, d_out, frame_stride, axis_stride, d_err);
    note_launch("k_xtc_decode", s);
    if (timing) {
        cudaEventRecord(e2, s); cudaEventSynchronize(e2);
        float a = 0, b = 0; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
        XtcFrameInfo h{}; cudaMemcpy(&h, d_info, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[mdgpu] xtc: %d frames x %u atoms: scan %.3f ms, decode %.3f ms; frame 0: %u groups, %u rounds, %u restages\n", B, num_atoms, a, b, h.ngroups, h.rounds, h.restages);
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    }
}

}  // namespace mdg
