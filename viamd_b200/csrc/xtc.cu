// xtc.cu — XTC frame decode on the device: the step right before the per-frame hot path. Compressed frames cross PCIe
// (about a third of the raw floats) and are expanded in HBM with the arithmetic of the reference's reader
// (md_xtc_decode_frame_data_soa_scaled md_xtc.c:747-931 as xtc_reader_load_frame :947-993 calls it: scale 10, nm -> Angstrom).
//
// The bit stream of a frame is sequential (field widths adapt as the stream goes), so the work is split in two kernels:
//   k_xtc_scan    one thread per frame walks the stream WITHOUT decoding: per group (one full-width coordinate + its run of small
//                 differences) it only needs the flag/run bits to know where the next group starts. It records (bit offset, first atom,
//                 small-integer index, run length) per group — a short dependent chain per group, all frames of the batch in parallel.
//   k_xtc_decode  one thread per (frame, group) expands its group from that record: unpack the mixed-radix integers, add the bias,
//                 int -> float, scale. Fully parallel, writes the SoA frame layout the property kernels read.
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
__global__ void k_xtc_scan(const uint8_t* __restrict__ blob, const unsigned long long* __restrict__ frame_off, uint32_t num_atoms, int B,
                           XtcFrameInfo* __restrict__ info, uint2* __restrict__ rec, uint16_t* __restrict__ rec_state, size_t rec_stride) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    XtcFrameInfo fi{}; fi.status = 1;
    const uint8_t* fr = blob + frame_off[f];
    const unsigned long long nbytes = frame_off[f + 1] - frame_off[f];
    uint2* r = rec + (size_t)f * rec_stride; uint16_t* rs = rec_state + (size_t)f * rec_stride;
    do {
        if (nbytes < 56 || be32(fr) != 1995u) break;                                      // decode_header md_xtc.c:419-434
        if ((uint32_t)be32(fr + 4) != num_atoms || (uint32_t)be32(fr + 52) != num_atoms) break;   // :776-781
        if (num_atoms <= 9) { if (nbytes < 56 + 12ull * num_atoms) break; fi.status = 0; fi.ngroups = 0; fi.data_off = 56; break; }
        if (nbytes < 92) break;
        fi.precision = __uint_as_float(be32(fr + 56));
        for (int k = 0; k < 3; ++k) { fi.minint[k] = (int)be32(fr + 60 + 4 * k); const int mx = (int)be32(fr + 72 + 4 * k); fi.sizeint[k] = (uint32_t)(mx - fi.minint[k] + 1); }
        int smallidx = (int)be32(fr + 84);
        if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) break;
        const unsigned long long data_bytes = be32(fr + 88);
        if (nbytes < 92 + data_bytes) break;
        fi.data_off = 92;
        uint32_t step_bits;
        if ((fi.sizeint[0] | fi.sizeint[1] | fi.sizeint[2]) > 0xffffffu) {
            fi.bitsize = 0; step_bits = 0;
            for (int k = 0; k < 3; ++k) { fi.bitsizeint[k] = (uint32_t)sizeofint_dev(fi.sizeint[k]); step_bits += fi.bitsizeint[k]; }
        } else { fi.bitsize = (uint32_t)sizeofints_dev(fi.sizeint); step_bits = fi.bitsize; }
        const uint8_t* stream = fr + 92;
        const unsigned long long total_bits = data_bytes * 8ull;
        unsigned long long bit = 0; uint32_t atom = 0, g = 0; int run = 0, run_count = 0; bool ok = true;
        while (atom < num_atoms) {   // md_xtc.c:850-929, positions only
            if (bit + step_bits + 1 > total_bits) { ok = false; break; }
            r[g] = make_uint2((uint32_t)bit, atom);
            bit += step_bits;
            const uint32_t data = (uint32_t)peek_bits(stream, bit, 6);
            const uint32_t flag = data & 32u;
            bit += flag ? 6 : 1;
            int is_smaller = 0;
            if (flag) { run = (int)(data & 31u); run_count = run / 3; is_smaller = run % 3; run -= is_smaller; is_smaller--; }
            if (atom + (uint32_t)run_count + 1u > num_atoms) { ok = false; break; }       // "Buffer overrun during decompression" :875
            const int rc = run > 0 ? run_count : 0;
            rs[g] = (uint16_t)((uint32_t)smallidx | ((uint32_t)rc << 8));
            bit += (unsigned long long)rc * (unsigned)smallidx; atom += (uint32_t)rc + 1u;
            smallidx += is_smaller;
            if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) { ok = false; break; }   // :914
            ++g;
        }
        if (!ok || bit > total_bits + 7) break;
        fi.ngroups = g; fi.status = 0;
    } while (false);
    info[f] = fi;
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

void launch_xtc_decode(const uint8_t* d_blob, const unsigned long long* d_frame_off, uint32_t num_atoms, int B, XtcFrameInfo* d_info,
                       uint2* d_rec, uint16_t* d_rec_state, size_t rec_stride, float* d_out, size_t frame_stride, size_t axis_stride, int* d_err, cudaStream_t s) {
    if (B <= 0) return;
    k_xtc_scan<<<(B + 31) / 32, 32, 0, s>>>(d_blob, d_frame_off, num_atoms, B, d_info, d_rec, d_rec_state, rec_stride);
    note_launch("k_xtc_scan", s);
    const uint32_t per_frame = num_atoms <= 9 ? 1u : min((num_atoms + 255u) / 256u, 64u);
    dim3 grid(per_frame, B);
    k_xtc_decode<<<grid, 256, 0, s>>>(d_blob, d_frame_off, num_atoms, d_info, d_rec, d_rec_state, rec_stride, d_out, frame_stride, axis_stride, d_err);
    note_launch("k_xtc_decode", s);
}

}  // namespace mdg
