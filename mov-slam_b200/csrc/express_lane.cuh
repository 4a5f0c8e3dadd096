// express_lane.cuh — EXPRESS (include/EXPRESS.h:79-192) evaluated at THREAD level: one lane per block (births) or per 8-row half
// of a block (candidate descriptors), on a window of the grey plane staged in shared memory.
//
// The warp-level forms in extract.cu spend one lane per pixel (ballots) or one lane per block row (SWAR over 16 pixels); every
// block then costs a full set of warp-wide shuffles, ballots and bookkeeping, and propagation is bound by instruction issue.
// Here a lane owns a whole block half: four pixels per 32-bit operation (VABSDIFF4 against the replicated centre, one carry-less
// per-byte compare), no cross-lane traffic at all, and 32 independent blocks (or 16 block pairs) per warp step.
//
// The functions are __host__ __device__ so that tests/test_express_lane.py can run them on the CPU against the oracle's
// restatement of EXPRESS.h (bit-exact on random and wrap-around inputs) before anything runs on a GPU.
//
// Window layout (written by cp.async, 8 bytes per lane): 16 rows of ROW_WORDS = 6 words (24 bytes), row r = image row y0 + r,
// bytes [xw, xw + 24) with xw = (first needed column) & ~7. The first needed column is mx + 1 for a candidate descriptor (the
// p++-before-read of EXPRESS.h:98-109) and mx for compute_express (which also walks the true block mask); `sb` = that column -
// xw (0..7). Windows of consecutive slots are WIN_STRIDE words apart: 2 * 49, so that the 64-bit row loads of 16 lanes (eight
// slots, two halves 48 words apart) fall on 16 different bank pairs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XL_HD __host__ __device__ __forceinline__
#else
#define XL_HD static inline
#endif

namespace xl {

constexpr int ROW_WORDS = 6;
constexpr int ROW_BYTES = ROW_WORDS * 4;
constexpr int WIN_WORDS = 16 * ROW_WORDS;
constexpr int WIN_STRIDE = WIN_WORDS + 2;
constexpr int MAX_THR = 127;  // the band arithmetic below assumes threshold < 128 (any threshold: the warp-level kernels)

XL_HD uint32_t absdiff4(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vabsdiffu4(a, b);  // VABSDIFF4.U8: one instruction on sm_100a
#else
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        const int x = (a >> (8 * k)) & 0xff, y = (b >> (8 * k)) & 0xff;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * k);
    }
    return r;
#endif
}
XL_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {  // s in {0, 8, 16, 24}
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
XL_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
XL_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
#endif
}

// EXPRESS.h:91-94: low = uint8(centre - thr), high = uint8(centre + thr); a pixel is out of band when low > p || high < p.
// For thr < 128: if neither limit wraps the test is |p - centre| > thr; if one does, low > high and EVERY pixel is out of band.
struct Band {
    uint32_t c4;  // centre replicated
    uint32_t k4;  // per byte 0x7f - thr; 0x80 when every pixel is out of band
};
XL_HD Band band_of(int centre, int thr) {
    Band b;
    b.c4 = (uint32_t)centre * 0x01010101u;
    const bool all = centre < thr || centre + thr > 255;
    b.k4 = all ? 0x80808080u : (uint32_t)(0x7f - thr) * 0x01010101u;
    return b;
}
// bit 7 of byte k: pixel k of v is out of band
XL_HD uint32_t flags4(uint32_t v, Band b) {
    const uint32_t a = absdiff4(v, b.c4);
    const uint32_t s = (a & 0x7f7f7f7fu) + b.k4;  // bit 7: (a & 0x7f) > thr, no carry between bytes (0x7f + 0x7f < 0x100)
    return (a | s) & 0x80808080u;                 // a > thr  (a >= 128 > thr, or the low seven bits decide)
}
// the four flags (bits 7, 15, 23, 31) as a nibble, pixel k -> bit k: 7 + 8k + (21 - 7k) = 28 + k, no two partial products meet
XL_HD uint32_t nib(uint32_t r) { return (r * 0x00204081u) >> 28; }

// Out-of-band flags of 16 (17 when `extra`) consecutive pixels of a window row, from row byte sb (0..7) on. e: the row's words.
XL_HD uint32_t row_flags(const uint32_t (&e)[ROW_WORDS], int sb, Band b, bool extra) {
    const bool o = (sb & 4) != 0;
    const uint32_t s = (uint32_t)(sb & 3) * 8u;
    const uint32_t v0 = o ? e[1] : e[0], v1 = o ? e[2] : e[1], v2 = o ? e[3] : e[2], v3 = o ? e[4] : e[3], v4 = o ? e[5] : e[4];
    const uint32_t p0 = funnel_r(v0, v1, s), p1 = funnel_r(v1, v2, s), p2 = funnel_r(v2, v3, s), p3 = funnel_r(v3, v4, s);
    uint32_t m = nib(flags4(p0, b)) | (nib(flags4(p1, b)) << 4) | (nib(flags4(p2, b)) << 8) | (nib(flags4(p3, b)) << 12);
    if (extra) m |= ((flags4(v4 >> s, b) >> 7) & 1u) << 16;
    return m;
}

XL_HD void load_row(const uint32_t *win, int r, uint32_t (&e)[ROW_WORDS]) {
#if defined(__CUDA_ARCH__)
    const uint2 *p = reinterpret_cast<const uint2 *>(win + r * ROW_WORDS);  // windows are 8-byte aligned: three 64-bit loads
    const uint2 a = p[0], b = p[1], c = p[2];
    e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y; e[4] = c.x; e[5] = c.y;
#else
    for (int k = 0; k < ROW_WORDS; k++) e[k] = win[r * ROW_WORDS + k];
#endif
}

// compute_center (EXPRESS.h:79-88): at(row = cols/2, col = rows/2) and its three upper-left neighbours. sb0: row byte of the
// block's column 0 (may be -1 for candidate windows, whose first byte is column 1; the centre columns are >= 3).
XL_HD int centre_of(const uint32_t *win, int sb0, int rows, int cols) {
    const uint8_t *w = reinterpret_cast<const uint8_t *>(win);
    const int cr = rows >> 1, cc = cols >> 1;
    const uint8_t *p = w + cc * ROW_BYTES + sb0 + cr;
    return ((int)p[0] + (int)p[-ROW_BYTES - 1] + (int)p[-1] + (int)p[-ROW_BYTES]) / 4;
}

// ---- candidate descriptor (EXPRESS.h:90-110), one lane per 8-row half ---------------------------------------------------------
// half h of a rows x cols block (rows, cols in {8, 16}; h = 0 for 8-row blocks): the four descriptor words this half owns, in the
// reference's layout (bit y*rows + x, OR-ed): rows == 16 -> words 4h .. 4h+3; rows == 8 -> words 0 .. 2 (d[3] = 0).
// sb = row byte of column 1 (0..7).
XL_HD void half_descriptor(const uint32_t *win, int sb, int rows, int cols, int half, Band b, uint32_t (&d)[4]) {
    const uint32_t cm = cols == 16 ? 0xffffu : 0xffu;
    uint32_t h[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 8; r++) {
        uint32_t e[ROW_WORDS];
        load_row(win, 8 * half + r, e);
        h[r] = row_flags(e, sb, b, false) & cm;
    }
    if (rows == 16) {
        d[0] = h[0] | (h[1] << 16);
        d[1] = h[2] | (h[3] << 16);
        d[2] = h[4] | (h[5] << 16);
        d[3] = h[6] | (h[7] << 16);
    } else {  // bit = 8y + x: 16-column rows overlap their successor by eight bits
        d[0] = h[0] | (h[1] << 8) | (h[2] << 16) | (h[3] << 24);
        d[1] = (h[3] >> 8) | h[4] | (h[5] << 8) | (h[6] << 16) | (h[7] << 24);
        d[2] = h[7] >> 8;
        d[3] = 0;
    }
}

// ---- compute_express + descriptor (EXPRESS.h:117-192, :90-110), one lane per block ---------------------------------------------
XL_HD void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &s, uint32_t &cy) {
    s = a ^ b ^ c;
    cy = (a & b) | (c & (a ^ b));
}
XL_HD bool has_run(uint32_t bits, int rounds) {  // a run of `rounds` (4, 6 or 8) consecutive set bits
    const uint32_t x2 = bits & (bits >> 1), x4 = x2 & (x2 >> 2);
    const uint32_t x = rounds == 4 ? x4 : rounds == 6 ? (x4 & (x2 >> 4)) : (x4 & (x4 >> 4));
    return x != 0;
}
// bit d of plane i: bit i of ceil(len(d) / 2), len(d) = min(d + 1, rows, cols, rows + cols - 1 - d) (the tables EXPRESS.h:20-38)
constexpr uint32_t half_len_plane(int rows, int cols, int i) {
    uint32_t m = 0;
    for (int d = 0; d < rows + cols - 1; d++) {
        int len = d + 1;
        if (rows < len) len = rows;
        if (cols < len) len = cols;
        if (rows + cols - 1 - d < len) len = rows + cols - 1 - d;
        m |= (uint32_t)((((len + 1) / 2) >> i) & 1) << d;
    }
    return m;
}
template <int ROWS, int COLS, int I>
struct HalfLenPlane {
    static constexpr uint32_t v = half_len_plane(ROWS, COLS, I);  // evaluated at compile time: a literal in device code
};

// sb = row byte of column 0 (0..7). Returns compute_express; when it passes, desc = the block's descriptor.
XL_HD bool block_express(const uint32_t *win, int sb, int rows, int cols, int thr, uint32_t (&desc)[8]) {
    const Band b = band_of(centre_of(win, sb, rows, cols), thr);
    const uint32_t cm = cols == 16 ? 0xffffu : 0xffu;
    uint32_t m0[16], m1[16];  // true block mask (diagonal walk) and the p++-shifted mask (pre-check, descriptor), per row
    int f = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 16; r++) {
        m0[r] = 0;
        m1[r] = 0;
        if (r < 8 || rows == 16) {
            uint32_t e[ROW_WORDS];
            load_row(win, r, e);
            const uint32_t q = row_flags(e, sb, b, true);
            m0[r] = q & cm;
            m1[r] = (q >> 1) & cm;
            f += popc32(m1[r]);
        }
    }
    // pre-check (:122-139): the running count only grows and is tested at row ends, so "reaches precheck at some row end" ==
    // "total >= precheck" (the uint8 counter cannot wrap before the break: it stops at the first row end >= precheck <= 32)
    if (f < rows * cols / 8) return false;
    // diagonal walk (:141-190): diagonal d of direction 1 is the set of cells with c - r + rows - 1 == d; shifting row r left by
    // rows - 1 - r lines every diagonal up in one bit column, and the win count of all diagonals is a bit-sliced sum of the rows.
    const int slices = rows + cols - 1;
    const int rounds = slices == 31 ? 8 : slices == 23 ? 6 : 4;  // round(slices * .25)
    const uint32_t valid = 0xffffffffu >> (32 - slices);
    uint32_t t0, t1, t2, t3, t4;  // planes of ceil(len / 2)
    if (rows == 16 && cols == 16) {
        t0 = HalfLenPlane<16, 16, 0>::v; t1 = HalfLenPlane<16, 16, 1>::v; t2 = HalfLenPlane<16, 16, 2>::v; t3 = HalfLenPlane<16, 16, 3>::v; t4 = HalfLenPlane<16, 16, 4>::v;
    } else if (rows == 8 && cols == 8) {
        t0 = HalfLenPlane<8, 8, 0>::v; t1 = HalfLenPlane<8, 8, 1>::v; t2 = HalfLenPlane<8, 8, 2>::v; t3 = HalfLenPlane<8, 8, 3>::v; t4 = HalfLenPlane<8, 8, 4>::v;
    } else if (rows == 16) {
        t0 = HalfLenPlane<16, 8, 0>::v; t1 = HalfLenPlane<16, 8, 1>::v; t2 = HalfLenPlane<16, 8, 2>::v; t3 = HalfLenPlane<16, 8, 3>::v; t4 = HalfLenPlane<16, 8, 4>::v;
    } else {
        t0 = HalfLenPlane<8, 16, 0>::v; t1 = HalfLenPlane<8, 16, 1>::v; t2 = HalfLenPlane<8, 16, 2>::v; t3 = HalfLenPlane<8, 16, 3>::v; t4 = HalfLenPlane<8, 16, 4>::v;
    }
    bool ok = false;
    for (int a = 0; a < 2 && !ok; a++) {
        uint32_t x[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 16; r++) {
            // a == 0: diagonal(img, i, true), stepping down-right; a == 1: the same walk on the column-reversed rows
            const uint32_t bits = a == 0 ? m0[r] : (brev32(m0[r]) >> (32 - cols));
            x[r] = (r < 8 || rows == 16) ? bits << (rows - 1 - r) : 0u;
        }
        // carry-save sum of the 16 rows: win count of diagonal d = bit d of the planes w0 (ones) .. w4 (sixteens)
        uint32_t s0, c0, s1, c1, s2, c2, s3, c3, s4, c4, s5, c5, s6, c6;
        full_add(x[0], x[1], x[2], s0, c0);
        full_add(x[3], x[4], x[5], s1, c1);
        full_add(x[6], x[7], x[8], s2, c2);
        full_add(x[9], x[10], x[11], s3, c3);
        full_add(x[12], x[13], x[14], s4, c4);
        full_add(s0, s1, s2, s5, c5);
        full_add(s3, s4, x[15], s6, c6);
        const uint32_t w0 = s5 ^ s6, c7 = s5 & s6;
        uint32_t u0, d0, u1, d1, w1, d3;
        full_add(c0, c1, c2, u0, d0);
        full_add(c3, c4, c5, u1, d1);
        const uint32_t u2 = c6 ^ c7, d2 = c6 & c7;
        full_add(u0, u1, u2, w1, d3);
        uint32_t v0, e0;
        full_add(d0, d1, d2, v0, e0);
        const uint32_t w2 = v0 ^ d3, e1 = v0 & d3;
        const uint32_t w3 = e0 ^ e1, w4 = e0 & e1;
        // win >= loss (:171)  <=>  win >= ceil(len / 2); bit-sliced compare, least significant plane first
        uint32_t ge = 0xffffffffu;
        ge = (w0 & ~t0) | (~(w0 ^ t0) & ge);
        ge = (w1 & ~t1) | (~(w1 ^ t1) & ge);
        ge = (w2 & ~t2) | (~(w2 ^ t2) & ge);
        ge = (w3 & ~t3) | (~(w3 ^ t3) & ge);
        ge = (w4 & ~t4) | (~(w4 ^ t4) & ge);
        // sticky run counters (:169-184): wins reaches `rounds` iff `rounds` consecutive win diagonals exist, the same for losses;
        // the early break (:185) only fires when the verdict is already false
        if (has_run(ge & valid, rounds) && has_run(~ge & valid, rounds)) ok = true;
    }
    if (!ok) return false;
    if (rows == 16) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 8; k++) desc[k] = m1[2 * k] | (m1[2 * k + 1] << 16);
    } else {
        desc[0] = m1[0] | (m1[1] << 8) | (m1[2] << 16) | (m1[3] << 24);
        desc[1] = (m1[3] >> 8) | m1[4] | (m1[5] << 8) | (m1[6] << 16) | (m1[7] << 24);
        desc[2] = m1[7] >> 8;
        desc[3] = desc[4] = desc[5] = desc[6] = desc[7] = 0;
    }
    return true;
}

}  // namespace xl
