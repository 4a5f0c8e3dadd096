// grid.cu — the per-pixel slot grid (VideoImage::mvi, CV_32SC4) from a frame's ordered hop list.
// Replaces the pixel loop of VideoDecoder::NextImage (src/VideoDecoder.cc:330-345): for every pixel, slots 0..2
// are the first three hops (in push_back order) whose source rectangle covers it and slot 3 is the LAST covering
// hop when four or more cover it; untouched slots are -1.
//
// This is the HBM-bound kernel of the front-end: 16 bytes are written per pixel, nothing is read back.
// Design:
//  * per-block ownership, no atomics: a CTA owns an 8-row band of one frame, a warp owns 32x8-pixel tiles of it,
//    a lane owns one pixel column; every pixel is written exactly once with one 128-bit streaming store, so a warp
//    store covers 512 contiguous bytes and the -1 fill is implicit.
//  * the band's hops are gathered once, in list order, into shared memory (ordered ballot compaction, 32-hop chunks
//    skipped by their y-extent); chunks of the staged list carry an x-extent so a tile only scans chunks near it.
//  * a tile's candidates are processed 31 at a time, candidate i of a chunk sitting in lane 30-i: one 32x32 bit
//    transpose gives every lane the candidates covering its column, one ballot per row gives the candidates covering
//    that row, and the slots fall out of bit scans — highest bit = first hop, lowest bit = last hop. Lane 31 holds
//    the index -1, so "no such hop" shuffles -1 out without a select.
//  * rows are independent, so the running state is four slot registers and a count, not a tile of registers.
#include <cstdio>

#include "common.cuh"

namespace {

constexpr int GRID_MAX_WARPS = 16;
constexpr int GRID_LIST_CAP = 2048;    // band-list entries staged in shared memory (2 words each = 16 KB); multiple of 32
constexpr int GRID_CHUNK_CAP = 1024;   // surviving 32-hop chunk ids per band
constexpr int TILE_Q = 124;            // a tile's candidate queue: 4 chunks of 31

__device__ __forceinline__ int bfind(unsigned x) {  // position of the highest set bit, -1 when x == 0
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ unsigned shl_clamp(unsigned v, int s) {  // shift amounts >= 32 (incl. -1) give 0
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

// 32x32 bit-matrix transpose across a warp: lane r ends with bit c == (lane c's input bit r).
__device__ __forceinline__ unsigned transpose32(unsigned x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const unsigned m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
    }
    return x;
}

__device__ __forceinline__ unsigned col_mask(unsigned wx, int tx) {
    const int x0 = (int)(wx & 0xffffu), x1 = (int)(wx >> 16);
    const int lo = max(x0 - tx, 0), hi = min(x1 - tx, 31);
    return ((2u << (hi - lo)) - 1u) << lo;  // hi-lo in [0,31]; 2u<<31 wraps to 0 -> all ones
}

struct Slots {
    int s0, s1, s2, s3;
    int cnt;
};

// Fold one chunk's candidates covering this pixel (mask m over lanes, bit 30-i = candidate i of the chunk, `idx` =
// this lane's candidate index or -1) into the slots. FIRST: cnt == 0 is known (first chunk of the tile).
template <bool FIRST, bool NEED_CNT>
__device__ __forceinline__ void fold(Slots &st, unsigned m, int idx) {
    const int p0 = bfind(m);
    const unsigned m1 = m ^ shl_clamp(1u, p0);
    const int p1 = bfind(m1);
    const unsigned m2 = m1 ^ shl_clamp(1u, p1);
    const int p2 = bfind(m2);
    const unsigned m3 = m2 ^ shl_clamp(1u, p2);
    const int v0 = __shfl_sync(0xffffffffu, idx, p0);  // p == -1 reads lane 31 == -1
    const int v1 = __shfl_sync(0xffffffffu, idx, p1);
    const int v2 = __shfl_sync(0xffffffffu, idx, p2);
    if (FIRST) {
        st.s0 = v0;
        st.s1 = v1;
        st.s2 = v2;
        st.s3 = __shfl_sync(0xffffffffu, idx, __ffs(m3) - 1);  // lowest remaining bit = last hop in order
        if (NEED_CNT) st.cnt = min(3, __popc(m));
    } else {
        const int c = st.cnt;
        const unsigned rem = c == 0 ? m3 : c == 1 ? m2 : c == 2 ? m1 : m;
        const int vl = __shfl_sync(0xffffffffu, idx, __ffs(rem) - 1);
        if (c == 0) {
            st.s0 = v0;
            st.s1 = v1;
            st.s2 = v2;
        } else if (c == 1) {
            st.s1 = v0;
            st.s2 = v1;
        } else if (c == 2) {
            st.s2 = v0;
        }
        if (rem) st.s3 = vl;
        st.cnt = min(3, c + __popc(m));
    }
}

// Exclusive prefix of per-warp counts (wcnt[0..nwarps)) for this warp, and the total. nwarps <= 16.
__device__ __forceinline__ void warp_counts_prefix(const int *wcnt, int nwarps, int warp, int lane, int &before, int &total) {
    int v = lane < nwarps ? wcnt[lane] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    total = __shfl_sync(0xffffffffu, incl, 15);
    before = __shfl_sync(0xffffffffu, incl - v, warp);
}

constexpr int SB_ROWS = 32;  // rows a CTA owns: four 8-row bands share one staged hop list

// list_i word: hop index (22 bits) | first row (5 bits) << 22 | last row (5 bits) << 27, rows relative to the CTA's band
__device__ __forceinline__ unsigned row_mask8(uint32_t wi, int sub) {
    const int r0 = (int)((wi >> 22) & 31u) - 8 * sub, r1 = (int)(wi >> 27) - 8 * sub;
    const int a = max(r0, 0), b = min(r1, 7);
    return b >= a ? ((2u << (b - a)) - 1u) << a : 0u;
}

__global__ void __launch_bounds__(GRID_MAX_WARPS * 32, 2)
grid_kernel(WinParams p, int NSB, int NT, const HopRect *__restrict__ hop_rects, const int32_t *__restrict__ nhops,
            const int32_t *__restrict__ chunk_bbox, int4 *__restrict__ grid) {
    extern __shared__ uint32_t smem[];
    uint32_t *list_x = smem;                                   // [CAP] x0 | x1<<16
    uint32_t *list_i = smem + GRID_LIST_CAP;                   // [CAP] hop index | r0<<22 | r1<<27
    int32_t  *clist = (int32_t *)(smem + 2 * GRID_LIST_CAP);   // [CHUNK_CAP] surviving chunk ids; reused as x-extents
    __shared__ int32_t wcnt[GRID_MAX_WARPS];
    __shared__ uint32_t cext[GRID_LIST_CAP / 32];              // per list chunk: rows touched, bit r = some entry covers row r
    __shared__ uint32_t cand[GRID_MAX_WARPS][2][TILE_Q + 36];  // per-warp candidate queue (x word, i word)

    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int band = blockIdx.x % NSB;
    const int sg = blockIdx.x / NSB;  // s*n_out + g
    const int s = sg / p.n_out, g = sg - s * p.n_out;
    const int ylo = band * SB_ROWS, yhi = min(ylo + SB_ROWS - 1, p.H - 1);
    const int n_h = nhops[s * p.n_in + g];
    const HopRect *rects = hop_rects + (size_t)sg * p.max_hops;
    const int32_t *bbox = chunk_bbox + (size_t)sg * p.max_chunks;
    const int nchunks = (n_h + 31) >> 5;

    // ---- phase 1a: ordered list of the 32-hop chunks whose y-extent touches the band --------------------------
    int n_cl = 0;  // identical in every thread
    for (int base = 0; base < nchunks; base += blockDim.x) {
        const int c = base + threadIdx.x;
        bool pred = false;
        if (c < nchunks) {
            const int bb = __ldg(&bbox[c]);
            const int ymin = (int16_t)(bb & 0xffff), ymax = bb >> 16;
            pred = ymax >= ylo && ymin <= yhi;
        }
        const unsigned b = __ballot_sync(0xffffffffu, pred);
        if (lane == 0) wcnt[warp] = __popc(b);
        __syncthreads();
        int before, tot;
        warp_counts_prefix(wcnt, nwarps, warp, lane, before, tot);
        const int pos = n_cl + before + __popc(b & lt);
        if (pred && pos < GRID_CHUNK_CAP) clist[pos] = c;
        n_cl += tot;
        __syncthreads();
    }
    const bool chunk_overflow = n_cl > GRID_CHUNK_CAP;

    // ---- phase 1b: ordered list of the hops touching the band, staged in shared memory -----------------------
    int n_list = 0;
    if (!chunk_overflow) {
        for (int base = 0; base < n_cl; base += nwarps) {
            const int ci = base + warp;
            bool pred = false;
            HopRect r = {0, 32767, -1, -32768};
            int h = 0;
            if (ci < n_cl) {
                h = clist[ci] * 32 + lane;
                if (h < n_h) {
                    r = rects[h];
                    pred = r.y1 >= ylo && r.y0 <= yhi;
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, pred);
            if (lane == 0) wcnt[warp] = __popc(b);
            __syncthreads();
            int before, tot;
            warp_counts_prefix(wcnt, nwarps, warp, lane, before, tot);
            const int pos = n_list + before + __popc(b & lt);
            if (pred && pos < GRID_LIST_CAP) {
                const int r0 = max((int)r.y0, ylo) - ylo, r1 = min((int)r.y1, yhi) - ylo;
                list_x[pos] = (uint32_t)(uint16_t)r.x0 | ((uint32_t)(uint16_t)r.x1 << 16);
                list_i[pos] = (uint32_t)h | ((uint32_t)r0 << 22) | ((uint32_t)r1 << 27);
            }
            n_list += tot;
            __syncthreads();
        }
    }
    const bool direct = chunk_overflow || n_list > GRID_LIST_CAP;  // pathological input: tiles scan global memory

    // ---- phase 1c: x-extent and touched rows of every 32-entry chunk of the staged list (reuses clist) --------
    const int n_lc = direct ? 0 : (n_list + 31) >> 5;
    for (int c = warp; c < n_lc; c += nwarps) {
        const int e = c * 32 + lane;
        int xmin = 65535, xmax = -1;
        unsigned rows = 0;
        if (e < n_list) {
            const uint32_t wx = list_x[e], wi = list_i[e];
            xmin = (int)(wx & 0xffffu);
            xmax = (int)(wx >> 16);
            const int r0 = (wi >> 22) & 31u, r1 = wi >> 27;
            rows = ((2u << (r1 - r0)) - 1u) << r0;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            rows |= __shfl_xor_sync(0xffffffffu, rows, o);
        }
        if (lane == 0) {
            clist[c] = xmin | (xmax << 16);
            cext[c] = rows;
        }
    }
    __syncthreads();

    // ---- phase 2: a warp takes (8-row band, 64-px strip) items; no barrier from here on ------------------------
    uint32_t *qx = cand[warp][0], *qi = cand[warp][1];
    const int n_strips = (NT + 1) >> 1;  // a warp gathers once for a 64-px strip = two adjacent tiles
    const int n_sub = (yhi - ylo + 8) >> 3;
    for (int item = warp; item < n_sub * n_strips; item += nwarps) {
        const int sub = item / n_strips, strip = item - sub * n_strips;
        const int sx = strip * 64;
        const int by = ylo + 8 * sub;  // first row of this 8-row band
        const unsigned submask = 0xffu << (8 * sub);
        // gather the strip's candidates (ascending hop order) into the queue
        int nq = 0;
        bool overflow = direct;
        if (!direct) {
            // list chunks that can matter to this item (x-extent meets the strip, some entry covers one of its rows): 32
            // chunks are tested per ballot, only the survivors are visited, in ascending order
            for (int cb = 0; cb < n_lc && !overflow; cb += 32) {
                const int cc = cb + lane;
                bool rel = false;
                if (cc < n_lc) {
                    const int ext = clist[cc];
                    rel = (ext >> 16) >= sx && (ext & 0xffff) <= sx + 63 && (cext[cc] & submask);
                }
                unsigned todo = __ballot_sync(0xffffffffu, rel);
                while (todo) {
                    const int c = cb + __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int e = c * 32 + lane;
                    bool pred = false;
                    uint32_t wx = 0, wi = 0;
                    if (e < n_list) {
                        wx = list_x[e];
                        wi = list_i[e];
                        pred = (int)(wx >> 16) >= sx && (int)(wx & 0xffffu) <= sx + 63 && row_mask8(wi, sub) != 0;
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, pred);
                    if (nq + __popc(b) > TILE_Q) {
                        overflow = true;
                        break;
                    }
                    if (pred) {
                        const int pos = nq + __popc(b & lt);
                        qx[pos] = wx;
                        qi[pos] = wi;
                    }
                    nq += __popc(b);
                }
            }
        }
        __syncwarp();
        const int nchk = (nq + 30) / 31;
        // per-chunk lane data shared by both tiles of the strip: candidate i of chunk c sits in lane 30-i
        uint32_t cwx[4];
        int idx[4];
        unsigned rm[4];
        if (!overflow) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                cwx[c] = 0x0000ffffu;  // x0 = 65535 > x1 = 0: covers no column
                idx[c] = -1;
                rm[c] = 0;
                const int e = c * 31 + (30 - lane);
                if (c < nchk && lane < 31 && e < nq) {
                    const uint32_t wi = qi[e];
                    cwx[c] = qx[e];
                    idx[c] = (int)(wi & 0x3fffffu);
                    rm[c] = row_mask8(wi, sub);
                }
            }
        }
        for (int tsub = 0; tsub < 2; tsub++) {
        const int tx = sx + 32 * tsub;
        if (tx >= p.W) break;
        const int x = tx + lane;
        int4 *out = grid + ((size_t)sg * p.H + by) * p.W + x;
        if (!overflow) {
            // fast path: <= 4 chunks of 31 candidates, column masks in registers. Rows are folded independently, and a
            // row whose covering set equals the previous row's (block edges are sparse) reuses its slots.
            const bool full = tx + 31 < p.W && by + 7 < p.H;  // warp-uniform: no per-store bounds test needed
            const size_t rstride = (size_t)p.W;
            if (nchk <= 1) {
                const unsigned cm = (lane < 31 && (int)(cwx[0] >> 16) >= tx && (int)(cwx[0] & 0xffffu) <= tx + 31) ? col_mask(cwx[0], tx) : 0u;
                const unsigned col0 = transpose32(cm, lane);
                Slots st = {-1, -1, -1, -1, 0};
                unsigned prev0 = 0;
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    const unsigned rowm0 = __ballot_sync(0xffffffffu, (rm[0] >> y) & 1u);
                    if (y == 0 || rowm0 != prev0) {  // warp-uniform
                        fold<true, false>(st, col0 & rowm0, idx[0]);
                        prev0 = rowm0;
                    }
                    if (full || (x < p.W && by + y < p.H)) st_cs_v4(out, make_int4(st.s0, st.s1, st.s2, st.s3));
                    out += rstride;
                }
            } else {
                unsigned col[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    col[c] = 0;
                    if (c < nchk) {  // warp-uniform
                        const unsigned cm = (lane < 31 && (int)(cwx[c] >> 16) >= tx && (int)(cwx[c] & 0xffffu) <= tx + 31) ? col_mask(cwx[c], tx) : 0u;
                        col[c] = transpose32(cm, lane);
                    }
                }
                Slots st = {-1, -1, -1, -1, 0};
                unsigned prev[4] = {0, 0, 0, 0};
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    unsigned rowm[4];
                    bool same = y != 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        rowm[c] = c < nchk ? __ballot_sync(0xffffffffu, (rm[c] >> y) & 1u) : 0u;
                        same = same && rowm[c] == prev[c];
                    }
                    if (!same) {  // warp-uniform
                        fold<true, true>(st, col[0] & rowm[0], idx[0]);
#pragma unroll
                        for (int c = 1; c < 4; c++)
                            if (c < nchk) fold<false, true>(st, col[c] & rowm[c], idx[c]);
#pragma unroll
                        for (int c = 0; c < 4; c++) prev[c] = rowm[c];
                    }
                    if (full || (x < p.W && by + y < p.H)) st_cs_v4(out, make_int4(st.s0, st.s1, st.s2, st.s3));
                    out += rstride;
                }
            }
        } else {
            // slow path (more than 124 candidates in one tile, or the band list did not fit): one row at a time,
            // streaming every source entry again and folding 31 candidates per step.
            const int n_src = direct ? n_h : n_list;
            for (int y = 0; y < 8; y++) {
                if (by + y >= p.H) break;
                Slots st = {-1, -1, -1, -1, 0};
                int nq = 0;
                for (int base = 0; base <= n_src; base += 32) {  // one extra, empty pass flushes the queue
                    const int e = base + lane;
                    bool pred = false;
                    uint32_t wx = 0, wi = 0;
                    if (e < n_src) {
                        if (!direct) {
                            wx = list_x[e];
                            wi = list_i[e];
                            pred = (int)(wx >> 16) >= tx && (int)(wx & 0xffffu) <= tx + 31 && ((row_mask8(wi, sub) >> y) & 1u);
                            wi &= 0x3fffffu;
                        } else {
                            const HopRect r = rects[e];
                            pred = r.y1 >= by + y && r.y0 <= by + y && r.x1 >= tx && r.x0 <= tx + 31;
                            wx = (uint32_t)(uint16_t)r.x0 | ((uint32_t)(uint16_t)r.x1 << 16);
                            wi = (uint32_t)e;
                        }
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, pred);
                    if (pred) {
                        const int pos = nq + __popc(b & lt);
                        qx[pos] = wx;
                        qi[pos] = wi;
                    }
                    nq += __popc(b);
                    __syncwarp();
                    const bool last = base + 32 > n_src;
                    while (nq >= 31 || (last && nq > 0)) {
                        const int take = min(nq, 31);
                        const int e2 = 30 - lane;
                        unsigned cm = 0;
                        int id = -1;
                        if (lane < 31 && e2 < take) {
                            cm = col_mask(qx[e2], tx);
                            id = (int)qi[e2];
                        }
                        const unsigned colm = transpose32(cm, lane);
                        fold<false, true>(st, colm, id);
                        __syncwarp();
                        // move the remainder to the front (reads 31.., writes 0..: disjoint for nq-take <= 32)
                        uint32_t a = 0, c2 = 0;
                        const bool mv = lane < nq - take;
                        if (mv) {
                            a = qx[take + lane];
                            c2 = qi[take + lane];
                        }
                        __syncwarp();
                        if (mv) {
                            qx[lane] = a;
                            qi[lane] = c2;
                        }
                        nq -= take;
                        __syncwarp();
                    }
                }
                if (x < p.W) st_cs_v4(out + (size_t)y * p.W, make_int4(st.s0, st.s1, st.s2, st.s3));
            }
        }
        __syncwarp();
        }  // tiles of the strip
    }
}

}  // namespace

int movfe_grid_launch(movfe_ctx *ctx, const WinParams &p) {
    ProfScope prof(ctx, MOVFE_STAGE_GRID);
    prof.launches(1);
    // a CTA owns a 32-row band = 4 x n_strips (8-row band, 64-px strip) items; warps per CTA: the largest divisor of the
    // item count that is <= 16 keeps every warp equally loaded
    const int n_strips = (ctx->NT + 1) / 2;
    const int nsb = (p.H + SB_ROWS - 1) / SB_ROWS;
    int nw = 8;
    for (int w = GRID_MAX_WARPS; w >= 4; w--)
        if ((4 * n_strips) % w == 0) {
            nw = w;
            break;
        }
    const size_t smem = (2 * GRID_LIST_CAP + GRID_CHUNK_CAP) * sizeof(uint32_t);
    MOVFE_CUDA(ctx, cudaFuncSetAttribute(grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = p.S * p.n_out * nsb;
    grid_kernel<<<blocks, nw * 32, smem, ctx->stream>>>(p, nsb, ctx->NT, ctx->d_hop_rect, ctx->d_nhops, ctx->d_chunk_bbox,
                                                       ctx->d_grid);
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}
