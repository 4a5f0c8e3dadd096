// grid.cu — the per-pixel slot grid (VideoImage::mvi, CV_32SC4) from a frame's ordered hop list.
// Replaces the pixel loop of VideoDecoder::NextImage (src/VideoDecoder.cc:330-345): for every pixel, slots 0..2
// are the first three hops (in push_back order) whose source rectangle covers it and slot 3 is the LAST covering
// hop when four or more cover it; untouched slots are -1.
//
// This is the HBM-bound kernel of the front-end: 16 bytes are written per pixel, nothing is read back.
// Design:
//  * per-block ownership, no atomics: a CTA owns a 32-row band of one frame, a warp owns 32x32-pixel tiles of it,
//    a lane owns one pixel column; every pixel is written exactly once with one 128-bit streaming store, so a warp
//    store covers 512 contiguous bytes and the -1 fill is implicit.
//  * the band's hops are gathered once, in list order, into shared memory (ordered ballot compaction, 32-hop chunks
//    skipped by their y-extent); chunks of the staged list carry an x-extent so a tile only scans chunks near it.
//  * hops are rectangles, so the set of hops covering pixel (x, y) is (hops covering column x) AND (hops covering row
//    y), and both change only at block edges. A tile's columns fall into a few runs with the same column set, its
//    rows into a few runs with the same row set; the slots are computed ONCE per (column run, row run) cell - a lane
//    per cell - into a small shared-memory table, and a pixel's 16 bytes are one table read away.
//  * candidates sit 31 to a chunk, candidate i in lane 30-i: one 32x32 bit transpose turns per-candidate column (row)
//    masks into per-column (per-row) candidate sets, and the slots fall out of bit scans - highest bit = first hop,
//    lowest bit = last hop. Lane 31 holds the index -1, so "no such hop" shuffles -1 out without a select.
#include <cstdio>

#include "common.cuh"

namespace {

constexpr int GRID_MAX_WARPS = 16;
constexpr int GRID_CHUNK_CAP = 1024;   // surviving 32-hop chunk ids per band
constexpr int TILE_CHUNKS = 6;         // a tile's candidates: up to 6 chunks of 31
constexpr int TILE_Q = 31 * TILE_CHUNKS;
constexpr int TQ_STRIDE = TILE_Q + 6;  // queue words per tile
constexpr int CELL_CAP = 128;          // (column run, row run) cells resolved per batch
constexpr int SB_ROWS = 32;            // rows a CTA owns = rows of a tile
constexpr int MAX_TILES = 32;          // tiles a CTA owns (wider frames are split in x): "lane = tile" bookkeeping

// per-warp scratch in shared memory
struct __align__(16) WarpScratch {
    int4 tab[CELL_CAP];                // slots of the cells of the current batch, index (row run - first run) * ncc + column run
    uint32_t amask[MAX_TILES];         // phase 1: lanes of this warp's chunk whose first tile is t ...
    uint32_t cmask[MAX_TILES];         //          ... and whose second tile is t
    uint8_t repc[32], repr[32];        // phase 2: first column / row of every run
};

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

// queue word i: hop index (22 bits) | first row (5 bits) << 22 | last row (5 bits) << 27, rows relative to the CTA's band
__device__ __forceinline__ unsigned row_mask32(uint32_t wi) {
    const int r0 = (int)((wi >> 22) & 31u), r1 = (int)(wi >> 27);
    return ((2u << (r1 - r0)) - 1u) << r0;  // r1 >= r0; 2u << 31 wraps to 0 -> all ones
}

// 65535 / n + 1 for n = 1..32: (j * magic) >> 16 == j / n for j < 2048
__constant__ uint32_t c_magic[33] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115, 2048};

// Fast path of one 32x32 tile whose candidates (at most NCH chunks of 31) are queued in shared memory.
// FUSED: the cells are the result (common.cuh: TileCells) - run maps to `runs`, slot vectors to `cells`, nothing per pixel.
template <int NCH, bool FUSED>
__device__ __forceinline__ void fast_tile(WarpScratch &ws, const uint32_t *qx, const uint32_t *qi, int nq, int tx, int lane, unsigned lt,
                                          int4 *out, int W, int nrows, bool xin, int32_t *dim, uint8_t *runs, int4 *cells) {
    const int nchk = (nq + 30) / 31;
    unsigned col[NCH], row[NCH];  // lane = column (row): candidates of the chunk covering it
    int idx[NCH];                         // lane 30-i: hop index of candidate i of the chunk; lane 31: -1
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        col[c] = row[c] = 0;
        idx[c] = -1;
        if (c < nchk) {  // warp-uniform
            const int e = c * 31 + (30 - lane);
            unsigned cm = 0, rmk = 0;
            if (lane < 31 && e < nq) {
                const uint32_t wi = qi[e];
                idx[c] = (int)(wi & 0x3fffffu);
                cm = col_mask(qx[e], tx);  // queue entries meet the tile in x
                rmk = row_mask32(wi);
            }
            col[c] = transpose32(cm, lane);
            row[c] = transpose32(rmk, lane);
        }
    }
    // runs of columns (rows) with the same covering set
    bool dcol = lane == 0, drow = lane == 0;
#pragma unroll
    for (int c = 0; c < NCH; c++)
        if (c < nchk) {
            const unsigned pc = __shfl_up_sync(0xffffffffu, col[c], 1), pr = __shfl_up_sync(0xffffffffu, row[c], 1);
            if (lane > 0) {
                dcol = dcol || pc != col[c];
                drow = drow || pr != row[c];
            }
        }
    const unsigned cb = __ballot_sync(0xffffffffu, dcol), rb = __ballot_sync(0xffffffffu, drow);
    const int ncc = __popc(cb), nrc = __popc(rb);
    const unsigned le = lt | (1u << lane);
    const int mycc = __popc(cb & le) - 1;  // run of column `lane`
    if (dcol) ws.repc[mycc] = (uint8_t)lane;
    if (drow) ws.repr[__popc(rb & le) - 1] = (uint8_t)lane;
    if (FUSED) {
        runs[lane] = (uint8_t)mycc;
        runs[32 + lane] = (uint8_t)(__popc(rb & le) - 1);
        if (lane == 0) *dim = ncc | (nrc << 8);
    }
    __syncwarp();
    // cells are numbered densely, cell j = (row run j / ncc, column run j % ncc); 32 cells are resolved per fold pass,
    // up to CELL_CAP per batch. j / ncc by a multiply: exact for j < 2048 (ncc <= 32).
    const unsigned magic = c_magic[ncc];
    const int per_batch = (int)((CELL_CAP * magic) >> 16);  // row runs per batch: CELL_CAP / ncc >= 4
    for (int kr0 = 0; kr0 < nrc; kr0 += per_batch) {
        const int kr1 = min(kr0 + per_batch, nrc);
        const int ncell = (kr1 - kr0) * ncc;
        for (int j0 = 0; j0 < ncell; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < ncell;
            const int q = (int)(((unsigned)j * magic) >> 16);
            const int sc = valid ? ws.repc[j - q * ncc] : 0, sr = valid ? ws.repr[kr0 + q] : 0;
            Slots st = {-1, -1, -1, -1, 0};
#pragma unroll
            for (int c = 0; c < NCH; c++)
                if (c < nchk) {  // warp-uniform
                    const unsigned m = __shfl_sync(0xffffffffu, col[c], sc) & __shfl_sync(0xffffffffu, row[c], sr);
                    if (c == 0) {
                        fold<true, true>(st, m, idx[0]);
                    } else if (NCH <= 2 || __any_sync(0xffffffffu, m != 0)) {
                        // many chunks (dense small blocks): a pass's cells lie in one or two row runs, and hops arrive in
                        // raster order, so most chunks cover none of them and are skipped for the whole warp
                        fold<false, true>(st, m, idx[c]);
                    }
                }
            if (FUSED) {
                if (valid) cells[kr0 * ncc + j] = make_int4(st.s0, st.s1, st.s2, st.s3);
            } else if (valid) {
                ws.tab[j] = make_int4(st.s0, st.s1, st.s2, st.s3);
            }
        }
        if (FUSED) continue;  // warp-uniform
        __syncwarp();
        // the rows of these runs: one table read per run and column, one 128-bit store per pixel
        const int y0 = ws.repr[kr0], y1 = min(kr1 < nrc ? (int)ws.repr[kr1] : 32, nrows);
        int cell = mycc - ncc;
        int4 v = make_int4(-1, -1, -1, -1);
        int4 *o = out + (size_t)y0 * W;
        for (int y = y0; y < y1; y++) {
            if ((rb >> y) & 1u) {  // warp-uniform: a new row run starts
                cell += ncc;
                v = ws.tab[cell];
            }
            if (xin) st_cs_v4(o, v);
            o += W;
        }
        __syncwarp();
    }
}

// grid = (32-row bands, frames of the window x x-splits, streams). FUSED: the per-tile cell tables (common.cuh: TileCells) are
// stored instead of expanding them to pixels.
template <bool FUSED>
__global__ void __launch_bounds__(GRID_MAX_WARPS * 32, 2)
grid_kernel(WinParams p, int nxs, int NTC, const HopRect *__restrict__ hop_rects, const int32_t *__restrict__ nhops,
            const int2 *__restrict__ chunk_bbox, int4 *__restrict__ grid, int32_t *__restrict__ tc_dim, uint8_t *__restrict__ tc_runs,
            int4 *__restrict__ tc_cells, int NT) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch *wsa = reinterpret_cast<WarpScratch *>(smem);
    WarpScratch &ws = wsa[warp];
    uint32_t *tq_x = smem + nwarps * (sizeof(WarpScratch) / 4);  // [NTC][TQ_STRIDE] per-tile candidate queues: x0 | x1<<16
    uint32_t *tq_i = tq_x + NTC * TQ_STRIDE;                     // [NTC][TQ_STRIDE]                             hop | r0<<22 | r1<<27
    int32_t  *clist = (int32_t *)(tq_i + NTC * TQ_STRIDE);       // [CHUNK_CAP] surviving chunk ids
    __shared__ int32_t wcnt[GRID_MAX_WARPS];
    // phase 1b: entries warp w's chunk adds to tile t, one byte each (<= 32), a tile's 16 warps in one 128-bit word
    __shared__ __align__(16) uint8_t tcnt[MAX_TILES][GRID_MAX_WARPS];

    const unsigned lt = lanemask_lt();
    const int band = blockIdx.x;
    const int g = nxs == 1 ? blockIdx.y : blockIdx.y / nxs;
    const int xs = blockIdx.y - g * nxs;
    const int s = blockIdx.z;
    const int sg = s * p.n_out + g;
    const int ylo = band * SB_ROWS, yhi = min(ylo + SB_ROWS - 1, p.H - 1);
    const int X0 = xs * NTC * 32, X1 = min(X0 + NTC * 32, p.W) - 1;   // pixel columns of this CTA
    const int ntl = (X1 - X0 + 32) >> 5;                              // its tiles (the last split may own fewer than NTC)
    const int n_h = nhops[s * p.n_in + g];
    const HopRect *rects = hop_rects + (size_t)sg * p.max_hops;
    const int2 *bbox = chunk_bbox + (size_t)sg * p.max_chunks;
    const int nchunks = (n_h + 31) >> 5;

    // ---- phase 1a: ordered list of the 32-hop chunks whose extent touches the band (and this x-split) ---------
    int n_cl = 0;  // identical in every thread
    for (int base = 0; base < nchunks; base += blockDim.x) {
        const int c = base + threadIdx.x;
        bool pred = false;
        if (c < nchunks) {
            const int2 bb = __ldg(&bbox[c]);
            const int ymin = (int16_t)(bb.x & 0xffff), ymax = bb.x >> 16, xmin = (int16_t)(bb.y & 0xffff), xmax = bb.y >> 16;
            pred = ymax >= ylo && ymin <= yhi && xmax >= X0 && xmin <= X1;  // the x test matters for x-split frames
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
    const bool direct = n_cl > GRID_CHUNK_CAP;  // pathological input: every tile streams the whole hop list

    // ---- phase 1b: the band's hops go straight into per-tile queues, in list order ---------------------------------
    // A warp takes one surviving chunk per round (lower warp = earlier chunk). A hop meets one or two tiles (blocks are
    // narrower than a tile): match.any groups the lanes by tile, so a lane knows its rank inside the chunk for each of its
    // tiles and lane t knows the chunk's count for tile t; counts are prefixed across the warps of the round.
    int run_total = 0;  // lane t: entries queued so far in tile t (the same in every warp)
    uint4 all_m, before_m;  // byte j = 1 for warps j < nwarps / j < warp
    {
        auto ones_below = [](int n, int word) {  // bytes 4*word .. 4*word+3: 1 where the byte index is < n
            const int k = min(max(n - 4 * word, 0), 4);
            return k == 0 ? 0u : (0x01010101u >> (8 * (4 - k)));
        };
        all_m = make_uint4(ones_below(nwarps, 0), ones_below(nwarps, 1), ones_below(nwarps, 2), ones_below(nwarps, 3));
        before_m = make_uint4(ones_below(warp, 0), ones_below(warp, 1), ones_below(warp, 2), ones_below(warp, 3));
    }
    if (!direct) {
        const HopRect none = {0, 32767, -1, -32768};  // overlaps nothing
        HopRect r_next = none;
        int h_next = 0;
        if (warp < n_cl) {
            h_next = clist[warp] * 32 + lane;
            if (h_next < n_h) r_next = rects[h_next];
        }
        for (int base = 0; base < n_cl; base += nwarps) {
            const int ci = base + warp;
            bool pred = false;
            uint32_t wx = 0, wi = 0;
            int t0 = 0, t1 = 0;
            const HopRect r = r_next;
            const int h = h_next;
            r_next = none;
            if (ci + nwarps < n_cl) {  // the next round's rectangle is in flight across this round's barriers
                h_next = clist[ci + nwarps] * 32 + lane;
                if (h_next < n_h) r_next = rects[h_next];
            }
            {
                {
                    pred = r.y1 >= ylo && r.y0 <= yhi && r.x1 >= X0 && r.x0 <= X1;
                    if (pred) {
                        const int r0 = max((int)r.y0, ylo) - ylo, r1 = min((int)r.y1, yhi) - ylo;
                        wx = (uint32_t)(uint16_t)r.x0 | ((uint32_t)(uint16_t)r.x1 << 16);
                        wi = (uint32_t)h | ((uint32_t)r0 << 22) | ((uint32_t)r1 << 27);
                        t0 = (max((int)r.x0, X0) - X0) >> 5;
                        t1 = (min((int)r.x1, X1) - X0) >> 5;
                    }
                }
            }
            const bool two = pred && t1 != t0;
            const bool wide = __any_sync(0xffffffffu, pred && t1 - t0 > 1);  // blocks wider than a tile: generic ballots
            int rank0 = 0, rank1 = 0, mycnt = 0, tmin = 0, tmax = -1;
            if (!wide) {
                ws.amask[lane] = 0;
                ws.cmask[lane] = 0;
                __syncwarp();
                const unsigned p0 = __match_any_sync(0xffffffffu, pred ? t0 : 64 + lane);
                const unsigned p1 = __match_any_sync(0xffffffffu, two ? t1 : 64 + lane);
                if (pred && (p0 & lt) == 0) ws.amask[t0] = p0;
                if (two && (p1 & lt) == 0) ws.cmask[t1] = p1;
                __syncwarp();
                if (pred) rank0 = __popc((ws.amask[t0] | ws.cmask[t0]) & lt);
                if (two) rank1 = __popc((ws.amask[t1] | ws.cmask[t1]) & lt);
                mycnt = __popc(ws.amask[lane] | ws.cmask[lane]);
            } else {
                tmin = __reduce_min_sync(0xffffffffu, pred ? t0 : MAX_TILES);
                tmax = __reduce_max_sync(0xffffffffu, pred ? t1 : -1);
                for (int t = tmin; t <= tmax; t++) {
                    const unsigned b = __ballot_sync(0xffffffffu, pred && t0 <= t && t <= t1);
                    if (lane == t) mycnt = __popc(b);
                }
            }
            tcnt[lane][warp] = (uint8_t)mycnt;
            __syncthreads();
            // lane t: sum of the tile's counts over all warps and over the warps before this one (byte-wise dot products)
            const uint4 cw = *reinterpret_cast<const uint4 *>(tcnt[lane]);
            const int tot = __dp4a(cw.x, all_m.x, __dp4a(cw.y, all_m.y, __dp4a(cw.z, all_m.z, __dp4a(cw.w, all_m.w, 0u))));
            const int before = __dp4a(cw.x, before_m.x, __dp4a(cw.y, before_m.y, __dp4a(cw.z, before_m.z, __dp4a(cw.w, before_m.w, 0u))));
            const int basepos = run_total + before;
            run_total += tot;
            if (!wide) {
                const int b0 = __shfl_sync(0xffffffffu, basepos, t0), b1 = __shfl_sync(0xffffffffu, basepos, t1);
                if (pred) {
                    const int pos = b0 + rank0;
                    if (pos < TILE_Q) {
                        tq_x[t0 * TQ_STRIDE + pos] = wx;
                        tq_i[t0 * TQ_STRIDE + pos] = wi;
                    }
                }
                if (two) {
                    const int pos = b1 + rank1;
                    if (pos < TILE_Q) {
                        tq_x[t1 * TQ_STRIDE + pos] = wx;
                        tq_i[t1 * TQ_STRIDE + pos] = wi;
                    }
                }
            } else {
                for (int t = tmin; t <= tmax; t++) {
                    const bool in = pred && t0 <= t && t <= t1;
                    const unsigned b = __ballot_sync(0xffffffffu, in);
                    const int pos = __shfl_sync(0xffffffffu, basepos, t) + __popc(b & lt);
                    if (in && pos < TILE_Q) {
                        tq_x[t * TQ_STRIDE + pos] = wx;
                        tq_i[t * TQ_STRIDE + pos] = wi;
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- phase 2: a warp takes 32x32-pixel tiles of the band; no block barrier from here on -----------------------
    const int nrows = yhi - ylo + 1;
    for (int t = warp; t < ntl; t += nwarps) {
        const int tx = X0 + t * 32;
        const int x = tx + lane;
        const bool xin = x < p.W;
        int4 *out = FUSED ? nullptr : grid + ((size_t)sg * p.H + ylo) * p.W + x;
        uint32_t *qx = tq_x + t * TQ_STRIDE, *qi = tq_i + t * TQ_STRIDE;
        const int nq = __shfl_sync(0xffffffffu, run_total, t);
        const bool overflow = direct || nq > TILE_Q;
        // fused mode: this tile's cell table
        const size_t tile = FUSED ? ((size_t)sg * gridDim.x + band) * NT + (size_t)xs * NTC + t : 0;
        int32_t *dim = FUSED ? tc_dim + tile : nullptr;
        uint8_t *runs = FUSED ? tc_runs + tile * 64 : nullptr;
        int4 *cells = FUSED ? tc_cells + tile * MOVFE_TILE_CELLS : nullptr;
        if (!overflow && nq == 0) {
            // no hop touches the tile (I frames, intra blocks): the implicit fill
            if (FUSED) {
                if (lane == 0) *dim = 0;
            } else {
                const int4 v = make_int4(-1, -1, -1, -1);
                if (xin)
                    for (int y = 0; y < nrows; y++) st_cs_v4(out + (size_t)y * p.W, v);
            }
        } else if (!overflow) {
            // ---- fast path: specialised for the usual one or two chunks of candidates ------------------------------
            if (nq <= 62) fast_tile<2, FUSED>(ws, qx, qi, nq, tx, lane, lt, out, p.W, nrows, xin, dim, runs, cells);
            else fast_tile<TILE_CHUNKS, FUSED>(ws, qx, qi, nq, tx, lane, lt, out, p.W, nrows, xin, dim, runs, cells);
        } else {
            // slow path (more than 186 candidates in one tile, or the chunk list did not fit): one row at a time,
            // streaming the band's chunks again from global memory and folding 31 candidates per step.
            const int n_src = direct ? nchunks : n_cl;
            if (FUSED) {  // every pixel its own cell: identity run maps
                runs[lane] = (uint8_t)lane;
                runs[32 + lane] = (uint8_t)lane;
                if (lane == 0) *dim = 32 | (32 << 8);
            }
            for (int y = 0; y < nrows; y++) {
                Slots st = {-1, -1, -1, -1, 0};
                int nq2 = 0;
                for (int k = 0; k <= n_src; k++) {  // one extra, empty pass flushes the queue
                    bool pred = false;
                    uint32_t wx = 0, wi = 0;
                    if (k < n_src) {
                        const int e = (direct ? k : clist[k]) * 32 + lane;
                        if (e < n_h) {
                            const HopRect r = rects[e];
                            pred = r.y1 >= ylo + y && r.y0 <= ylo + y && r.x1 >= tx && r.x0 <= tx + 31;
                            wx = (uint32_t)(uint16_t)r.x0 | ((uint32_t)(uint16_t)r.x1 << 16);
                            wi = (uint32_t)e;
                        }
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, pred);
                    if (pred) {
                        const int pos = nq2 + __popc(b & lt);
                        qx[pos] = wx;
                        qi[pos] = wi;
                    }
                    nq2 += __popc(b);
                    __syncwarp();
                    const bool last = k == n_src;
                    while (nq2 >= 31 || (last && nq2 > 0)) {
                        const int take = min(nq2, 31);
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
                        // move the remainder to the front (reads 31.., writes 0..: disjoint for nq2-take <= 32)
                        uint32_t a = 0, c2 = 0;
                        const bool mv = lane < nq2 - take;
                        if (mv) {
                            a = qx[take + lane];
                            c2 = qi[take + lane];
                        }
                        __syncwarp();
                        if (mv) {
                            qx[lane] = a;
                            qi[lane] = c2;
                        }
                        nq2 -= take;
                        __syncwarp();
                    }
                }
                if (FUSED) cells[y * 32 + lane] = make_int4(st.s0, st.s1, st.s2, st.s3);
                else if (xin) st_cs_v4(out + (size_t)y * p.W, make_int4(st.s0, st.s1, st.s2, st.s3));
            }
        }
        __syncwarp();
    }
}

}  // namespace

int movfe_grid_launch(movfe_ctx *ctx, const WinParams &p, RasterBuf &w) {
    ProfScope prof(ctx, MOVFE_STAGE_GRID, ctx->raster_stream);
    prof.launches(1);
    // a CTA owns a 32-row band (of an x-split of at most 32 tiles when the frame is wider than 1024 px) = NTC tiles of
    // 32x32 pixels; warps per CTA: the largest divisor of the tile count that is <= 16 keeps every warp equally loaded
    const int nsb = (p.H + SB_ROWS - 1) / SB_ROWS;
    const int nxs = (ctx->NT + MAX_TILES - 1) / MAX_TILES;
    const int ntc = (ctx->NT + nxs - 1) / nxs;
    int nw = 4;
    for (int k = GRID_MAX_WARPS; k >= 4; k--)
        if (ntc % k == 0) {
            nw = k;
            break;
        }
    if (nw == 4 && ntc % 4 != 0 && ntc > 4) nw = 8;
    const size_t smem = (size_t)nw * sizeof(WarpScratch) + ((size_t)2 * ntc * TQ_STRIDE + GRID_CHUNK_CAP) * sizeof(uint32_t);
    if (smem > (size_t)ctx->smem_optin) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "grid: %zu bytes of shared memory per CTA exceed the device limit %d", smem, ctx->smem_optin);
    MOVFE_CUDA(ctx, optin_dynamic_smem(grid_kernel<false>, ctx->smem_optin));  // same value from every context
    MOVFE_CUDA(ctx, optin_dynamic_smem(grid_kernel<true>, ctx->smem_optin));
    if ((size_t)p.n_out * nxs > 65535 || p.S > 65535) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "grid: launch grid out of range");
    dim3 blocks(nsb, p.n_out * nxs, p.S);
    if (ctx->fused)
        grid_kernel<true><<<blocks, nw * 32, smem, ctx->raster_stream>>>(p, nxs, ntc, w.d_hop_rect, w.d_nhops, w.d_chunk_bbox, nullptr, w.d_tc_dim,
                                                                          w.d_tc_runs, w.d_tc_cells, ctx->NT);
    else
        grid_kernel<false><<<blocks, nw * 32, smem, ctx->raster_stream>>>(p, nxs, ntc, w.d_hop_rect, w.d_nhops, w.d_chunk_bbox, w.d_grid, nullptr,
                                                                           nullptr, nullptr, ctx->NT);
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}
