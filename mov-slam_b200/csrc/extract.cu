// extract.cu — reference-chained track propagation + EXPRESS descriptors.
// Replaces MOVExtractor::operator() (src/MOVExtractor.cc:63-455) and include/EXPRESS.h:79-192, batched over
// streams; frames of a stream are processed in order (frame f's table is the input of frame f+1).
// Compiled with -fmad=false (positions are binary32 sums that must round like the reference).
//
// Per frame, three launches. All three are bound by the latency of dependent gathers, so the work is
// arranged for loads in flight, not for bytes:
//   cand_kernel      a warp takes 32 tracks (in the reference's sorted order). Thread level: each lane walks its own
//                    track's chain order -> track -> slot-grid cell -> up to four hops. Warp level: per track, the loads
//                    of ALL candidate patches are issued back to back, then the warp-ballot EXPRESS descriptors are
//                    scored by Hamming distance. In-bounds tracks claim their hop's kps entry with an integer atomicMin
//                    on the sorted rank (order-independent result == the reference's first-come rule) and write their
//                    moved record to a staging table.
//   birth_kernel     unclaimed in-bounds candidate-keypoint blocks: compute_express + descriptor from one pass over
//                    the block's 17 columns.
//   finalize_kernel  one CTA per stream: ordered compaction of the survivors (a thread owns 8 consecutive ranks, one
//                    block scan per table), births appended in kps order with ids ++mCurrentId, optional coverage
//                    back-fill / I-frame seeding on the 16-px lattice, then the stable (age desc, popcount desc)
//                    order of the new table for the next frame: bitonic sort of unique 64-bit keys, register- and
//                    shuffle-resident for all but the widest steps.
// LK-carried features (cv::calcOpticalFlowPyrLK; MOVExtractor.cc:81-120,161-243,337-377): the LK arithmetic is OpenCV's
// and stays on the host, its RESULTS are handed in with movfe_set_lk_results and merged by finalize_kernel exactly where
// the reference merges them (I-frame carry-over of every track, coverage tracks after the propagated ones, lost-
// relocalisation seeds first). A frame that gets no results drops its carried tracks - the oracle's lk_status == NULL
// mode - and counts them (movfe_dropped_lk_tracks).
#include <algorithm>
#include <cstdio>

#include "common.cuh"
#include "express_lane.cuh"

namespace {

#ifndef MOVFE_CAND_WARPS
#define MOVFE_CAND_WARPS 4   // 128-thread CTAs pack better around the pose / finalize / raster CTAs they share SMs with
#endif
constexpr int CAND_WARPS = MOVFE_CAND_WARPS;
#ifndef MOVFE_FIN_THREADS
#define MOVFE_FIN_THREADS 512   // a 1024-thread CTA holds every register of its SM: with 512 two candidate CTAs fit beside it (3.22 against 3.27 ms per C2 step)
#endif
constexpr int FIN_THREADS = MOVFE_FIN_THREADS;
constexpr int FIN_WARPS = FIN_THREADS / 32;
constexpr int MAX_TRACKS_CAP = 8192;

struct ExtParams {
    int S, W, H, maxT, max_kps, max_hops, maxM, n_out, n_in, RING, TSLOTS;
    int s0;          // first video stream of this launch (the streams of a context are propagated in independent groups)
    int P;           // row pitch of the grey ring (power of two >= W)
    int fi;          // raster-window slot of this frame
    int gslot;       // ring slot of this frame (grey, flags)
    int tslot_prev, tslot_cur;
    int thr;         // EXPRESS threshold
    int has_grey;
    int use_lk;      // this frame consumes the host LK results installed with movfe_set_lk_results
    int fused;       // MOVFE_CFG_NO_GRID: slots come from the per-tile cell tables (common.cuh: resolve_slots), not from a slot grid
    int NT, tiles;   // tiles per tile row / per frame
    double cov_thr;
};

// raster results of one frame as the propagation kernels read them: the slot grid (grid-output mode) or the tile cell tables
struct SlotSource {
    const int4 *grid;
    const int32_t *tc_dim;
    const uint8_t *tc_runs;
    const int4 *tc_cells;
};

__device__ __forceinline__ TileCells frame_cells(const ExtParams &p, const SlotSource &src, int s) {
    const size_t fr = (size_t)s * p.n_out + p.fi;
    TileCells q;
    q.dim = src.tc_dim + fr * p.tiles;
    q.runs = src.tc_runs + fr * p.tiles * 64;
    q.cells = src.tc_cells + fr * p.tiles * MOVFE_TILE_CELLS;
    q.NT = p.NT;
    return q;
}

// Host LK hand-over (movfe_set_lk_results), per stream. n < 0: nothing installed for this frame.
struct LkBuf {
    int32_t *n;               // [S]
    uint8_t *status;          // [S][maxT]
    float2  *pts;             // [S][maxT]
    int32_t *n_reloc;         // [S]
    movfe_reloc_seed *reloc;  // [S][maxT]
    unsigned long long *dropped;
};

// ------------------------------------------------------------------------------------------------ EXPRESS -----
struct Band {
    int low, high;  // uint8 wrap-around already applied (EXPRESS.h:93-94)
};

// compute_center (EXPRESS.h:79-88): at(row = cols/2, col = rows/2) and its three upper-left neighbours.
__device__ __forceinline__ Band express_band(const uint8_t *__restrict__ roi, int stride, int rows, int cols, int thr) {
    const int cr = rows / 2, cc = cols / 2;
    const int center = ((int)roi[cc * stride + cr] + (int)roi[(cc - 1) * stride + (cr - 1)] + (int)roi[cc * stride + (cr - 1)] +
                        (int)roi[(cc - 1) * stride + cr]) / 4;
    Band b;
    b.low = (uint8_t)(center - thr);
    b.high = (uint8_t)(center + thr);
    return b;
}

// compute_descriptor (EXPRESS.h:90-110), one warp per block. `shift` = 1 reproduces the p++-before-read
// off-by-one of the row scans; `shift` = 0 gives the true block mask the diagonal walk reads.
// Returns the number of out-of-band pixels; desc (bit y*rows+x, OR-ed) is uniform across the warp.
//
// Fast paths for the four shapes H.264 produces (compile-time rows/cols: no integer division, and the bit
// scatter of the non-square shapes is a closed form of the ballot word); any other shape takes the generic loop.
template <int ROWS, int COLS, bool ROWMAJOR>
__device__ __forceinline__ int express_mask_t(const uint8_t *__restrict__ roi, int stride, Band bd, int shift,
                                              uint32_t desc[8], int lane) {
    constexpr int N = ROWS * COLS, ITERS = N / 32;
    constexpr int LC = COLS == 16 ? 4 : 3;
#pragma unroll
    for (int i = 0; i < 8; i++) desc[i] = 0;
    int vals[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; it++) {  // all loads first: ITERS independent requests in flight
        const int p = it * 32 + lane;
        vals[it] = roi[(p >> LC) * stride + (p & (COLS - 1)) + shift];
    }
    int count = 0;
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const unsigned b = __ballot_sync(0xffffffffu, bd.low > vals[it] || bd.high < vals[it]);
        count += __popc(b);
        if (ROWMAJOR || ROWS == COLS) {
            desc[it] = b;  // bit y*rows+x == raster index
        } else if (ROWS == 16 && COLS == 8) {
            // rows 4it..4it+3, 8 px each; bit = y*16 + x: two rows per word at offsets 0 and 16
            desc[2 * it] = (b & 0xffu) | (((b >> 8) & 0xffu) << 16);
            desc[2 * it + 1] = ((b >> 16) & 0xffu) | (((b >> 24) & 0xffu) << 16);
        } else {  // ROWS == 8 && COLS == 16: bit = y*8 + x, consecutive rows overlap by 8 bits and are OR-ed
            const unsigned c = (b & 0xffffu) | ((b >> 16) << 8);  // 24 bits starting at bit 16*it
            if (it & 1) {
                desc[it >> 1] |= c << 16;
                desc[(it >> 1) + 1] |= c >> 16;
            } else {
                desc[it >> 1] |= c;
            }
        }
    }
    return count;
}

__device__ __forceinline__ int express_mask_generic(const uint8_t *__restrict__ roi, int stride, int rows, int cols, Band bd,
                                                    int shift, bool rowmajor_bits, uint32_t desc[8], int lane);

__device__ __forceinline__ int express_mask(const uint8_t *__restrict__ roi, int stride, int rows, int cols, Band bd,
                                            int shift, bool rowmajor_bits, uint32_t desc[8], int lane) {
    if (rows == 16 && cols == 16) return express_mask_t<16, 16, true>(roi, stride, bd, shift, desc, lane);
    if (rows == 8 && cols == 8) return express_mask_t<8, 8, true>(roi, stride, bd, shift, desc, lane);
    if (rows == 16 && cols == 8)
        return rowmajor_bits ? express_mask_t<16, 8, true>(roi, stride, bd, shift, desc, lane)
                             : express_mask_t<16, 8, false>(roi, stride, bd, shift, desc, lane);
    if (rows == 8 && cols == 16)
        return rowmajor_bits ? express_mask_t<8, 16, true>(roi, stride, bd, shift, desc, lane)
                             : express_mask_t<8, 16, false>(roi, stride, bd, shift, desc, lane);
    return express_mask_generic(roi, stride, rows, cols, bd, shift, rowmajor_bits, desc, lane);
}

__device__ __forceinline__ int express_mask_generic(const uint8_t *__restrict__ roi, int stride, int rows, int cols, Band bd,
                                                    int shift, bool rowmajor_bits, uint32_t desc[8], int lane) {
#pragma unroll
    for (int i = 0; i < 8; i++) desc[i] = 0;
    int count = 0;
    const int n = rows * cols;
    const bool direct = rowmajor_bits || rows == cols;  // bit index == raster index p
#pragma unroll
    for (int it = 0; it < 8; it++) {
        if (it * 32 < n) {  // warp-uniform
            const int p = it * 32 + lane;
            bool oob = false;
            int y = 0, x = 0;
            if (p < n) {
                y = p / cols;
                x = p - y * cols;
                const int v = roi[y * stride + x + shift];
                oob = bd.low > v || bd.high < v;
            }
            const unsigned b = __ballot_sync(0xffffffffu, oob);
            count += __popc(b);
            if (direct) {
                desc[it] = b;
            } else {
                // generic shapes (16x8, 8x16, 4-px blocks): bit = y*rows + x, set bits are OR-ed (EXPRESS.h:106)
                unsigned rest = b;
                while (rest) {
                    const int l = __ffs(rest) - 1;
                    rest &= rest - 1;
                    const int pp = it * 32 + l;
                    const int yy = pp / cols, xx = pp - yy * cols;
                    const int bit = yy * rows + xx;
#pragma unroll
                    for (int w = 0; w < 8; w++)
                        if ((bit >> 5) == w) desc[w] |= 1u << (bit & 31);
                }
            }
        }
    }
    return count;
}

__device__ __forceinline__ int hamming256(const uint32_t a[8], const uint32_t b[8]) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
    return d;
}

__device__ __forceinline__ bool has_run(uint32_t bits, int r) {
    uint32_t x = bits;
    for (int k = 1; k < r; k++) x &= bits >> k;
    return x != 0;
}

// compute_express (EXPRESS.h:117-192), one warp per block. smem8: 8 words of per-warp scratch.
__device__ __forceinline__ bool express_test(const uint8_t *__restrict__ roi, int stride, int rows, int cols, int thr,
                                             uint32_t *smem8, int lane) {
    const Band bd = express_band(roi, stride, rows, cols, thr);
    uint32_t m[8];
    // pre-check (:122-139): the running count only grows and is tested per row, so "reaches precheck at some row
    // end" == "total >= precheck" (the uint8 counter cannot wrap before the break)
    const int f = express_mask(roi, stride, rows, cols, bd, 1, true, m, lane);
    const int precheck = (uint8_t)(rows * cols * .125);
    if (f < precheck) return false;
    if (!((rows == 8 || rows == 16) && (cols == 8 || cols == 16))) return false;  // diagonal() undefined (:66-69)
    express_mask(roi, stride, rows, cols, bd, 0, true, m, lane);  // true block mask, bit = y*cols + x
    __syncwarp();
    if (lane < 8) smem8[lane] = m[lane];
    __syncwarp();
    const int slices = rows + cols - 1;
    const int rounds = (int)roundf(slices * .25f);  // 8 / 6 / 4 for 31 / 23 / 15 slices (exact in float)
    const uint32_t valid = slices >= 32 ? 0xffffffffu : ((1u << slices) - 1u);
    bool ok = false;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const bool direction = a == 0;
        // lane d walks diagonal d (closed form of the tables EXPRESS.h:20-38, see oracle/express.cc)
        bool winbit = false;
        if (lane < slices) {
            const int d = lane;
            const int len = min(min(d + 1, rows), min(cols, slices - d));
            const int r0 = max(rows - 1 - d, 0);
            const int c1 = max(0, d - (rows - 1));
            const int c0 = direction ? c1 : cols - 1 - c1;
            const int dc = direction ? 1 : -1;
            int win = 0;
            for (int r = 0; r < len; r++) {
                const int bit = (r0 + r) * cols + (c0 + dc * r);
                win += (smem8[bit >> 5] >> (bit & 31)) & 1u;
            }
            winbit = win >= len - win;  // win >= loss (:171); "loss > win" is its complement (:179)
        }
        const uint32_t wb = __ballot_sync(0xffffffffu, winbit) & valid;
        // sticky run counters (:169-184): wins reaches `rounds` iff `rounds` consecutive win diagonals exist; the
        // early break (:185) only fires when the verdict is already false.
        if (has_run(wb, rounds) && has_run(~wb & valid, rounds)) ok = true;
    }
    __syncwarp();
    return ok;
}

__device__ __forceinline__ bool rect_in_bounds(int x, int y, int w, int h, int cols, int rows) {
    return x >= 0 && y >= 0 && (x + w) < cols && (y + h) < rows;
}

// ------------------------------------------------------------------------------------ batched patch access -----
// The propagation kernels are bound by the latency of dependent gathers, not by bandwidth, so every global load of a
// track's patches is ISSUED before any is consumed: patch_issue only loads, patch_words only consumes.
template <int ROWS, int COLS, int STRIDE>
__device__ __forceinline__ void patch_issue(const uint8_t *__restrict__ img, unsigned origin, int stride_rt, int shift, int lane,
                                            int (&vals)[ROWS * COLS / 32], int (&cen)[4]) {
    const int stride = STRIDE ? STRIDE : stride_rt;  // compile-time pitch: every row offset below is an immediate
    // one 64-bit address per patch and lane; with a compile-time pitch every load below is [base + immediate]
    constexpr int LC = COLS == 16 ? 4 : 3;
    const uint8_t *pc = img + origin;
    const uint8_t *pq = pc + ((lane >> LC) * stride + (lane & (COLS - 1)) + shift);
    const int step = (32 >> LC) * stride;
#pragma unroll
    for (int it = 0; it < ROWS * COLS / 32; it++) vals[it] = pq[it * step];
    constexpr int cr = ROWS / 2, cc = COLS / 2;  // compute_center (EXPRESS.h:79-88): at(row = cols/2, col = rows/2)
    cen[0] = pc[cc * stride + cr];
    cen[1] = pc[(cc - 1) * stride + (cr - 1)];
    cen[2] = pc[cc * stride + (cr - 1)];
    cen[3] = pc[(cc - 1) * stride + cr];
}

__device__ __forceinline__ Band band_of(const int (&cen)[4], int thr) {
    const int center = (cen[0] + cen[1] + cen[2] + cen[3]) / 4;
    Band b;
    b.low = (uint8_t)(center - thr);
    b.high = (uint8_t)(center + thr);
    return b;
}

// out-of-band ballots of a patch, raster order: bit p of word p/32 (p = y*COLS + x)
template <int N32>
__device__ __forceinline__ void patch_words(const int (&vals)[N32], Band bd, uint32_t (&b)[N32]) {
#pragma unroll
    for (int it = 0; it < N32; it++) b[it] = __ballot_sync(0xffffffffu, bd.low > vals[it] || bd.high < vals[it]);
}

// raster-order words -> the reference's descriptor layout, bit y*rows + x, OR-ed (EXPRESS.h:90-110)
template <int ROWS, int COLS>
__device__ __forceinline__ void desc_layout(const uint32_t (&b)[ROWS * COLS / 32], uint32_t (&desc)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) desc[i] = 0;
#pragma unroll
    for (int it = 0; it < ROWS * COLS / 32; it++) {
        if (ROWS == COLS) {
            desc[it] = b[it];
        } else if (ROWS == 16 && COLS == 8) {
            desc[2 * it] = (b[it] & 0xffu) | (((b[it] >> 8) & 0xffu) << 16);
            desc[2 * it + 1] = ((b[it] >> 16) & 0xffu) | (((b[it] >> 24) & 0xffu) << 16);
        } else {  // ROWS == 8 && COLS == 16
            const unsigned c = (b[it] & 0xffffu) | ((b[it] >> 16) << 8);
            if (it & 1) {
                desc[it >> 1] |= c << 16;
                desc[(it >> 1) + 1] |= c >> 16;
            } else {
                desc[it >> 1] |= c;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- cand_kernel -----
// A warp takes 32 consecutive tracks (sorted rank). Thread level: each lane walks its own track's dependent chain
// order -> track -> slot-grid cell -> up to four hops, so 32 chains are in flight per warp. Warp level: the tracks that
// need descriptors are visited one by one, all loads of all their candidate patches issued back to back.
constexpr int CAND_THREADS = CAND_WARPS * 32;
constexpr int CW_MXY = 0;    // [4] candidate rectangle origin, mx | my << 16
constexpr int CW_INFO = 4;   // need (4 bits) | mw << 8 | mh << 16
constexpr int CW_DESC = 5;   // [8] the track's previous descriptor (fetched at thread level, 32 tracks at once)
constexpr int CW_HOP = 13;   // [4][3] the candidate hops' displacement (2 floats) and kps index: parked here across the warp-level loop
constexpr int CW_WORDS = 25;

// Loads of one candidate patch for the propagation kernel (mask pixels at column offset 1). For the shapes whose four
// centre pixels lie inside the loaded 1..COLS columns the centre costs four shuffles instead of four more loads.
template <int ROWS, int COLS>
struct CentreInPatch {
    static constexpr bool value = COLS / 2 < ROWS;  // rows cols/2-1, cols/2 exist; columns rows/2-1, rows/2 are within 1..COLS
};

template <int ROWS, int COLS, int STRIDE>
__device__ __forceinline__ void cand_issue(const uint8_t *__restrict__ img, unsigned origin, int stride_rt, int lane,
                                           int (&vals)[ROWS * COLS / 32], int (&cen)[4]) {
    const int stride = STRIDE ? STRIDE : stride_rt;
    constexpr int LC = COLS == 16 ? 4 : 3;
    const uint8_t *pc = img + origin;
    const uint8_t *pq = pc + ((lane >> LC) * stride + (lane & (COLS - 1)) + 1);
    const int step = (32 >> LC) * stride;
#pragma unroll
    for (int it = 0; it < ROWS * COLS / 32; it++) vals[it] = pq[it * step];
    if (!CentreInPatch<ROWS, COLS>::value) {
        constexpr int cr = ROWS / 2, cc = COLS / 2;
        cen[0] = pc[cc * stride + cr];
        cen[1] = pc[(cc - 1) * stride + (cr - 1)];
        cen[2] = pc[cc * stride + (cr - 1)];
        cen[3] = pc[(cc - 1) * stride + cr];
    }
}

template <int ROWS, int COLS>
__device__ __forceinline__ Band cand_band(const int (&vals)[ROWS * COLS / 32], const int (&cen)[4], int thr) {
    if (!CentreInPatch<ROWS, COLS>::value) return band_of(cen, thr);
    // pixel (r, c) of the block sits at raster index r*COLS + (c-1) of the loaded columns
    constexpr int cr = ROWS / 2, cc = COLS / 2;
    constexpr int p0 = cc * COLS + (cr - 1), p1 = (cc - 1) * COLS + (cr - 2), p2 = cc * COLS + (cr - 2), p3 = (cc - 1) * COLS + (cr - 1);
    const int c4[4] = {__shfl_sync(0xffffffffu, vals[p0 >> 5], p0 & 31), __shfl_sync(0xffffffffu, vals[p1 >> 5], p1 & 31),
                       __shfl_sync(0xffffffffu, vals[p2 >> 5], p2 & 31), __shfl_sync(0xffffffffu, vals[p3 >> 5], p3 & 31)};
    return band_of(c4, thr);
}

// Candidates are evaluated two at a time (the loads of both are in flight together): four at a time keeps 48 registers of
// pixels alive and halves the warps an SM can hold, and occupancy is what hides the L2 latency of these gathers.
template <int ROWS, int COLS, int STRIDE>
__device__ __forceinline__ void cand_eval_pair(const uint8_t *__restrict__ img, int stride_rt, int thr, const int (&mxy)[4], unsigned need, int j0,
                                               const uint32_t (&pd)[8], int lane, uint32_t (&best_d)[8], int &best, int &chosen) {
    constexpr int IT = ROWS * COLS / 32;
    int vals[2][IT], cen[2][4];
    const int stride = STRIDE ? STRIDE : stride_rt;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int j = j0 + k;
        if ((need >> j) & 1u) {  // warp-uniform
            cand_issue<ROWS, COLS, STRIDE>(img, (unsigned)((mxy[j] >> 16) * stride + (int16_t)(mxy[j] & 0xffff)), stride_rt, lane, vals[k], cen[k]);
        } else {
#pragma unroll
            for (int it = 0; it < IT; it++) vals[k][it] = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) cen[k][q] = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int j = j0 + k;
        if ((need >> j) & 1u) {  // warp-uniform
            uint32_t b[IT], d[8];
            patch_words<IT>(vals[k], cand_band<ROWS, COLS>(vals[k], cen[k], thr), b);
            desc_layout<ROWS, COLS>(b, d);
            const int dist = hamming256(pd, d);
            // :292-296 strict '<' from 256. Candidate 0 is also the default choice (:270), so taking it at dist == 256
            // gives the reference's result (same hop, same descriptor, gate fails).
            if (j == 0 || dist < best) {
                best = dist;
                chosen = j;
#pragma unroll
                for (int q = 0; q < 8; q++) best_d[q] = d[q];
            }
        }
    }
}

template <int ROWS, int COLS, int STRIDE>
__device__ __forceinline__ int cand_eval(const uint8_t *__restrict__ img, int stride_rt, int thr, const int (&mxy)[4], unsigned need,
                                         const uint32_t (&pd)[8], int lane, uint32_t (&best_d)[8], int &best) {
    int chosen = -1;
    best = 256;
    cand_eval_pair<ROWS, COLS, STRIDE>(img, stride_rt, thr, mxy, need, 0, pd, lane, best_d, best, chosen);
    if (need & 0xcu) cand_eval_pair<ROWS, COLS, STRIDE>(img, stride_rt, thr, mxy, need, 2, pd, lane, best_d, best, chosen);  // warp-uniform
    return chosen;
}

// any other block shape (synthetic 4-px blocks, ...): one candidate after the other through the generic mask
__device__ __noinline__ int cand_eval_generic(const uint8_t *__restrict__ img, int stride, int thr, int mw, int mh, const int (&mxy)[4],
                                              unsigned need, const uint32_t (&pd)[8], int lane, uint32_t (&best_d)[8], int &best) {
    int chosen = -1;
    best = 256;
    for (int j = 0; j < 4; j++) {
        if (!((need >> j) & 1u)) continue;
        const uint8_t *roi = img + (size_t)(mxy[j] >> 16) * stride + (int16_t)(mxy[j] & 0xffff);
        uint32_t d[8];
        express_mask(roi, stride, mh, mw, express_band(roi, stride, mh, mw, thr), 1, false, d, lane);
        const int dist = hamming256(pd, d);
        if (j == 0 || dist < best) {
            best = dist;
            chosen = j;
#pragma unroll
            for (int k = 0; k < 8; k++) best_d[k] = d[k];
        }
    }
    return chosen;
}


// ---- asynchronous patch staging (PIPE variant of cand_kernel) -----------------------------------------------------------
// The candidate patches are gathers whose latency (an L2 round trip each) is what bounds propagation. Instead of holding the
// pixels of the patches in flight in registers, a warp keeps PF_DEPTH patch WINDOWS in flight as asynchronous global->shared
// copies (cp.async, 16 bytes per lane: one instruction fetches a 32-byte x 16-row window that contains the block's columns
// 1..cols and its centre pixels), and reads a window's pixels from shared memory when it has landed. No registers are tied
// up by the loads, so twice as many warps fit on an SM, and the depth of the prefetch no longer depends on how many
// candidates a track happens to have.
constexpr int PF_DEPTH = 8;   // even: two windows are consumed per step (one per half-warp)
constexpr int WIN_ROW = 32, WIN_ROWS = 16, WIN_BYTES = WIN_ROW * WIN_ROWS;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// window: rows my .. my+15, bytes xa .. xa+31 (xa a multiple of 16). Propagation reads a block's columns 1..cols: xa = (mx + 1) & ~15;
// births read columns 0..cols: xa = mx & ~15.
__device__ __forceinline__ void window_issue(const uint8_t *__restrict__ img, int stride, int xa, int my, uint8_t *win, int lane) {
    cp_async16(win + (lane >> 1) * WIN_ROW + (lane & 1) * 16, img + (size_t)(my + (lane >> 1)) * stride + xa + (lane & 1) * 16);
}

// descriptor of a ROWS x COLS block from its staged window (EXPRESS.h:79-110: centre, band, out-of-band bits of columns 1..COLS)
template <int ROWS, int COLS>
__device__ __forceinline__ void window_descriptor(const uint8_t *win, int mx, int thr, int lane, uint32_t (&desc)[8]) {
    constexpr int IT = ROWS * COLS / 32, LC = COLS == 16 ? 4 : 3;
    const int xo = mx - ((mx + 1) & ~15);  // window byte of the block's column 0 (-1 .. 14)
    constexpr int cr = ROWS / 2, cc = COLS / 2;  // compute_center: at(row = cols/2, col = rows/2) and its upper-left neighbours
    const int center = ((int)win[cc * WIN_ROW + xo + cr] + (int)win[(cc - 1) * WIN_ROW + xo + cr - 1] + (int)win[cc * WIN_ROW + xo + cr - 1] +
                        (int)win[(cc - 1) * WIN_ROW + xo + cr]) / 4;
    Band bd;
    bd.low = (uint8_t)(center - thr);
    bd.high = (uint8_t)(center + thr);
    const uint8_t *pq = win + (lane >> LC) * WIN_ROW + xo + 1 + (lane & (COLS - 1));
    int vals[IT];
#pragma unroll
    for (int it = 0; it < IT; it++) vals[it] = pq[it * (32 >> LC) * WIN_ROW];
    uint32_t b[IT];
    patch_words<IT>(vals, bd, b);
    desc_layout<ROWS, COLS>(b, desc);
}


// ---- SWAR evaluation of a staged window: a HALF-WARP per patch, a lane per block row, 16 pixels per lane -------------------
// One lane compares its row's pixels four to a 32-bit word (per-byte subtract and compare without carries between bytes) and
// ends with the row's out-of-band bits; the 256-bit descriptor is those rows side by side, so the Hamming distance to the
// track's previous descriptor is one popcount per lane and a half-warp sum. A quarter of the instructions of the
// pixel-per-lane form (eight byte loads, compares and ballots per lane and patch), and two patches per warp step.
struct BandSwar {
    uint32_t lowm;   // (low replicated) & 0x7f7f7f7f
    uint32_t nlow;   // ~(low replicated)
    uint32_t k4;     // per byte 0x7f - (t & 0x7f), t = high - low; 0x80 when every pixel is out of band (low > high: uint8 wrap)
    bool big;        // t >= 128
};
__device__ __forceinline__ BandSwar band_swar(int center, int thr) {
    const int low = (uint8_t)(center - thr), high = (uint8_t)(center + thr);  // EXPRESS.h:93-94, wrap included
    BandSwar b;
    const bool all = low > high;  // (low > p) || (high < p) holds for every p
    const uint32_t l4 = all ? 0u : (uint32_t)low * 0x01010101u;
    const int t = high - low;
    b.lowm = l4 & 0x7f7f7f7fu;
    b.nlow = ~l4;
    b.k4 = all ? 0x80808080u : (uint32_t)(0x7f - (t & 0x7f)) * 0x01010101u;
    b.big = !all && t >= 128;
    return b;
}
// bit 7 of byte k: pixel k of v lies outside [low, high]
__device__ __forceinline__ uint32_t oob4(uint32_t v, const BandSwar &b) {
    const uint32_t H = 0x80808080u;
    const uint32_t d = ((v | H) - b.lowm) ^ ((v ^ b.nlow) & H);  // per-byte v - low (mod 256)
    const uint32_t g = (d & ~H) + b.k4;                          // bit 7: (d & 0x7f) > (t & 0x7f)   [or always, k4 = 0x80]
    return (b.big ? (g & d) : (g | d)) & H;                      // d > t
}
// the four bit-7 flags of a word as a nibble (pixel k -> bit k)
__device__ __forceinline__ uint32_t nib(uint32_t r) { return ((r >> 7) * 0x10204080u) >> 28; }

// Out-of-band bits of row `r` of the rows x cols block whose column 0 sits at window byte xo (-1..14): columns 1..cols, i.e.
// window bytes xo+1 .. xo+cols (the p++-before-read of EXPRESS.h:98-109). 0 for rows outside the block.
__device__ __forceinline__ uint32_t row_bits(const uint8_t *win, int xo, int r, int rows, int cols, const BandSwar &b) {
    if (r >= rows) return 0u;
    const int sft = xo + 1;                      // 0..15
    const uint32_t *w = reinterpret_cast<const uint32_t *>(win + r * WIN_ROW) + (sft >> 2);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];   // (sft >> 2) + 4 <= 7: inside the 32-byte row
    const int bs = (sft & 3) * 8;
    const uint32_t p0 = __funnelshift_r(w0, w1, bs), p1 = __funnelshift_r(w1, w2, bs), p2 = __funnelshift_r(w2, w3, bs), p3 = __funnelshift_r(w3, w4, bs);
    const uint32_t m = nib(oob4(p0, b)) | (nib(oob4(p1, b)) << 4) | (nib(oob4(p2, b)) << 8) | (nib(oob4(p3, b)) << 12);
    return m & ((1u << cols) - 1u);
}

// Half-warp-wide: the 16 half-words of the block's descriptor (bit y*rows + x, OR-ed: EXPRESS.h:106), half-word q in lane q of
// the half. `row` = row_bits of this lane's row. rows == 16: half-word y is row y. rows == 8: rows are 8 bits apart and, for
// 16-column blocks, overlap their successor by 8 bits.
__device__ __forceinline__ uint32_t desc_halfword(uint32_t row, int rows, int hl) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, row, 1, 16);
    const uint32_t byte = (row & 0xffu) | (hl > 0 ? (up >> 8) : 0u);          // byte hl of the 8-row layout (hl = 0..8)
    const uint32_t b0 = __shfl_sync(0xffffffffu, byte, 2 * hl, 16), b1 = __shfl_sync(0xffffffffu, byte, 2 * hl + 1, 16);
    const uint32_t hw8 = hl < 8 ? (b0 | (b1 << 8)) : 0u;                        // lanes 8..15: 2*hl wraps, nothing there
    return rows == 16 ? row : hw8;
}

// Thread-level state of one track across the phases of the candidate kernels (lane = sorted rank inside the warp's 32-track chunk).
struct TrackState {
    uint4 a0, a1;    // pt_x, pt_y, mb.x | mb.y << 16, mb.w | mb.h << 16;  track_id, age, q_indx, flags
    float ptx, pty, hw, hh;
    unsigned need;   // candidates whose block lies inside the image (:286)
    int chosen;
    bool act, alive, multi, warp_job;
};

// ---- thread level: the track's own chain order -> track -> slots -> up to four hops; candidate rectangles, previous descriptor and the
// hops themselves are parked in the warp's shared-memory words smw[CW_*][lane] for the warp-level phase.
__device__ __forceinline__ TrackState track_chain(const ExtParams &p, const movfe_track *__restrict__ prev, const uint16_t *__restrict__ ord,
                                                  const int4 *__restrict__ g, const TileCells &tq, const movfe_hop *__restrict__ hp, bool has_img,
                                                  unsigned long long *__restrict__ stats, int i, int n_prev, int lane, int (*smw)[32]) {
    TrackState ts;
    ts.act = i < n_prev;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = make_uint4(0, 0, 0, 0), d0 = a0, d1 = a0;
    if (ts.act) {
        const int oidx = ord[i];
        const uint4 *tp = reinterpret_cast<const uint4 *>(prev + oidx);
        a0 = __ldg(tp);      // pt_x, pt_y, mb.x | mb.y << 16, mb.w | mb.h << 16
        a1 = __ldg(tp + 1);  // track_id, age, q_indx, flags
        if (has_img) {       // previous descriptor: one more dependent latency if it were fetched per track at warp level
            d0 = __ldg(tp + 2);
            d1 = __ldg(tp + 3);
        }
    }
    ts.a0 = a0;
    ts.a1 = a1;
    const float ptx = __uint_as_float(a0.x), pty = __uint_as_float(a0.y);
    const int mw = (int16_t)(a0.w & 0xffffu), mh = (int16_t)(a0.w >> 16);
    bool alive = ts.act && !(a1.w & MOVFE_TRACK_COVERAGE);  // :258-262 coverage tracks go to the host LK step
    int4 sl = make_int4(-1, -1, -1, -1);
    {
        const int x = (int)ptx, y = (int)pty;  // :264
        if (x < 0 || y < 0 || x >= p.W || y >= p.H) alive = false;  // unchecked .at<>() in the reference (UB)
        if (alive) sl = p.fused ? resolve_slots(tq, x, y) : __ldg(&g[(size_t)y * p.W + x]);
    }
    if (sl.x == -1) alive = false;  // :265-268
    const int sj[4] = {sl.x, sl.y, sl.z, sl.w};
    bool vj[4];
    vj[0] = alive;
#pragma unroll
    for (int j = 1; j < 4; j++) vj[j] = vj[j - 1] && sj[j] != -1;  // :277-278 stop at the first empty slot
    {   // workload counters (diagnostic): tracks looked up and their candidate hops
        const int nt = __reduce_add_sync(0xffffffffu, alive ? 1 : 0);
        const int nc = __reduce_add_sync(0xffffffffu, (int)vj[0] + (int)vj[1] + (int)vj[2] + (int)vj[3]);
        if (lane == 0 && nt) {
            atomicAdd(&stats[0], (unsigned long long)nt);
            atomicAdd(&stats[1], (unsigned long long)nc);
        }
    }
    int4 hv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) hv[j] = vj[j] ? __ldg(reinterpret_cast<const int4 *>(hp + sj[j])) : make_int4(0, 0, -1, 0);
    const float hw = (float)(mw / 2), hh = (float)(mh / 2);
    int mxy[4];
    unsigned need = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float px = __fadd_rn(ptx, __int_as_float(hv[j].x));  // :283
        const float py = __fadd_rn(pty, __int_as_float(hv[j].y));
        const int mx = (int)__fsub_rn(px, hw), my = (int)__fsub_rn(py, hh);  // :284
        mxy[j] = (int)((uint32_t)(mx & 0xffff) | ((uint32_t)my << 16));
        if (vj[j] && rect_in_bounds(mx, my, mw, mh, p.W, p.H)) need |= 1u << j;  // :286
        // the hop itself is parked in shared memory: only the chosen one is needed again, after the warp-level phase
        smw[CW_HOP + 3 * j + 0][lane] = hv[j].x;
        smw[CW_HOP + 3 * j + 1][lane] = hv[j].y;
        smw[CW_HOP + 3 * j + 2][lane] = hv[j].z;
    }
    // chosen candidate when no descriptor is involved: single-candidate pixels keep slot 0 (:270); with several
    // candidates and a flat image every distance is 0, so the first in-bounds one wins
    ts.chosen = (!has_img && sl.y >= 0 && need) ? __ffs(need) - 1 : 0;
    ts.multi = sl.y >= 0;
    ts.warp_job = alive && need != 0 && has_img;
    if (ts.warp_job) {
        // a later candidate whose block lands on the same pixels as an earlier in-bounds one has the same descriptor and
        // distance, and the strict '<' of :292 never prefers it: it is not evaluated
        unsigned need_eval = need;
#pragma unroll
        for (int j = 1; j < 4; j++)
#pragma unroll
            for (int k = 0; k < j; k++)
                if (((need >> k) & 1u) && mxy[j] == mxy[k]) need_eval &= ~(1u << j);
#pragma unroll
        for (int j = 0; j < 4; j++) smw[CW_MXY + j][lane] = mxy[j];
        smw[CW_INFO][lane] = (int)need_eval | (mw << 8) | (mh << 16);
        smw[CW_DESC + 0][lane] = (int)d0.x;
        smw[CW_DESC + 1][lane] = (int)d0.y;
        smw[CW_DESC + 2][lane] = (int)d0.z;
        smw[CW_DESC + 3][lane] = (int)d0.w;
        smw[CW_DESC + 4][lane] = (int)d1.x;
        smw[CW_DESC + 5][lane] = (int)d1.y;
        smw[CW_DESC + 6][lane] = (int)d1.z;
        smw[CW_DESC + 7][lane] = (int)d1.w;
    }
    ts.ptx = ptx;
    ts.pty = pty;
    ts.hw = hw;
    ts.hh = hh;
    ts.need = need;
    ts.alive = alive;
    return ts;
}

// ---- thread level: move, bounds, gate, claim (:301-316). my_best: distance of the chosen candidate's descriptor (with an image).
__device__ __forceinline__ void track_move(const ExtParams &p, const TrackState &ts, int my_best, bool has_img, movfe_track *__restrict__ st,
                                           int2 *__restrict__ ci, int32_t *__restrict__ cl, int i, int lane, int (*smw)[32]) {
    if (!ts.act) return;
    const int chosen = ts.chosen;
    const float cx = __fadd_rn(ts.ptx, __int_as_float(smw[CW_HOP + 3 * chosen + 0][lane]));  // :303
    const float cy = __fadd_rn(ts.pty, __int_as_float(smw[CW_HOP + 3 * chosen + 1][lane]));
    const int cd = smw[CW_HOP + 3 * chosen + 2][lane];
    const int cmx = (int)__fsub_rn(cx, ts.hw), cmy = (int)__fsub_rn(cy, ts.hh);  // :304
    const int cxy = (int)((uint32_t)(cmx & 0xffff) | ((uint32_t)cmy << 16));
    const bool inb = ts.alive && ((ts.need >> chosen) & 1u);  // :306 (the claim test itself happens in finalize)
    int fl = 0;
    if (inb) {
        fl = 1;
        // with an image the chosen candidate's descriptor was evaluated at warp level (need bit set => warp job)
        if (!has_img || my_best <= 40) fl |= 2;  // :311-316
        uint4 *o = reinterpret_cast<uint4 *>(st + i);
        o[0] = make_uint4(__float_as_uint(cx), __float_as_uint(cy), (uint32_t)cxy, ts.a0.w);
        o[1] = make_uint4(ts.a1.x, ts.a1.y + 1, (uint32_t)i, 0u);  // trackId, age + 1, qIndx, flags
        if (!has_img) {  // MV-only mode: descriptors are all zero
            o[2] = make_uint4(0, 0, 0, 0);
            o[3] = make_uint4(0, 0, 0, 0);
        }
        if (cd >= 0 && cd < p.max_kps) atomicMin(&cl[cd], i);  // first-come in sorted order (:306-309)
    }
    if (ts.a1.w & MOVFE_TRACK_COVERAGE) fl = 4;  // carried by the host LK step (:258-262), merged in finalize
    ci[i] = make_int2(ts.alive ? cd : -1, fl);
}


#ifndef CAND_MINB
#define CAND_MINB (16 / MOVFE_CAND_WARPS)  // 128 registers per thread
#endif
#ifndef CAND_PIPE_MINB
#define CAND_PIPE_MINB 6  // 80 registers per thread: no spills; 64 registers (8 CTAs) spill 200 bytes and measure 5-9 % slower
#endif
template <int PITCH, bool PIPE>
__global__ void __launch_bounds__(CAND_THREADS, PIPE ? CAND_PIPE_MINB : CAND_MINB)
cand_kernel(ExtParams p, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
            const uint16_t *__restrict__ order, SlotSource src, const movfe_hop *__restrict__ hops,
            const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, movfe_track *__restrict__ stage,
            int2 *__restrict__ cinfo, int32_t *__restrict__ claim, unsigned long long *__restrict__ stats) {
    __shared__ int sm[CAND_WARPS][CW_WORDS][32];
    __shared__ __align__(16) uint8_t swin[PIPE ? CAND_WARPS : 1][PF_DEPTH][WIN_BYTES];  // PIPE: patch windows in flight
    __shared__ uint8_t slist[PIPE ? CAND_WARPS : 1][128];                              // PIPE: the chunk's (track | candidate << 5) list
    pdl_wait();     // the previous frame's finalize_kernel wrote the tables read below
    pdl_trigger();  // after the wait: at most one dependent grid is resident and waiting
    const int s = p.s0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_prev = ntracks[s * p.TSLOTS + p.tslot_prev];
    if (!(fflags[s * p.RING + p.gslot] & MOVFE_FRAME_P)) return;  // I frame: nothing is propagated
    const movfe_track *prev = tracks + ((size_t)s * p.TSLOTS + p.tslot_prev) * p.maxT;
    const uint16_t *ord = order + (size_t)s * p.maxT;
    const int4 *g = p.fused ? nullptr : src.grid + ((size_t)s * p.n_out + p.fi) * ((size_t)p.W * p.H);
    TileCells tq = {};
    if (p.fused) tq = frame_cells(p, src, s);
    const movfe_hop *hp = hops + ((size_t)s * p.n_out + p.fi) * p.max_hops;
    const uint8_t *img = p.has_grey ? grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.P * p.H) : nullptr;
    movfe_track *st = stage + (size_t)s * p.maxT;
    int2 *ci = cinfo + (size_t)s * p.maxT;
    int32_t *cl = claim + (size_t)s * p.max_kps;  // lbFound (MOVExtractor.cc:253), one entry per kps of the frame

    for (int c = blockIdx.x * CAND_WARPS + warp; c * 32 < n_prev; c += gridDim.x * CAND_WARPS) {
        const int i = c * 32 + lane;  // sorted rank
        TrackState ts = track_chain(p, prev, ord, g, tq, hp, img != nullptr, stats, i, n_prev, lane, sm[warp]);
        const bool warp_job = ts.warp_job, multi = ts.multi;
        int chosen = ts.chosen;
        __syncwarp();
        // ---- warp level: descriptors of the candidate patches -------------------------------------------------------
        // The winning descriptor of track t goes straight to the staging record of its rank (written for every evaluated
        // track; finalize only reads the records of tracks that pass bounds and gate).
        int my_best = 0;
        unsigned todo = __ballot_sync(0xffffffffu, warp_job);
        unsigned odd = 0;  // tracks whose block is none of the four H.264 shapes: second loop (its arrays live in local memory)
        if (PIPE) {
            // the chunk's evaluations as one flat list in (track, candidate) order: PF_DEPTH windows are always in flight
            const int stride = PITCH ? PITCH : p.P;
            int info = warp_job ? sm[warp][CW_INFO][lane] : 0;
            const bool std_shape = ((info >> 8 & 0xff) == 16 || (info >> 8 & 0xff) == 8) && ((info >> 16) == 16 || (info >> 16) == 8);
            odd = __ballot_sync(0xffffffffu, warp_job && !std_shape);
            const unsigned mine = (warp_job && std_shape) ? (unsigned)(info & 0xf) : 0u;
            int cnt = __popc(mine), incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const int n_ev = __shfl_sync(0xffffffffu, incl, 31);
            {
                int pos = incl - cnt;
                unsigned m = mine;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    slist[warp][pos++] = (uint8_t)(lane | (j << 5));
                }
            }
            __syncwarp();
            auto issue = [&](int k) {
                const int e = slist[warp][k], t = e & 31, j = e >> 5;
                const int m = sm[warp][CW_MXY + j][t];
                window_issue(img, stride, ((int16_t)(m & 0xffff) + 1) & ~15, m >> 16, swin[warp][k % PF_DEPTH], lane);
            };
#pragma unroll
            for (int k = 0; k < PF_DEPTH; k++) {
                if (k < n_ev) issue(k);
                cp_async_commit();
            }
            // two evaluations per step: half-warp h takes list entry k + h, lane hl of the half takes block row hl
            const int half = lane >> 4, hl = lane & 15;
            int cur_t = -1, best = 256, ch = -1;
            auto flush = [&]() {  // the finished track's verdict goes to its lane (its descriptor is already in its staging record)
                if (lane == cur_t) {
                    if (multi && ch >= 0) chosen = ch;  // single-candidate pixels never compare (:272): slot 0 stays chosen
                    my_best = best;
                }
            };
            const bool swar_ok = p.thr >= 0;  // (the band arithmetic below covers every threshold; kept as a switch for A/B)
            for (int k = 0; k < n_ev; k += 2) {
                cp_async_wait<PF_DEPTH - 2>();  // all but the newest PF_DEPTH-2 groups have landed: windows k and k+1 are complete
                __syncwarp();
                const int ke = k + half;
                const bool on = ke < n_ev;
                const int e = on ? slist[warp][ke] : 0, t = e & 31, j = e >> 5;
                // (an idle half - odd list length - computes on a 16x16 dummy: the per-track words of lane 0 may be stale)
                const int tinfo = on ? sm[warp][CW_INFO][t] : ((16 << 8) | (16 << 16));
                const int cols = (tinfo >> 8) & 0xff, rows = tinfo >> 16;   // block width / height
                const int mx = on ? (int16_t)(sm[warp][CW_MXY + j][t] & 0xffff) : 0;
                const uint8_t *win = swin[warp][ke % PF_DEPTH];
                const int xo = mx - ((mx + 1) & ~15);
                // compute_center (EXPRESS.h:79-88): at(row = cols/2, col = rows/2) and its upper-left neighbours
                const int cr = rows >> 1, cc = cols >> 1;
                const int center = ((int)win[cc * WIN_ROW + xo + cr] + (int)win[(cc - 1) * WIN_ROW + xo + cr - 1] +
                                    (int)win[cc * WIN_ROW + xo + cr - 1] + (int)win[(cc - 1) * WIN_ROW + xo + cr]) / 4;
                const BandSwar bd = band_swar(center, p.thr);
                const uint32_t row = swar_ok ? row_bits(win, xo, hl, rows, cols, bd) : 0u;
                const uint32_t hw = desc_halfword(row, rows, hl);
                // previous descriptor, half-word hl (bitset<256> bit i at word i>>5, bit i&31)
                const uint32_t pw = (uint32_t)sm[warp][CW_DESC + (hl >> 1)][t];
                const uint32_t pdh = (hl & 1) ? (pw >> 16) : (pw & 0xffffu);
                // both halves' distances from one warp-wide sum: half 1 counts in the upper 16 bits (a distance is at most 256)
                const int part = on ? __popc(hw ^ pdh) : 0;
                const unsigned both = __reduce_add_sync(0xffffffffu, (unsigned)part << (16 * half));
                const int d0v = (int)(both & 0xffffu), d1v = (int)(both >> 16);
                const int e1 = __shfl_sync(0xffffffffu, e, 16), e0 = __shfl_sync(0xffffffffu, e, 0);
#pragma unroll
                for (int h = 0; h < 2; h++) {  // the two results are folded in list order (warp-uniform)
                    if (k + h >= n_ev) break;
                    const int eh = h ? e1 : e0, th = eh & 31, jh = eh >> 5, dh = h ? d1v : d0v;
                    if (th != cur_t) {
                        if (cur_t >= 0) flush();
                        cur_t = th;
                        best = 256;
                        ch = -1;
                    }
                    // :292-296 strict '<' from 256; candidate 0 is also the default choice (:270), see cand_eval_pair
                    if (jh == 0 || dh < best) {
                        best = dh;
                        ch = jh;
                        // the winner so far: its descriptor goes to the track's staging record, half-word hl from lane hl of half h
                        // (a later, better candidate of the same track overwrites it: same warp, program order)
                        if (half == h) reinterpret_cast<uint16_t *>((st + c * 32 + th)->desc)[hl] = (uint16_t)hw;
                    }
                }
                __syncwarp();  // every lane has read windows k, k+1 before their buffers are refilled
                if (k + PF_DEPTH < n_ev) issue(k + PF_DEPTH);
                cp_async_commit();
                if (k + PF_DEPTH + 1 < n_ev) issue(k + PF_DEPTH + 1);
                cp_async_commit();
            }
            if (cur_t >= 0) flush();
            cp_async_wait<0>();
            todo = 0;
        }
        while (!PIPE && todo) {
            const int t = __ffs(todo) - 1;
            todo &= todo - 1;
            const int info = sm[warp][CW_INFO][t];
            const unsigned nd = info & 0xf;
            const int tw = (info >> 8) & 0xff, th = info >> 16;
            const bool std_shape = (tw == 16 || tw == 8) && (th == 16 || th == 8);
            if (!std_shape) {  // warp-uniform
                odd |= 1u << t;
                continue;
            }
            int cm[4];
#pragma unroll
            for (int j = 0; j < 4; j++) cm[j] = sm[warp][CW_MXY + j][t];
            uint32_t pd[8];
#pragma unroll
            for (int k = 0; k < 8; k++) pd[k] = (uint32_t)sm[warp][CW_DESC + k][t];
            uint32_t bd[8];
            int best, ch;
            if (tw == 16 && th == 16) ch = cand_eval<16, 16, PITCH>(img, p.P, p.thr, cm, nd, pd, lane, bd, best);
            else if (tw == 8 && th == 8) ch = cand_eval<8, 8, PITCH>(img, p.P, p.thr, cm, nd, pd, lane, bd, best);
            else if (tw == 8 && th == 16) ch = cand_eval<16, 8, PITCH>(img, p.P, p.thr, cm, nd, pd, lane, bd, best);
            else ch = cand_eval<8, 16, PITCH>(img, p.P, p.thr, cm, nd, pd, lane, bd, best);
            if (lane == t) {
                // single-candidate pixels never compare (:272): slot 0 stays chosen; its descriptor is the one evaluated
                if (multi && ch >= 0) chosen = ch;
                my_best = best;
                uint4 *o = reinterpret_cast<uint4 *>(st + i);
                o[2] = make_uint4(bd[0], bd[1], bd[2], bd[3]);
                o[3] = make_uint4(bd[4], bd[5], bd[6], bd[7]);
            }
        }
        while (odd) {
            const int t = __ffs(odd) - 1;
            odd &= odd - 1;
            int cm[4];
#pragma unroll
            for (int j = 0; j < 4; j++) cm[j] = sm[warp][CW_MXY + j][t];
            const int info = sm[warp][CW_INFO][t];
            uint32_t pd[8];
#pragma unroll
            for (int k = 0; k < 8; k++) pd[k] = (uint32_t)sm[warp][CW_DESC + k][t];
            uint32_t bd[8];
            int best;
            const int ch = cand_eval_generic(img, p.P, p.thr, (info >> 8) & 0xff, info >> 16, cm, info & 0xf, pd, lane, bd, best);
            if (lane == t) {
                if (multi && ch >= 0) chosen = ch;
                my_best = best;
                uint4 *o = reinterpret_cast<uint4 *>(st + i);
                o[2] = make_uint4(bd[0], bd[1], bd[2], bd[3]);
                o[3] = make_uint4(bd[4], bd[5], bd[6], bd[7]);
            }
        }
        __syncwarp();
        ts.chosen = chosen;
        track_move(p, ts, my_best, img != nullptr, st, ci, cl, i, lane, sm[warp]);
    }
}

// ---------------------------------------------------------------------------------------- cand_lane_kernel -----
// The same propagation step with the descriptors evaluated at THREAD level (express_lane.cuh): a lane pair per candidate block,
// one lane per 8-row half, 16 blocks per warp step. The warp-level forms above pay a set of warp-wide shuffles, ballots and
// bookkeeping per block (cand_kernel<.., true>: ~85 warp instructions per block, instruction issue is what bounds it); here the
// block's pixels are compared four to an instruction by one lane and nothing crosses lanes until the distance of a block is the
// sum of its two halves. The windows are staged with 8-byte cp.async copies, 24 bytes per row from an 8-byte aligned column, so
// that a row is three conflict-free 64-bit shared loads.
constexpr int LN_BATCH = 16;                       // blocks per warp step
#ifndef LN_BUFS
#define LN_BUFS 1   // 2: the next step's windows are staged while this step is evaluated - measured 2.5 % slower (3 CTAs per SM instead of 5)
#endif
constexpr int LN_LIST = 128;                       // evaluations of a 32-track chunk (4 candidates each)
constexpr int CW_BEST = CW_WORDS;                  // one more parked word per track: (best distance << 3 | candidate) of the evaluations so far
constexpr int CWL_WORDS = CW_WORDS + 1;
constexpr size_t LN_SMEM = (size_t)CAND_WARPS * LN_BUFS * LN_BATCH * xl::WIN_STRIDE * sizeof(uint32_t);
#ifndef CAND_LANE_MINB
#define CAND_LANE_MINB (LN_BUFS == 2 ? 3 : 5)  // CTAs per SM that the shared memory of the window slots allows
#endif

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Stages the windows of `n` (<= BATCH) list entries into the warp's window slots: 48 copies of 8 bytes per window (16 rows x 3),
// two windows per three warp-wide copies, no predicates: an odd last window is fetched twice (the second copy lands in an unused
// slot), and the rows below an 8-row block are fetched like the others (the grey ring ends with 16 rows of slack).
// org[k] = byte offset of the window's first row in the grey plane.
struct StageLane {
    unsigned src[3], dst[3];  // byte offsets of this lane's three copies: source relative to the window origin, destination in the pair's slots
    bool second[3];           // the copy belongs to the second window of the pair
};
__device__ __forceinline__ StageLane stage_lane(int stride, int lane) {
    StageLane sl;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int m = lane + 32 * i;  // 0..95: copy m of the window pair
        const int which = m >= 48 ? 1 : 0, mm = m - 48 * which, row = mm / 3, part = mm - 3 * row;
        sl.second[i] = which != 0;
        sl.src[i] = (unsigned)(row * stride + part * 8);
        sl.dst[i] = (unsigned)(which * (xl::WIN_STRIDE * 4) + mm * 8);
    }
    return sl;
}
template <int BATCH>
__device__ __forceinline__ void stage_windows(const uint8_t *__restrict__ img, const StageLane &sl, const uint32_t *org, int n, uint32_t *win) {
#pragma unroll
    for (int pp = 0; pp < BATCH / 2; pp++) {
        if (2 * pp >= n) break;  // warp-uniform
        const uint32_t o0 = org[2 * pp], o1 = org[min(2 * pp + 1, n - 1)];
        uint8_t *dst = reinterpret_cast<uint8_t *>(win) + pp * (2 * xl::WIN_STRIDE * 4);
#pragma unroll
        for (int i = 0; i < 3; i++) cp_async8(dst + sl.dst[i], img + ((sl.second[i] ? o1 : o0) + sl.src[i]));
    }
    cp_async_commit();
}

template <int PITCH>
__global__ void __launch_bounds__(CAND_THREADS, CAND_LANE_MINB)
cand_lane_kernel(ExtParams p, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
                 const uint16_t *__restrict__ order, SlotSource src, const movfe_hop *__restrict__ hops,
                 const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, movfe_track *__restrict__ stage,
                 int2 *__restrict__ cinfo, int32_t *__restrict__ claim, unsigned long long *__restrict__ stats) {
    __shared__ int sm[CAND_WARPS][CWL_WORDS][32];
    // window slots, double-buffered: the next step's windows land while this step is evaluated (dynamic: above the static limit)
    extern __shared__ __align__(16) uint32_t swin_all[];
    uint32_t (*swin)[LN_BUFS][LN_BATCH * xl::WIN_STRIDE] = reinterpret_cast<uint32_t (*)[LN_BUFS][LN_BATCH * xl::WIN_STRIDE]>(swin_all);
    __shared__ uint32_t sorg[CAND_WARPS][LN_LIST];   // window origin of every evaluation of the chunk, in (track, candidate) order
    __shared__ uint8_t slist[CAND_WARPS][LN_LIST];   // track | candidate << 5
    pdl_wait();     // the previous frame's finalize_kernel wrote the tables read below
    pdl_trigger();
    const int s = p.s0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_prev = ntracks[s * p.TSLOTS + p.tslot_prev];
    if (!(fflags[s * p.RING + p.gslot] & MOVFE_FRAME_P)) return;  // I frame: nothing is propagated
    const movfe_track *prev = tracks + ((size_t)s * p.TSLOTS + p.tslot_prev) * p.maxT;
    const uint16_t *ord = order + (size_t)s * p.maxT;
    const int4 *g = p.fused ? nullptr : src.grid + ((size_t)s * p.n_out + p.fi) * ((size_t)p.W * p.H);
    TileCells tq = {};
    if (p.fused) tq = frame_cells(p, src, s);
    const movfe_hop *hp = hops + ((size_t)s * p.n_out + p.fi) * p.max_hops;
    const uint8_t *img = grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.P * p.H);  // this kernel is only launched with an image
    movfe_track *st = stage + (size_t)s * p.maxT;
    int2 *ci = cinfo + (size_t)s * p.maxT;
    int32_t *cl = claim + (size_t)s * p.max_kps;
    const int stride = PITCH ? PITCH : p.P;
    const int q = lane >> 1, half = lane & 1;
    const StageLane stl = stage_lane(stride, lane);

    for (int c = blockIdx.x * CAND_WARPS + warp; c * 32 < n_prev; c += gridDim.x * CAND_WARPS) {
        const int i = c * 32 + lane;  // sorted rank
        TrackState ts = track_chain(p, prev, ord, g, tq, hp, true, stats, i, n_prev, lane, sm[warp]);
        // ---- the chunk's evaluations as one flat list in (track, candidate) order ------------------------------------------------
        const int info = ts.warp_job ? sm[warp][CW_INFO][lane] : 0;
        const int tw = (info >> 8) & 0xff, th = info >> 16;
        const bool std_shape = (tw == 16 || tw == 8) && (th == 16 || th == 8);
        unsigned odd = __ballot_sync(0xffffffffu, ts.warp_job && !std_shape);  // none of the four H.264 shapes: warp-level loop below
        const unsigned mine = (ts.warp_job && std_shape) ? (unsigned)(info & 0xf) : 0u;
        const int cnt = __popc(mine);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int n_ev = __shfl_sync(0xffffffffu, incl, 31);
        {
            int pos = incl - cnt;
            unsigned m = mine;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int mv = sm[warp][CW_MXY + j][lane];
                const int mx = (int16_t)(mv & 0xffff), my = mv >> 16;
                slist[warp][pos] = (uint8_t)(lane | (j << 5));
                sorg[warp][pos] = (uint32_t)(my * stride + ((mx + 1) & ~7));
                pos++;
            }
        }
        __syncwarp();
        // ---- lane level: a lane pair per block, 16 blocks per step --------------------------------------------------------------
        int carry_t = -1, carry_key = 0;   // the track whose evaluations straddle two steps: its best key so far (warp-uniform)
        if (LN_BUFS == 2 && n_ev > 0) stage_windows<LN_BATCH>(img, stl, sorg[warp], min(LN_BATCH, n_ev), swin[warp][0]);
        for (int b0 = 0, bi = 0; b0 < n_ev; b0 += LN_BATCH, bi++) {
            const int nb = min(LN_BATCH, n_ev - b0);
            if (LN_BUFS == 2) {  // the next step's windows: one commit group per step, empty at the end
                if (b0 + LN_BATCH < n_ev) stage_windows<LN_BATCH>(img, stl, sorg[warp] + b0 + LN_BATCH, min(LN_BATCH, n_ev - b0 - LN_BATCH), swin[warp][(bi + 1) & 1]);
                else cp_async_commit();
            } else {
                stage_windows<LN_BATCH>(img, stl, sorg[warp] + b0, nb, swin[warp][0]);
            }
            const bool on = q < nb;
            const int e = on ? slist[warp][b0 + q] : 0, t = e & 31, j = e >> 5;
            // (an idle pair computes on a 16x16 dummy: the per-track words of lane 0 may be stale)
            const int tinfo = on ? sm[warp][CW_INFO][t] : ((16 << 8) | (16 << 16));
            const int cols = (tinfo >> 8) & 0xff, rows = tinfo >> 16;
            const int mx = on ? (int16_t)(sm[warp][CW_MXY + j][t] & 0xffff) : 0;
            const int xw = (mx + 1) & ~7;
            uint32_t pd[4];
#pragma unroll
            for (int k = 0; k < 4; k++) pd[k] = (uint32_t)sm[warp][CW_DESC + 4 * half + k][t];
            cp_async_wait<LN_BUFS - 1>();  // all but the newest group: this step's windows have landed
            __syncwarp();
            const uint32_t *win = swin[warp][LN_BUFS == 2 ? (bi & 1) : 0] + q * xl::WIN_STRIDE;
            const xl::Band bd = xl::band_of(xl::centre_of(win, mx - xw, rows, cols), p.thr);
            uint32_t d[4];
            xl::half_descriptor(win, mx + 1 - xw, rows, cols, half, bd, d);
            if (half && rows == 8) d[0] = d[1] = d[2] = d[3] = 0;  // an 8-row block has no second half
            // (desc1 ^ desc2).count() (EXPRESS.h:112-115): this half's four words, the block's distance is the sum of both halves
            int part = __popc(d[0] ^ pd[0]) + __popc(d[1] ^ pd[1]) + __popc(d[2] ^ pd[2]) + __popc(d[3] ^ pd[3]);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            // :292-296: candidates in order, strict '<' from 256, candidate 0 is the default choice (:270): the winner is the smallest
            // (distance, candidate) key. Evaluations of a track are adjacent list entries: a segmented minimum over <= 3 neighbours.
            const int key = on ? ((part << 3) | j) : 0x7fff;
            const int tk = on ? ((t << 16) | key) : (0xff << 16) | 0x7fff;
            int seg_min = 0x7fff;  // best key of the OTHER evaluations of this track in this step
#pragma unroll
            for (int dlt = 1; dlt <= 3; dlt++) {
                const int up = __shfl_up_sync(0xffffffffu, tk, 2 * dlt), dn = __shfl_down_sync(0xffffffffu, tk, 2 * dlt);
                if (q >= dlt && (up >> 16) == t) seg_min = min(seg_min, up & 0xffff);
                if (q + dlt < LN_BATCH && (dn >> 16) == t) seg_min = min(seg_min, dn & 0xffff);
            }
            const int prior = (on && t == carry_t) ? carry_key : 0x7fff;  // this track's best key of the previous step
            const bool winner = on && key < seg_min && key < prior;
            if (winner) {
                // the best descriptor so far goes to the track's staging record (a better candidate of a later step overwrites it:
                // same warp, program order); the key is parked for the track's own lane
                reinterpret_cast<uint4 *>((st + c * 32 + t)->desc)[half] = make_uint4(d[0], d[1], d[2], d[3]);
                if (!half) sm[warp][CW_BEST][t] = key;
            }
            // the last block's track may continue in the next step
            const int last = 2 * (nb - 1);
            const int lt = __shfl_sync(0xffffffffu, t, last), lk = __shfl_sync(0xffffffffu, min(min(key, seg_min), prior), last);
            carry_t = lt;
            carry_key = lk;
            __syncwarp();  // every lane has read its window before the slots are refilled
        }
        __syncwarp();
        int my_best = 0;
        if (ts.warp_job && std_shape) {
            const int key = sm[warp][CW_BEST][lane];
            const int dist = key >> 3, j = key & 7;
            // single-candidate pixels never compare (:272): slot 0 stays chosen. A first evaluated candidate j > 0 at distance 256
            // is not taken by the strict '<' (its track fails the bounds test of :306 at slot 0 anyway).
            if (ts.multi && (dist < 256 || j == 0)) ts.chosen = j;
            my_best = dist;
        }
        while (odd) {
            const int t = __ffs(odd) - 1;
            odd &= odd - 1;
            int cm[4];
#pragma unroll
            for (int j = 0; j < 4; j++) cm[j] = sm[warp][CW_MXY + j][t];
            const int oinfo = sm[warp][CW_INFO][t];
            uint32_t pd[8];
#pragma unroll
            for (int k = 0; k < 8; k++) pd[k] = (uint32_t)sm[warp][CW_DESC + k][t];
            uint32_t bd[8];
            int best;
            const int ch = cand_eval_generic(img, p.P, p.thr, (oinfo >> 8) & 0xff, oinfo >> 16, cm, oinfo & 0xf, pd, lane, bd, best);
            if (lane == t) {
                if (ts.multi && ch >= 0) ts.chosen = ch;
                my_best = best;
                uint4 *o = reinterpret_cast<uint4 *>(st + i);
                o[2] = make_uint4(bd[0], bd[1], bd[2], bd[3]);
                o[3] = make_uint4(bd[4], bd[5], bd[6], bd[7]);
            }
        }
        __syncwarp();
        track_move(p, ts, my_best, true, st, ci, cl, i, lane, sm[warp]);
    }
}

// ------------------------------------------------------------------------------------------ birth_kernel -----
// compute_express (EXPRESS.h:117-192) + descriptor of one block from ONE pass over its 17 columns: the true block mask
// (diagonal walk) and the p++-shifted mask (pre-check, descriptor) differ by one column.
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

template <int ROWS, int COLS, int STRIDE>
__device__ __forceinline__ bool express_birth(const uint8_t *__restrict__ img, unsigned origin, int stride_rt, int thr, uint32_t *smem8,
                                              int lane, uint32_t (&desc)[8]) {
    constexpr int IT = ROWS * COLS / 32;
    const int stride = STRIDE ? STRIDE : stride_rt;
    int v0[IT], cen[4];
    patch_issue<ROWS, COLS, STRIDE>(img, origin, stride_rt, 0, lane, v0, cen);
    const int vx = lane < ROWS ? (int)img[origin + (unsigned)(lane * stride + COLS)] : 0;  // column COLS: in bounds, x + w < cols (:388)
    const Band bd = band_of(cen, thr);
    uint32_t m0[IT], m1[IT];
    patch_words<IT>(v0, bd, m0);
    const unsigned cx = __ballot_sync(0xffffffffu, lane < ROWS && (bd.low > vx || bd.high < vx));
    int f = 0;
#pragma unroll
    for (int it = 0; it < IT; it++) {
        if (COLS == 16) {
            m1[it] = ((m0[it] >> 1) & 0x7fff7fffu) | (((cx >> (2 * it)) & 1u) << 15) | (((cx >> (2 * it + 1)) & 1u) << 31);
        } else {
            m1[it] = (m0[it] >> 1) & 0x7f7f7f7fu;
#pragma unroll
            for (int k = 0; k < 4; k++) m1[it] |= ((cx >> (4 * it + k)) & 1u) << (8 * k + 7);
        }
        f += __popc(m1[it]);
    }
    // pre-check (:122-139): the running count only grows and is tested per row, so "reaches precheck at some row end" ==
    // "total >= precheck" (the uint8 counter cannot wrap before the break)
    constexpr int precheck = ROWS * COLS / 8;
    if (f < precheck) return false;
    // diagonal walk (:141-190). A block diagonal is the set of cells with constant c - r (direction 1, closed form of the
    // tables EXPRESS.h:20-38: d = c - r + ROWS-1) or constant c + r (direction 0, the same on the column-reversed row).
    // Lane r shifts its row left by ROWS-1-r so that diagonal d sits in bit d of every lane; one 32x32 bit transpose then
    // hands lane d the cells of diagonal d, and the win count is a popcount.
    __syncwarp();
#pragma unroll
    for (int it = 0; it < IT; it++)
        if (lane == it) smem8[it] = m0[it];
    __syncwarp();
    unsigned row = 0;
    if (lane < ROWS) row = COLS == 16 ? reinterpret_cast<const uint16_t *>(smem8)[lane] : reinterpret_cast<const uint8_t *>(smem8)[lane];
    constexpr int slices = ROWS + COLS - 1;
    constexpr int rounds = slices == 31 ? 8 : slices == 23 ? 6 : 4;  // roundf(slices * .25f)
    constexpr uint32_t valid = (1u << slices) - 1u;
    const int len = min(min(lane + 1, ROWS), min(COLS, slices - lane));
    bool ok = false;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        if (a == 1 && ok) break;  // warp-uniform: the first direction already passed (EXPRESS.h:186-189 returns there)
        const unsigned bits = a == 0 ? row : (__brev(row) >> (32 - COLS));
        const unsigned diag = transpose32(lane < ROWS ? bits << (ROWS - 1 - lane) : 0u, lane);
        const int win = __popc(diag);
        const bool winbit = lane < slices && win >= len - win;  // win >= loss (:171); "loss > win" is its complement (:179)
        const uint32_t wb = __ballot_sync(0xffffffffu, winbit) & valid;
        // sticky run counters (:169-184): wins reaches `rounds` iff `rounds` consecutive win diagonals exist; the
        // early break (:185) only fires when the verdict is already false.
        if (has_run(wb, rounds) && has_run(~wb & valid, rounds)) ok = true;
    }
    __syncwarp();
    if (ok) desc_layout<ROWS, COLS>(m1, desc);
    return ok;
}

constexpr int BIRTH_KPW = 8;  // PIPE: kps entries a warp takes per round (finer than 32: the unclaimed blocks are spread unevenly)

template <int PITCH, bool PIPE>
__global__ void __launch_bounds__(CAND_THREADS, 32 / MOVFE_CAND_WARPS)  // 64 registers per thread
birth_kernel(ExtParams p, const movfe_rect *__restrict__ kps, const int32_t *__restrict__ nkps,
             const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, const int32_t *__restrict__ claim,
             uint8_t *__restrict__ birth_flag, uint32_t *__restrict__ birth_desc) {
    __shared__ uint32_t scratch[CAND_WARPS][8];
    __shared__ int sm[CAND_WARPS][2][32];
    __shared__ __align__(16) uint8_t swin[PIPE ? CAND_WARPS : 1][PF_DEPTH][WIN_BYTES];  // PIPE: block windows in flight (cp.async)
    pdl_wait();  // cand_kernel wrote the claims
    pdl_trigger();
    const int s = p.s0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = nkps[s * p.n_in + p.fi];
    if (!(fflags[s * p.RING + p.gslot] & MOVFE_FRAME_P)) return;
    const uint8_t *img = grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.P * p.H);
    const movfe_rect *kp = kps + ((size_t)s * p.n_out + p.fi) * p.max_kps;
    constexpr int KPW = PIPE ? BIRTH_KPW : 32;
    for (int c = blockIdx.x * CAND_WARPS + warp; c * KPW < n; c += gridDim.x * CAND_WARPS) {
        const int i = c * KPW + lane;
        // thread level: which blocks are unclaimed and inside the image (:381,:388)
        bool job = false;
        int2 r = make_int2(0, 0);
        if (lane < KPW && i < n) {
            r = __ldg(reinterpret_cast<const int2 *>(kp + i));  // x | y << 16, w | h << 16
            const int x = (int16_t)(r.x & 0xffff), y = r.x >> 16, w = (int16_t)(r.y & 0xffff), h = r.y >> 16;
            const bool claimed = claim[(size_t)s * p.max_kps + i] != 0x7fffffff;  // lbFound[i]
            job = !claimed && rect_in_bounds(x, y, w, h, p.W, p.H);
            birth_flag[(size_t)s * p.max_kps + i] = 0;
        }
        sm[warp][0][lane] = r.x;
        sm[warp][1][lane] = r.y;
        __syncwarp();
        unsigned todo = __ballot_sync(0xffffffffu, job);
        unsigned odd = 0;  // blocks of none of the four H.264 shapes: second loop (the generic mask indexes its array dynamically)
        const int stride = PITCH ? PITCH : p.P;
        auto publish = [&](int t, const uint32_t (&d)[8]) {
            if (lane == 0) {
                const size_t o = (size_t)s * p.max_kps + (c * KPW + t);
                birth_flag[o] = 1;
                uint4 *bd4 = reinterpret_cast<uint4 *>(birth_desc + o * 8);
                bd4[0] = make_uint4(d[0], d[1], d[2], d[3]);
                bd4[1] = make_uint4(d[4], d[5], d[6], d[7]);
            }
        };
        if (PIPE) {
            // the round's standard-shape blocks, PF_DEPTH windows in flight (the same staging as cand_kernel)
            unsigned std_m = 0;
            {
                const int w = (int16_t)(r.y & 0xffff), h = r.y >> 16;
                const bool stdsh = (w == 16 || w == 8) && (h == 16 || h == 8);
                std_m = __ballot_sync(0xffffffffu, job && stdsh);
                odd = todo & ~std_m;
            }
            auto issue = [&](int t, int slot) {
                const int rx = sm[warp][0][t];
                const int x = (int16_t)(rx & 0xffff), y = rx >> 16;
                window_issue(img, stride, x & ~15, y, swin[warp][slot], lane);
            };
            unsigned pend = std_m;  // blocks whose window has not been requested yet
#pragma unroll
            for (int k = 0; k < PF_DEPTH; k++) {
                if (pend) {
                    issue(__ffs(pend) - 1, k);
                    pend &= pend - 1;
                }
                cp_async_commit();
            }
            int k = 0;
            for (unsigned run = std_m; run; run &= run - 1, k++) {
                const int t = __ffs(run) - 1;
                cp_async_wait<PF_DEPTH - 1>();
                __syncwarp();
                const int rx = sm[warp][0][t], ry = sm[warp][1][t];
                const int x = (int16_t)(rx & 0xffff), w = (int16_t)(ry & 0xffff), h = ry >> 16;
                const uint8_t *win = swin[warp][k % PF_DEPTH];
                const unsigned xo = (unsigned)(x & 15);
                uint32_t d[8];
                bool pass;
                if (w == 16 && h == 16) pass = express_birth<16, 16, WIN_ROW>(win, xo, WIN_ROW, p.thr, scratch[warp], lane, d);
                else if (w == 8 && h == 8) pass = express_birth<8, 8, WIN_ROW>(win, xo, WIN_ROW, p.thr, scratch[warp], lane, d);
                else if (w == 8 && h == 16) pass = express_birth<16, 8, WIN_ROW>(win, xo, WIN_ROW, p.thr, scratch[warp], lane, d);
                else pass = express_birth<8, 16, WIN_ROW>(win, xo, WIN_ROW, p.thr, scratch[warp], lane, d);
                if (pass) publish(t, d);
                __syncwarp();  // every lane has read the window before its buffer is refilled
                if (pend) {
                    issue(__ffs(pend) - 1, k % PF_DEPTH);
                    pend &= pend - 1;
                }
                cp_async_commit();
            }
            cp_async_wait<0>();
            todo = 0;
        }
        while (!PIPE && todo) {
            const int t = __ffs(todo) - 1;
            todo &= todo - 1;
            const int rx = sm[warp][0][t], ry = sm[warp][1][t];
            const int x = (int16_t)(rx & 0xffff), y = rx >> 16, w = (int16_t)(ry & 0xffff), h = ry >> 16;
            if (!((w == 16 || w == 8) && (h == 16 || h == 8))) {  // warp-uniform
                odd |= 1u << t;
                continue;
            }
            const unsigned origin = (unsigned)(y * stride + x);
            uint32_t d[8];
            bool pass;
            if (w == 16 && h == 16) pass = express_birth<16, 16, PITCH>(img, origin, p.P, p.thr, scratch[warp], lane, d);
            else if (w == 8 && h == 8) pass = express_birth<8, 8, PITCH>(img, origin, p.P, p.thr, scratch[warp], lane, d);
            else if (w == 8 && h == 16) pass = express_birth<16, 8, PITCH>(img, origin, p.P, p.thr, scratch[warp], lane, d);
            else pass = express_birth<8, 16, PITCH>(img, origin, p.P, p.thr, scratch[warp], lane, d);
            if (pass) publish(t, d);
        }
        while (odd) {
            const int t = __ffs(odd) - 1;
            odd &= odd - 1;
            const int rx = sm[warp][0][t], ry = sm[warp][1][t];
            const int x = (int16_t)(rx & 0xffff), y = rx >> 16, w = (int16_t)(ry & 0xffff), h = ry >> 16;
            const uint8_t *roi = img + (unsigned)(y * stride + x);
            uint32_t d[8];
            const bool pass = express_test(roi, stride, h, w, p.thr, scratch[warp], lane);  // :391
            if (pass) {
                express_mask(roi, stride, h, w, express_band(roi, stride, h, w, p.thr), 1, false, d, lane);
                publish(t, d);
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------- birth_lane_kernel -----
// compute_express + descriptor of the unclaimed in-bounds kps blocks at THREAD level (express_lane.cuh: block_express), a lane
// per block, 32 blocks per warp step. A warp scans its share of the kps list 32 entries at a time and gathers the jobs in a list;
// whenever 32 have come together their windows are staged and every lane evaluates one block on its own: no ballots, no
// transposes, the diagonal walk is a carry-save sum of the block's shifted rows.
constexpr int BL_BATCH = 32;
constexpr size_t BL_SMEM = (size_t)CAND_WARPS * BL_BATCH * xl::WIN_STRIDE * sizeof(uint32_t);

#ifndef BIRTH_LANE_MINB
#define BIRTH_LANE_MINB 4
#endif
template <int PITCH>
__global__ void __launch_bounds__(CAND_THREADS, BIRTH_LANE_MINB)
birth_lane_kernel(ExtParams p, const movfe_rect *__restrict__ kps, const int32_t *__restrict__ nkps,
                  const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, const int32_t *__restrict__ claim,
                  uint8_t *__restrict__ birth_flag, uint32_t *__restrict__ birth_desc) {
    extern __shared__ __align__(16) uint32_t swin_all[];  // [CAND_WARPS][BL_BATCH * xl::WIN_STRIDE]: 49 KB, above the static limit
    uint32_t (*swin)[BL_BATCH * xl::WIN_STRIDE] = reinterpret_cast<uint32_t (*)[BL_BATCH * xl::WIN_STRIDE]>(swin_all);
    __shared__ uint32_t sorg[CAND_WARPS][2 * BL_BATCH];  // job list: window origin in the grey plane
    __shared__ uint32_t sinf[CAND_WARPS][2 * BL_BATCH];  //           kps index | (x & 7) << 20 | (w == 16) << 23 | (h == 16) << 24
    __shared__ uint32_t scratch[CAND_WARPS][8];
    __shared__ int sm[CAND_WARPS][2][32];
    pdl_wait();  // cand_kernel wrote the claims
    pdl_trigger();
    const int s = p.s0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = nkps[s * p.n_in + p.fi];
    if (!(fflags[s * p.RING + p.gslot] & MOVFE_FRAME_P)) return;
    const uint8_t *img = grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.P * p.H);
    const movfe_rect *kp = kps + ((size_t)s * p.n_out + p.fi) * p.max_kps;
    const int stride = PITCH ? PITCH : p.P;
    const StageLane stl = stage_lane(stride, lane);
    const unsigned lt = lanemask_lt();
    int n_jobs = 0;  // warp-uniform: entries of the job list

    auto run_step = [&](int nb) {  // evaluates list entries [0, nb), nb <= BL_BATCH
        stage_windows<BL_BATCH>(img, stl, sorg[warp], nb, swin[warp]);
        const bool on = lane < nb;
        const uint32_t inf = on ? sinf[warp][lane] : ((1u << 23) | (1u << 24));  // an idle lane computes on a 16x16 dummy
        const int cols = (inf >> 23) & 1u ? 16 : 8, rows = (inf >> 24) & 1u ? 16 : 8;
        cp_async_wait<0>();
        __syncwarp();
        uint32_t d[8];
        const bool pass = xl::block_express(swin[warp] + lane * xl::WIN_STRIDE, (int)((inf >> 20) & 7u), rows, cols, p.thr, d);  // :391
        if (on && pass) {
            const size_t o = (size_t)s * p.max_kps + (inf & 0xfffffu);
            birth_flag[o] = 1;
            uint4 *bd4 = reinterpret_cast<uint4 *>(birth_desc + o * 8);
            bd4[0] = make_uint4(d[0], d[1], d[2], d[3]);
            bd4[1] = make_uint4(d[4], d[5], d[6], d[7]);
        }
        __syncwarp();  // every lane has read its window before the slots are refilled
    };

    for (int c = blockIdx.x * CAND_WARPS + warp; c * 32 < n; c += gridDim.x * CAND_WARPS) {
        const int i = c * 32 + lane;
        // thread level: which blocks are unclaimed and inside the image (:381,:388)
        bool job = false;
        int2 r = make_int2(0, 0);
        if (i < n) {
            r = __ldg(reinterpret_cast<const int2 *>(kp + i));  // x | y << 16, w | h << 16
            const int x = (int16_t)(r.x & 0xffff), y = r.x >> 16, w = (int16_t)(r.y & 0xffff), h = r.y >> 16;
            const bool claimed = claim[(size_t)s * p.max_kps + i] != 0x7fffffff;  // lbFound[i]
            job = !claimed && rect_in_bounds(x, y, w, h, p.W, p.H);
            birth_flag[(size_t)s * p.max_kps + i] = 0;
        }
        const int x = (int16_t)(r.x & 0xffff), y = r.x >> 16, w = (int16_t)(r.y & 0xffff), h = r.y >> 16;
        const bool stdsh = (w == 16 || w == 8) && (h == 16 || h == 8);
        const unsigned std_m = __ballot_sync(0xffffffffu, job && stdsh);
        unsigned odd = __ballot_sync(0xffffffffu, job && !stdsh);
        if (job && stdsh) {
            const int pos = n_jobs + __popc(std_m & lt);
            sorg[warp][pos] = (uint32_t)(y * stride + (x & ~7));
            sinf[warp][pos] = (uint32_t)i | ((uint32_t)(x & 7) << 20) | (w == 16 ? 1u << 23 : 0u) | (h == 16 ? 1u << 24 : 0u);
        }
        n_jobs += __popc(std_m);
        __syncwarp();
        if (n_jobs >= BL_BATCH) {  // warp-uniform
            run_step(BL_BATCH);
            n_jobs -= BL_BATCH;
            const uint32_t o = lane < n_jobs ? sorg[warp][BL_BATCH + lane] : 0u, f = lane < n_jobs ? sinf[warp][BL_BATCH + lane] : 0u;
            __syncwarp();
            if (lane < n_jobs) {
                sorg[warp][lane] = o;
                sinf[warp][lane] = f;
            }
            __syncwarp();
        }
        if (odd) {  // blocks of none of the four H.264 shapes: the warp-level generic test
            sm[warp][0][lane] = r.x;
            sm[warp][1][lane] = r.y;
            __syncwarp();
            while (odd) {
                const int t = __ffs(odd) - 1;
                odd &= odd - 1;
                const int rx = sm[warp][0][t], ry = sm[warp][1][t];
                const int ox = (int16_t)(rx & 0xffff), oy = rx >> 16, ow = (int16_t)(ry & 0xffff), oh = ry >> 16;
                const uint8_t *roi = img + (unsigned)(oy * stride + ox);
                uint32_t d[8];
                const bool pass = express_test(roi, stride, oh, ow, p.thr, scratch[warp], lane);  // :391
                if (pass) {
                    express_mask(roi, stride, oh, ow, express_band(roi, stride, oh, ow, p.thr), 1, false, d, lane);
                    if (lane == 0) {
                        const size_t o = (size_t)s * p.max_kps + (c * 32 + t);
                        birth_flag[o] = 1;
                        uint4 *bd4 = reinterpret_cast<uint4 *>(birth_desc + o * 8);
                        bd4[0] = make_uint4(d[0], d[1], d[2], d[3]);
                        bd4[1] = make_uint4(d[4], d[5], d[6], d[7]);
                    }
                }
            }
            __syncwarp();
        }
    }
    if (n_jobs > 0) run_step(n_jobs);
}

// --------------------------------------------------------------------------------------- finalize_kernel -----
__device__ __forceinline__ int block_excl_scan(int v, int *wsum, int &total) {
    // exclusive scan of one int per thread over the CTA (FIN_THREADS); wsum: FIN_WARPS ints of shared memory
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < FIN_WARPS ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        if (lane < FIN_WARPS) wsum[lane] = w;  // inclusive
    }
    __syncthreads();
    total = wsum[FIN_WARPS - 1];
    const int before = warp ? wsum[warp - 1] : 0;
    __syncthreads();
    return before + x - v;
}

__device__ __forceinline__ int popc256(const uint32_t d[8]) {
    int c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c += __popc(d[i]);
    return c;
}

// Sort key of table entry `idx`: the reference orders prev->mvVF by age descending, then descriptor popcount descending
// (MOVExtractor.cc:249-252), canonicalised to a stable sort — ascending unique 64-bit keys.
__device__ __forceinline__ unsigned long long sort_key(int age, int popc, int idx) {
    const uint32_t a = 0x7fffffffu - (uint32_t)max(age, 0);
    return ((unsigned long long)a << 32) | ((unsigned long long)(256 - popc) << 16) | (unsigned)idx;
}

constexpr int SORT_MIN_N = 256;  // keys are padded to at least this many (one full warp at E = 8)

// Bitonic sort (ascending) of N = 2^m >= SORT_MIN_N unique 64-bit keys in shared memory by the whole CTA. Thread t owns
// the E consecutive keys [tE, tE+E): compare-exchange steps with partner distance j < E stay in registers, j < 32E go
// through warp shuffles, only the rest (15 of 78 steps at N = 4096) goes through shared memory and a barrier.
template <int E>
__device__ void bitonic_sort(unsigned long long *keys, int N) {
    const int t = threadIdx.x, lane = t & 31;
    const bool act = t * E < N;  // whole warps (N >= 32E)
    unsigned long long k[E];
    if (act) {
#pragma unroll
        for (int e = 0; e < E; e++) k[e] = keys[t * E + e];
    }
    for (int kk = 2; kk <= N; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j < E) {
                if (act) {
#pragma unroll
                    for (int jj = E >> 1; jj > 0; jj >>= 1) {
                        if (j == jj) {
#pragma unroll
                            for (int e = 0; e < E; e++) {
                                if ((e & jj) == 0) {
                                    const bool up = ((t * E + e) & kk) == 0;
                                    const unsigned long long a = k[e], b = k[e | jj];
                                    if ((a > b) == up) {
                                        k[e] = b;
                                        k[e | jj] = a;
                                    }
                                }
                            }
                        }
                    }
                }
            } else if (j < 32 * E) {
                if (act) {
                    const int lj = j / E;
                    const bool keep_min = ((lane & lj) == 0) == (((t * E) & kk) == 0);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[e], lj);
                        k[e] = keep_min ? (k[e] < o ? k[e] : o) : (k[e] > o ? k[e] : o);
                    }
                }
            } else {
                if (act) {
#pragma unroll
                    for (int e = 0; e < E; e++) keys[t * E + e] = k[e];
                }
                __syncthreads();
                if (act) {
                    const int tj = j / E;
                    const bool keep_min = ((t & tj) == 0) == (((t * E) & kk) == 0);
                    const unsigned long long *o = keys + (t ^ tj) * E;
#pragma unroll
                    for (int e = 0; e < E; e++) k[e] = keep_min ? (k[e] < o[e] ? k[e] : o[e]) : (k[e] > o[e] ? k[e] : o[e]);
                }
                __syncthreads();
            }
        }
    }
    if (act) {
#pragma unroll
        for (int e = 0; e < E; e++) keys[t * E + e] = k[e];
    }
    __syncthreads();
}

// Sorts keys[0..n) (already filled; n <= capacity of the shared array) and writes the permutation.
__device__ void sort_keys(unsigned long long *keys, int n, uint16_t *__restrict__ order_out) {
    int N = SORT_MIN_N;
    while (N < n) N <<= 1;
    for (int i = n + threadIdx.x; i < N; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();
    if (N <= 4 * FIN_THREADS) bitonic_sort<4>(keys, N);
    else if (N <= 8 * FIN_THREADS) bitonic_sort<8>(keys, N);
    else bitonic_sort<16>(keys, N);
    for (int i = threadIdx.x; i < n; i += blockDim.x) order_out[i] = (uint16_t)(keys[i] & 0xffffu);
    __syncthreads();
}

// The table finalize builds is already ordered by age (survivors keep the previous order with age + 1, births and
// lattice entries have age 0 and come last), so the reference's (age desc, popcount desc) order only permutes entries
// INSIDE runs of equal age. Each run is ordered by one warp with a stable counting sort on the 257 popcount values -
// about a microsecond for a 4000-entry table where the full bitonic network takes tens. pk = 256 - popcount.
// Returns false (nothing written) when the ages are not non-increasing; the caller then takes the general sort.
constexpr int HIST_STRIDE = 264;

__device__ bool sort_runs(const int *age, const uint16_t *pk, int n, uint16_t *run_start, int *hist, int *wsum,
                          uint16_t *__restrict__ order_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    // run heads (ordered compaction) and the monotonicity check
    int n_runs = 0;
    bool bad = false;
    for (int base = 0; base < n; base += FIN_THREADS * 8) {
        const int i0 = base + threadIdx.x * 8;
        unsigned hm = 0;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int i = i0 + e;
            if (i < n) {
                const int a = age[i];
                if (i == 0) hm |= 1u << e;
                else {
                    const int ap = age[i - 1];
                    if (a != ap) hm |= 1u << e;
                    if (a > ap) bad = true;
                }
            }
        }
        int tot;
        int r = n_runs + block_excl_scan(__popc(hm), wsum, tot);
#pragma unroll
        for (int e = 0; e < 8; e++)
            if ((hm >> e) & 1u) run_start[r++] = (uint16_t)(i0 + e);
        n_runs += tot;
    }
    if (__syncthreads_or(bad)) return false;
    int *h = hist + warp * HIST_STRIDE;
    for (int r = warp; r < n_runs; r += FIN_WARPS) {
        const int s = run_start[r], e = r + 1 < n_runs ? run_start[r + 1] : n, L = e - s;
        if (L == 1) {
            if (lane == 0) order_out[s] = (uint16_t)s;
        } else if (L <= 32) {
            const int p = lane < L ? pk[s + lane] : 0x7fff;
            int cnt = 0;
            for (int k = 0; k < L; k++) {
                const int o = __shfl_sync(0xffffffffu, p, k);
                cnt += (o < p) || (o == p && k < lane);
            }
            if (lane < L) order_out[s + cnt] = (uint16_t)(s + lane);
        } else {
            for (int b = lane; b < 257; b += 32) h[b] = 0;
            __syncwarp();
            for (int i = s + lane; i < e; i += 32) atomicAdd(&h[pk[i]], 1);
            __syncwarp();
            int carry = 0;
            for (int b0 = 0; b0 < 257; b0 += 32) {  // exclusive prefix, ascending pk = descending popcount
                const int b = b0 + lane;
                const int v = b < 257 ? h[b] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                if (b < 257) h[b] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            for (int c0 = s; c0 < e; c0 += 32) {  // stable placement, 32 consecutive entries at a time
                const int i = c0 + lane;
                const bool valid = i < e;
                const int p = valid ? pk[i] : 1000 + lane;
                const unsigned peers = __match_any_sync(0xffffffffu, p);
                const int rk = __popc(peers & lt);
                const int bs = valid ? h[p] : 0;
                __syncwarp();
                if (valid && rk == 0) h[p] = bs + __popc(peers);
                __syncwarp();
                if (valid) order_out[s + bs + rk] = (uint16_t)i;
            }
        }
    }
    __syncthreads();
    return true;
}

// keys of entries [from, n) from the table in global memory (entries written outside the fused copy paths)
__device__ void fill_keys(const movfe_track *__restrict__ tab, int from, int n, unsigned long long *keys) {
    for (int i = from + threadIdx.x; i < n; i += blockDim.x) {
        const movfe_track &t = tab[i];
        keys[i] = sort_key(t.age, popc256(t.desc), i);
    }
    __syncthreads();
}

// 16-px lattice walk shared by the coverage back-fill (:418-451) and the I-frame seeding (:123-157).
__device__ void lattice_pass(const ExtParams &p, const uint8_t *__restrict__ img, const int4 *__restrict__ g, const TileCells &tq, bool need_uncovered,
                             uint32_t track_flags, movfe_track *__restrict__ cur, int &n_out, int &id, uint32_t (*scratch)[8],
                             int *lat_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = (p.W - 16 + 15) / 16, gh = (p.H - 16 + 15) / 16;  // x = 8,24,.. < W-8
    const int nb = gw * gh;
    for (int base = 0; base < nb; base += FIN_WARPS) {
        const int b = base + warp;
        bool pass = false;
        uint32_t d[8];
        int x = 0, y = 0;
        if (b < nb) {
            y = 8 + 16 * (b / gw);
            x = 8 + 16 * (b % gw);
            if (rect_in_bounds(x - 8, y - 8, 16, 16, p.W, p.H)) {
                if (express_birth<16, 16, 0>(img, (unsigned)((y - 8) * p.P + (x - 8)), p.P, p.thr, scratch[warp], lane, d) &&
                    !(need_uncovered && (p.fused ? resolve_slots(tq, x, y).x : __ldg(&g[(size_t)y * p.W + x]).x) >= 0))
                    pass = true;
            }
        }
        if (lane == 0) lat_flag[warp] = pass;
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < FIN_WARPS; w++) {
            before += w < warp ? lat_flag[w] : 0;
            tot += lat_flag[w];
        }
        if (pass && n_out + before < p.maxT && lane == 0) {
            movfe_track t;
            t.pt_x = (float)x;
            t.pt_y = (float)y;
            t.mb = {(int16_t)(x - 8), (int16_t)(y - 8), 16, 16};
            t.track_id = id + before + 1;
            t.age = 0;
            t.q_indx = -1;
            t.flags = track_flags;
#pragma unroll
            for (int k = 0; k < 8; k++) t.desc[k] = d[k];
            cur[n_out + before] = t;
        }
        n_out += tot;
        id += tot;
        __syncthreads();
    }
}

#ifndef MOVFE_FIN_IPT
#define MOVFE_FIN_IPT 8
#endif
constexpr int FIN_IPT = MOVFE_FIN_IPT;  // consecutive entries a thread owns per round of the ordered compactions

__global__ void __launch_bounds__(FIN_THREADS, 1024 / FIN_THREADS)
finalize_kernel(ExtParams p, movfe_track *__restrict__ tracks, int32_t *__restrict__ ntracks,
                int32_t *__restrict__ cur_id, uint16_t *__restrict__ order, const movfe_track *__restrict__ stage,
                const int2 *__restrict__ cinfo, int32_t *__restrict__ claim, const movfe_rect *__restrict__ kps,
                const int32_t *__restrict__ nkps, const double *__restrict__ cov, const uint8_t *__restrict__ birth_flag,
                const uint32_t *__restrict__ birth_desc, SlotSource src,
                const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, LkBuf lk) {
    extern __shared__ unsigned long long keys[];  // general sort: keys[N]; run sort: age[maxT] pk[maxT] run_start[maxT] hist[32][264]
    __shared__ int wsum[FIN_WARPS];
    __shared__ uint32_t scratch[FIN_WARPS][8];
    __shared__ int lat_flag[FIN_WARPS];
    int *s_age = reinterpret_cast<int *>(keys);
    uint16_t *s_pk = reinterpret_cast<uint16_t *>(s_age + p.maxT);
    uint16_t *s_run = s_pk + p.maxT;
    int *s_hist = reinterpret_cast<int *>(s_run + p.maxT);
    uint16_t *s_src = reinterpret_cast<uint16_t *>(s_hist + FIN_WARPS * HIST_STRIDE);  // [FIN_THREADS * FIN_IPT] ranks that survive a round
    pdl_wait();  // cand_kernel / birth_kernel wrote the staging tables
    pdl_trigger();
    const int s = p.s0 + blockIdx.x;
    movfe_track *cur = tracks + ((size_t)s * p.TSLOTS + p.tslot_cur) * p.maxT;
    const movfe_track *st = stage + (size_t)s * p.maxT;
    const int2 *ci = cinfo + (size_t)s * p.maxT;
    int32_t *cl = claim + (size_t)s * p.max_kps;  // lbFound (MOVExtractor.cc:253), one entry per kps of the frame
    const int n_prev = ntracks[s * p.TSLOTS + p.tslot_prev];
    const uint8_t ff = fflags[s * p.RING + p.gslot];
    const bool is_p = ff & MOVFE_FRAME_P;
    const int n_kps = nkps[s * p.n_in + p.fi];
    int id = cur_id[s * p.TSLOTS + p.tslot_prev];
    int n_out = 0;    // logical size of the new table (entries beyond maxT are dropped)
    int n_keyed = 0;  // entries whose sort key is already in shared memory
    const uint8_t *img = p.has_grey ? grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.P * p.H) : nullptr;
    const int4 *g = p.fused ? nullptr : src.grid + ((size_t)s * p.n_out + p.fi) * ((size_t)p.W * p.H);
    TileCells tq = {};
    if (p.fused) tq = frame_cells(p, src, s);
    const movfe_track *prev = tracks + ((size_t)s * p.TSLOTS + p.tslot_prev) * p.maxT;
    const int lk_n = p.use_lk ? lk.n[s] : -1;
    const uint8_t *lk_st = lk.status + (size_t)s * p.maxT;
    const float2 *lk_pt = lk.pts + (size_t)s * p.maxT;
    bool rekey_all = false;  // entries were written outside the fused copy paths ahead of keyed ones
    int n_dropped = 0;       // carried tracks that got no LK result

    if (is_p) {
        // lost relocalisation (:161-243): seeds that passed the host-side tests (:207-215) come first
        const int n_reloc = (p.use_lk && img) ? lk.n_reloc[s] : 0;
        if (n_reloc > 0) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const movfe_reloc_seed *rs = lk.reloc + (size_t)s * p.maxT;
            for (int base = 0; base < n_reloc; base += FIN_WARPS) {
                const int i = base + warp;
                bool pass = false;
                uint32_t d[8];
                movfe_reloc_seed sd = {0, 0, 0.f, 0.f};
                int mx = 0, my = 0;
                if (i < n_reloc) {
                    sd = rs[i];
                    mx = (int)__fsub_rn(sd.x, 8.f);  // :218 cv::Rect(float, ...) truncates
                    my = (int)__fsub_rn(sd.y, 8.f);
                    if (rect_in_bounds(mx, my, 16, 16, p.W, p.H)) {  // :219
                        const uint8_t *roi = img + (size_t)my * p.P + mx;
                        express_mask(roi, p.P, 16, 16, express_band(roi, p.P, 16, 16, p.thr), 1, false, d, lane);  // :221-223
                        pass = true;
                    }
                }
                if (lane == 0) lat_flag[warp] = pass;
                __syncthreads();
                int before = 0, tot = 0;
                for (int w = 0; w < FIN_WARPS; w++) {
                    before += w < warp ? lat_flag[w] : 0;
                    tot += lat_flag[w];
                }
                if (pass && n_out + before < p.maxT && lane == 0) {
                    movfe_track t;
                    t.pt_x = sd.x;
                    t.pt_y = sd.y;
                    t.mb = {(int16_t)mx, (int16_t)my, 16, 16};
                    t.track_id = sd.track_id;
                    t.age = 0;
                    t.q_indx = sd.q_indx;
                    t.flags = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) t.desc[k] = d[k];
                    cur[n_out + before] = t;
                }
                n_out += tot;
                __syncthreads();
            }
            rekey_all = true;
        }
        // survivors in sorted order (:254-334): a thread owns FIN_IPT consecutive ranks, so one block scan orders a round.
        // The scan only produces the list of surviving ranks (shared memory); the 64-byte records are then moved by ALL
        // threads, one record each per step, instead of by the few threads whose ranks survived, eight in a row.
        for (int base = 0; base < n_prev; base += FIN_THREADS * FIN_IPT) {
            const int i0 = base + threadIdx.x * FIN_IPT;
            int2 c[FIN_IPT];
#pragma unroll
            for (int e = 0; e < FIN_IPT; e++) c[e] = i0 + e < n_prev ? ci[i0 + e] : make_int2(-1, 0);
            unsigned accm = 0;
#pragma unroll
            for (int e = 0; e < FIN_IPT; e++) {
                if ((c[e].y & 3) == 3) {  // in bounds and through the descriptor gate
                    const bool mine = c[e].x < 0 || c[e].x >= p.max_kps || cl[c[e].x] == i0 + e;  // !lbFound at my turn
                    if (mine) accm |= 1u << e;
                }
            }
            int tot;
            int pos = block_excl_scan(__popc(accm), wsum, tot);  // barriers: every claim test is done
#pragma unroll
            for (int e = 0; e < FIN_IPT; e++) {
                if ((c[e].y & 1) && c[e].x >= 0 && c[e].x < p.max_kps) cl[c[e].x] = 0x7fffffff;  // claims are per frame
                if ((accm >> e) & 1u) s_src[pos++] = (uint16_t)(i0 + e - base);
            }
            __syncthreads();
            for (int k = threadIdx.x; k < tot; k += FIN_THREADS) {
                const int dpos = n_out + k;
                if (dpos < p.maxT) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(st + base + s_src[k]);
                    const uint4 r0 = src[0], r1 = src[1], r2 = src[2], r3 = src[3];
                    uint4 *dst = reinterpret_cast<uint4 *>(cur + dpos);
                    dst[0] = r0;
                    dst[1] = r1;
                    dst[2] = r2;
                    dst[3] = r3;
                    const int pc = __popc(r2.x) + __popc(r2.y) + __popc(r2.z) + __popc(r2.w) + __popc(r3.x) + __popc(r3.y) +
                                   __popc(r3.z) + __popc(r3.w);
                    s_age[dpos] = max((int)r1.y, 0);
                    s_pk[dpos] = (uint16_t)(256 - pc);
                }
            }
            n_out += tot;
            __syncthreads();  // s_src is reused by the next round / the births
        }
        // coverage tracks (:337-377): the i-th coverage track in sorted order owns LK result i; carried ones follow the
        // propagated survivors, keep block and descriptor, age + 1, coverage flag, qIndx = i
        {
            const uint16_t *ord = order + (size_t)s * p.maxT;
            int cov_seen = 0;
            for (int base = 0; base < n_prev; base += FIN_THREADS * FIN_IPT) {
                const int i0 = base + threadIdx.x * FIN_IPT;
                unsigned covm = 0;
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++)
                    if (i0 + e < n_prev && (ci[i0 + e].y & 4)) covm |= 1u << e;
                if (!__syncthreads_or(covm != 0)) continue;  // block-uniform: no coverage track in this round
                int tot;
                const int k0 = cov_seen + block_excl_scan(__popc(covm), wsum, tot);
                cov_seen += tot;
                if (lk_n < 0) continue;  // block-uniform: nothing installed, the tracks are dropped (counted below)
                unsigned okm = 0;
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++) {
                    if ((covm >> e) & 1u) {
                        const int k = k0 + __popc(covm & ((1u << e) - 1u));
                        if (k < lk_n && lk_st[k]) {
                            const float2 q = lk_pt[k];
                            if (!(q.x < 0.f || q.y < 0.f || q.x >= (float)p.W || q.y >= (float)p.H)) okm |= 1u << e;  // :354
                        }
                    }
                }
                int tot2;
                int pos = n_out + block_excl_scan(__popc(okm), wsum, tot2);
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++) {
                    if ((okm >> e) & 1u) {
                        if (pos < p.maxT) {
                            const int k = k0 + __popc(covm & ((1u << e) - 1u));
                            const uint4 *src = reinterpret_cast<const uint4 *>(prev + ord[i0 + e]);
                            const uint4 r0 = src[0], r1 = src[1], r2 = src[2], r3 = src[3];
                            const float2 q = lk_pt[k];
                            uint4 *dst = reinterpret_cast<uint4 *>(cur + pos);
                            dst[0] = make_uint4(__float_as_uint(q.x), __float_as_uint(q.y), r0.z, r0.w);
                            dst[1] = make_uint4(r1.x, r1.y + 1, (uint32_t)k, MOVFE_TRACK_COVERAGE);
                            dst[2] = r2;
                            dst[3] = r3;
                        }
                        pos++;
                    }
                }
                if (tot2) rekey_all = true;
                n_out += tot2;
            }
            if (lk_n < 0) n_dropped += cov_seen;
        }
        // births in kps order (:379-416)
        int mov_cnt = 0;
        if (img) {
            const movfe_rect *kp = kps + ((size_t)s * p.n_out + p.fi) * p.max_kps;
            const uint8_t *bf = birth_flag + (size_t)s * p.max_kps;
            for (int base = 0; base < n_kps; base += FIN_THREADS * FIN_IPT) {
                const int i0 = base + threadIdx.x * FIN_IPT;
                unsigned bm = 0;
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++)
                    if (i0 + e < n_kps && bf[i0 + e]) bm |= 1u << e;
                int tot;
                int r = block_excl_scan(__popc(bm), wsum, tot);
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++)
                    if ((bm >> e) & 1u) s_src[r++] = (uint16_t)(i0 + e - base);
                __syncthreads();
                for (int k = threadIdx.x; k < tot; k += FIN_THREADS) {  // one new track per thread and step
                    const int pos = n_out + k;
                    if (pos < p.maxT) {
                        const int i = base + s_src[k];
                        const movfe_rect mb = kp[i];
                        const uint4 *dp = reinterpret_cast<const uint4 *>(birth_desc + ((size_t)s * p.max_kps + i) * 8);
                        const uint4 d0 = dp[0], d1 = dp[1];
                        // (mb.br() + mb.tl()) * 0.5 on Point_<int>: saturate_cast<int>(double) rounds half to even (:385)
                        const float fx = (float)__double2int_rn((mb.x + mb.w + mb.x) * 0.5);
                        const float fy = (float)__double2int_rn((mb.y + mb.h + mb.y) * 0.5);
                        uint4 *dst = reinterpret_cast<uint4 *>(cur + pos);
                        dst[0] = make_uint4(__float_as_uint(fx), __float_as_uint(fy), (uint32_t)(uint16_t)mb.x | ((uint32_t)(uint16_t)mb.y << 16),
                                            (uint32_t)(uint16_t)mb.w | ((uint32_t)(uint16_t)mb.h << 16));
                        dst[1] = make_uint4((uint32_t)(id + k + 1), 0u, (uint32_t)-1, 0u);  // ++mCurrentId, age 0, qIndx -1
                        dst[2] = d0;
                        dst[3] = d1;
                        const int pc = __popc(d0.x) + __popc(d0.y) + __popc(d0.z) + __popc(d0.w) + __popc(d1.x) + __popc(d1.y) +
                                       __popc(d1.z) + __popc(d1.w);
                        s_age[pos] = 0;
                        s_pk[pos] = (uint16_t)(256 - pc);
                    }
                }
                n_out += tot;
                id += tot;
                mov_cnt += tot;
                __syncthreads();
            }
        }
        n_keyed = min(n_out, p.maxT);
        // coverage back-fill (:418-451)
        if (img && (cov[s * p.n_in + p.fi] < p.cov_thr || mov_cnt < 60))
            lattice_pass(p, img, g, tq, true, MOVFE_TRACK_COVERAGE, cur, n_out, id, scratch, lat_flag);
    } else if (n_prev == 0 && img) {
        // I frame without previous features: seeding on the 16-px lattice (:123-157)
        lattice_pass(p, img, g, tq, false, 0u, cur, n_out, id, scratch, lat_flag);
    } else if (n_prev > 0) {
        // I frame with previous features (:81-120): every track of the previous table, in TABLE order, is carried to its LK
        // position with block and descriptor kept, age + 1, qIndx = its index; no seeding happens
        if (lk_n >= 0) {
            for (int base = 0; base < n_prev; base += FIN_THREADS * FIN_IPT) {
                const int i0 = base + threadIdx.x * FIN_IPT;
                unsigned okm = 0;
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++) {
                    const int i = i0 + e;
                    if (i < n_prev && i < lk_n && lk_st[i]) {
                        const float2 q = lk_pt[i];
                        if (!(q.x < 0.f || q.y < 0.f || q.x >= (float)p.W || q.y >= (float)p.H)) okm |= 1u << e;  // :98
                    }
                }
                int tot;
                int pos = n_out + block_excl_scan(__popc(okm), wsum, tot);
#pragma unroll
                for (int e = 0; e < FIN_IPT; e++) {
                    if ((okm >> e) & 1u) {
                        if (pos < p.maxT) {
                            const int i = i0 + e;
                            const uint4 *src = reinterpret_cast<const uint4 *>(prev + i);
                            const uint4 r0 = src[0], r1 = src[1], r2 = src[2], r3 = src[3];
                            const float2 q = lk_pt[i];
                            uint4 *dst = reinterpret_cast<uint4 *>(cur + pos);
                            dst[0] = make_uint4(__float_as_uint(q.x), __float_as_uint(q.y), r0.z, r0.w);
                            dst[1] = make_uint4(r1.x, r1.y + 1, (uint32_t)i, 0u);
                            dst[2] = r2;
                            dst[3] = r3;
                        }
                        pos++;
                    }
                }
                n_out += tot;
            }
            rekey_all = true;
        } else {
            n_dropped += n_prev;
        }
    }
    if (rekey_all) n_keyed = 0;
    const int n_new = min(n_out, p.maxT);
    if (threadIdx.x == 0) {
        ntracks[s * p.TSLOTS + p.tslot_cur] = n_new;
        cur_id[s * p.TSLOTS + p.tslot_cur] = id;
        if (n_dropped) atomicAdd(lk.dropped, (unsigned long long)n_dropped);  // diagnostic counter, not on the data path
    }
    __syncthreads();  // cur[] and the shared age / popcount arrays are visible to the whole CTA
    for (int i = n_keyed + threadIdx.x; i < n_new; i += blockDim.x) {  // lattice entries were written outside the fused copies
        const movfe_track &t = cur[i];
        s_age[i] = max(t.age, 0);
        s_pk[i] = (uint16_t)(256 - popc256(t.desc));
    }
    __syncthreads();
    if (!sort_runs(s_age, s_pk, n_new, s_run, s_hist, wsum, order + (size_t)s * p.maxT)) {
        fill_keys(cur, 0, n_new, keys);  // ages not ordered (cannot happen for tables this kernel built): general sort
        sort_keys(keys, n_new, order + (size_t)s * p.maxT);
    }
}

// Sorts a table that was installed from the host (movfe_set_tracks).
__global__ void __launch_bounds__(FIN_THREADS, 1024 / FIN_THREADS)
sort_only_kernel(int maxT, int TSLOTS, int tslot, int stream, const movfe_track *__restrict__ tracks,
                 const int32_t *__restrict__ ntracks, uint16_t *__restrict__ order) {
    extern __shared__ unsigned long long keys[];
    const int s = stream;
    const int n = ntracks[s * TSLOTS + tslot];
    fill_keys(tracks + ((size_t)s * TSLOTS + tslot) * maxT, 0, n, keys);
    sort_keys(keys, n, order + (size_t)s * maxT);
}

__global__ void fill_i32(int32_t *p, size_t n, int32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct ExtScratch {
    movfe_track *stage;  // [S][maxT] moved tracks by sorted rank, written by cand_kernel for in-bounds tracks
    int2 *cinfo;         // [S][maxT] {hop's kps index, bit0 in bounds | bit1 through the descriptor gate}
    int32_t *claim;
    uint8_t *birth_flag;
    uint32_t *birth_desc;
    uint16_t *order;
    LkBuf lk;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

ExtScratch carve(const movfe_ctx *ctx, size_t *total) {
    const movfe_config &c = ctx->cfg;
    const size_t S = c.n_streams;
    uint8_t *base = (uint8_t *)ctx->d_ext_scratch;
    size_t off = 0;
    ExtScratch e;
    e.stage = (movfe_track *)(base + off);
    off += align256(S * c.max_tracks * sizeof(movfe_track));
    e.cinfo = (int2 *)(base + off);
    off += align256(S * c.max_tracks * sizeof(int2));
    e.claim = (int32_t *)(base + off);
    off += align256(S * (size_t)ctx->max_kps * sizeof(int32_t));
    e.birth_flag = (uint8_t *)(base + off);
    off += align256(S * (size_t)ctx->max_kps);
    e.birth_desc = (uint32_t *)(base + off);
    off += align256(S * (size_t)ctx->max_kps * 32);
    e.order = (uint16_t *)(base + off);
    off += align256(S * c.max_tracks * sizeof(uint16_t));
    e.lk.n = (int32_t *)(base + off);
    off += align256(S * sizeof(int32_t));
    e.lk.n_reloc = (int32_t *)(base + off);
    off += align256(S * sizeof(int32_t));
    e.lk.dropped = (unsigned long long *)(base + off);
    off += 256;
    e.lk.status = (uint8_t *)(base + off);
    off += align256(S * (size_t)c.max_tracks);
    e.lk.pts = (float2 *)(base + off);
    off += align256(S * (size_t)c.max_tracks * sizeof(float2));
    e.lk.reloc = (movfe_reloc_seed *)(base + off);
    off += align256(S * (size_t)c.max_tracks * sizeof(movfe_reloc_seed));
    if (total) *total = off;
    return e;
}

size_t sort_smem(int maxT) {
    int N = SORT_MIN_N;
    while (N < maxT) N <<= 1;
    const size_t general = (size_t)N * sizeof(unsigned long long);
    const size_t runs = (size_t)maxT * 8 + (size_t)FIN_WARPS * HIST_STRIDE * sizeof(int) + (size_t)FIN_THREADS * FIN_IPT * sizeof(uint16_t);
    return std::max(general, runs);
}

}  // namespace

size_t movfe_extract_scratch_bytes(const movfe_ctx *ctx) {
    size_t total = 0;
    carve(ctx, &total);
    return total;
}

// The LK hand-over buffers and the sorted order of the newest table, for movfe_lk_carry (lk.cu): the device-side producer of what
// movfe_set_lk_results installs from the host.
void movfe_lk_buffers(const movfe_ctx *ctx, movfe_lk_handover *out) {
    ExtScratch e = carve(ctx, nullptr);
    out->n = e.lk.n;
    out->status = e.lk.status;
    out->pts = reinterpret_cast<float *>(e.lk.pts);
    out->order = e.order;
}

static int tslot_of(const movfe_ctx *ctx, int64_t frame) {
    const int T = ctx->TSLOTS;
    return (int)(((frame % T) + T) % T);
}

int movfe_extract_init(movfe_ctx *ctx) {
    const movfe_config &c = ctx->cfg;
    if (c.max_tracks > MAX_TRACKS_CAP) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "max_tracks must be <= %d", MAX_TRACKS_CAP);
    ExtScratch e = carve(ctx, nullptr);
    const size_t n = (size_t)c.n_streams * ctx->max_kps;
    fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(e.claim, n, 0x7fffffff);
    MOVFE_CUDA(ctx, cudaMemsetAsync(e.order, 0, (size_t)c.n_streams * c.max_tracks * sizeof(uint16_t), ctx->stream));
    MOVFE_CUDA(ctx, cudaMemsetAsync(e.lk.n, 0xff, (size_t)c.n_streams * sizeof(int32_t), ctx->stream));  // -1: nothing installed
    MOVFE_CUDA(ctx, cudaMemsetAsync(e.lk.n_reloc, 0, (size_t)c.n_streams * sizeof(int32_t), ctx->stream));
    MOVFE_CUDA(ctx, cudaMemsetAsync(e.lk.dropped, 0, sizeof(unsigned long long), ctx->stream));
    // the attributes are per function and process-wide: set once to the device limit, never per launch (contexts on other
    // host threads launch the same kernels)
    int fin_limit = 0;
    MOVFE_CUDA(ctx, optin_dynamic_smem(finalize_kernel, ctx->smem_optin, &fin_limit));
    if ((int)sort_smem(c.max_tracks) > fin_limit)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "max_tracks=%d needs %zu bytes of shared memory, the device allows %d", c.max_tracks, sort_smem(c.max_tracks), fin_limit);
    MOVFE_CUDA(ctx, optin_dynamic_smem(sort_only_kernel, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(birth_lane_kernel<0>, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(cand_lane_kernel<0>, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(cand_lane_kernel<1024>, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(cand_lane_kernel<2048>, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(birth_lane_kernel<1024>, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(birth_lane_kernel<2048>, ctx->smem_optin));
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}

int movfe_extract_launch(movfe_ctx *ctx, int64_t first_frame, int n_frames) {
    const movfe_config &c = ctx->cfg;
    ExtScratch e = carve(ctx, nullptr);
    // frames [first, first+n) overwrite the tables of frames TSLOTS earlier: a pose launch still reading those must finish
    for (const auto &pl : ctx->pose_launches)
        if (pl.first >= 0 && pl.first <= first_frame + n_frames - 1 - ctx->TSLOTS)
            MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pl.done, 0));
    // the raster results of this window were produced on the raster stream
    RasterBuf &w = ctx->rb[ctx->rb_cur];
    MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, w.done, 0));
    {
    ProfScope prof(ctx, MOVFE_STAGE_EXTRACT);
    // fork: the other groups' streams start after everything enqueued so far on the primary stream
    const int G = ctx->n_groups;
    if (G > 1) {
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        for (int g = 1; g < G; g++) MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->ext_stream[g], ctx->ev_fork, 0));
    }
    const bool pdl = ctx->pdl_mode != 0;       // cand -> birth -> finalize edges
    const bool pdl_cand = ctx->pdl_mode == 1;  // finalize -> next frame's cand edge as well
    for (int k = 0; k < n_frames; k++) {  // which event announces frame k's table
        const int kend = pdl_cand ? std::min((k / ctx->ev_batch + 1) * ctx->ev_batch - 1, n_frames - 1) : k;
        ctx->ev_of_frame[(first_frame + k) % c.window_frames] = (int)((first_frame + kend) % c.window_frames);
    }
    for (int k = 0; k < n_frames; k++) {
        const int64_t a = first_frame + k;
        ExtParams p;
        p.S = c.n_streams;
        p.W = c.width;
        p.H = c.height;
        p.maxT = c.max_tracks;
        p.max_kps = ctx->max_kps;
        p.max_hops = ctx->max_hops;
        p.maxM = c.max_records_per_frame;
        p.n_out = w.nout;
        p.n_in = w.nin;
        p.RING = ctx->RING;
        p.TSLOTS = ctx->TSLOTS;
        p.fi = (int)(a - w.first);
        p.gslot = (int)(a % ctx->RING);
        p.tslot_prev = tslot_of(ctx, a - 1);
        p.tslot_cur = tslot_of(ctx, a);
        p.thr = c.express_threshold;
        p.has_grey = c.has_grey;
        p.use_lk = (k == 0 && ctx->lk_pending) ? 1 : 0;
        p.fused = ctx->fused ? 1 : 0;
        p.NT = ctx->NT;
        p.tiles = ctx->NT * ctx->NTR;
        const SlotSource src = {w.d_grid, w.d_tc_dim, w.d_tc_runs, w.d_tc_cells};
        p.P = ctx->grey_pitch;
        p.cov_thr = c.coverage_threshold;
        const bool batch_end = !pdl_cand || (k + 1) % ctx->ev_batch == 0 || k == n_frames - 1;
        for (int g = 0; g < G; g++) {
        const int s_lo = (int)((int64_t)c.n_streams * g / G), ns = (int)((int64_t)c.n_streams * (g + 1) / G) - s_lo;
        cudaStream_t gs = ctx->ext_stream[g];
        p.s0 = s_lo;
        // grid-stride over tracks / kps: enough CTAs to fill the chip, never one CTA per (mostly empty) capacity slot
        const int bps = ctx->cand_bps > 0 ? ctx->cand_bps : std::max(4, (64 * ctx->sm_count + c.n_streams * CAND_WARPS - 1) / (c.n_streams * CAND_WARPS));
        dim3 gc(std::min((c.max_tracks + CAND_THREADS - 1) / CAND_THREADS, bps), ns);  // a warp takes 32 tracks
        // thread-level descriptors (cand_lane_kernel / birth_lane_kernel) need an image and a threshold below 128
        const bool lane_mode = ctx->cand_lane && c.has_grey && c.express_threshold >= 0 && c.express_threshold <= xl::MAX_THR;
#define MOVFE_CAND(PITCH)                                                                                              \
    MOVFE_CUDA(ctx, launch_pdl(pdl_cand, lane_mode ? cand_lane_kernel<PITCH> : ctx->cand_pipe ? cand_kernel<PITCH, true> : cand_kernel<PITCH, false>, gc, dim3(CAND_THREADS), lane_mode ? LN_SMEM + (size_t)ctx->cand_pad_bytes : 0, gs, p, ctx->d_tracks, ctx->d_ntracks, e.order, \
                               src, w.d_hops, ctx->d_grey, ctx->d_fflags, e.stage, e.cinfo, e.claim, ctx->d_stats))
        switch (ctx->grey_pitch) {  // the usual pitches get compile-time row offsets
            case 1024: MOVFE_CAND(1024); break;
            case 2048: MOVFE_CAND(2048); break;
            default: MOVFE_CAND(0); break;
        }
#undef MOVFE_CAND
        int nl = 2;
        if (c.has_grey) {
            const int kpw = lane_mode ? 32 * ctx->birth_chunks : ctx->cand_pipe ? BIRTH_KPW : 32;
            dim3 gb(std::min((ctx->max_kps + kpw * CAND_WARPS - 1) / (kpw * CAND_WARPS), (ctx->cand_pipe && !lane_mode) ? 2 * bps : bps), ns);
#define MOVFE_BIRTH(PITCH)                                                                                             \
    MOVFE_CUDA(ctx, launch_pdl(pdl, lane_mode ? birth_lane_kernel<PITCH> : ctx->cand_pipe ? birth_kernel<PITCH, true> : birth_kernel<PITCH, false>, gb, dim3(CAND_THREADS), lane_mode ? BL_SMEM : 0, gs, p, w.d_kps, w.d_nkps, ctx->d_grey,  \
                               ctx->d_fflags, e.claim, e.birth_flag, e.birth_desc))
            switch (ctx->grey_pitch) {
                case 1024: MOVFE_BIRTH(1024); break;
                case 2048: MOVFE_BIRTH(2048); break;
                default: MOVFE_BIRTH(0); break;
            }
#undef MOVFE_BIRTH
            nl = 3;
        }
        MOVFE_CUDA(ctx, launch_pdl(pdl, finalize_kernel, dim3(ns), dim3(FIN_THREADS), sort_smem(c.max_tracks), gs, p, ctx->d_tracks,
                                   ctx->d_ntracks, ctx->d_cur_id, e.order, e.stage, e.cinfo, e.claim, w.d_kps, w.d_nkps, w.d_cov,
                                   e.birth_flag, e.birth_desc, src, ctx->d_grey, ctx->d_fflags, e.lk));
        prof.launches(nl);
        // the tables up to frame a are complete for this group: the pose stream may start on them while propagation goes on.
        // One event per ev_batch frames (and at the end of the call): an event record between two kernels breaks their
        // programmatic dependency.
        if (batch_end) MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_frame[(size_t)g * c.window_frames + a % c.window_frames], gs));
        }
        if (p.use_lk) {  // the results belonged to this frame only (every group has consumed them once the groups join)
            for (int g = 1; g < G; g++) {
                MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_join[g], ctx->ext_stream[g]));
                MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[g], 0));
            }
            MOVFE_CUDA(ctx, cudaMemsetAsync(e.lk.n, 0xff, (size_t)c.n_streams * sizeof(int32_t), ctx->stream));
            MOVFE_CUDA(ctx, cudaMemsetAsync(e.lk.n_reloc, 0, (size_t)c.n_streams * sizeof(int32_t), ctx->stream));
            ctx->lk_pending = false;
        }
    }
    // join: whatever follows on the primary stream sees every group's tables
    for (int g = 1; g < G; g++) {
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_join[g], ctx->ext_stream[g]));
        MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[g], 0));
    }
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
    // the next raster into this buffer and later pushes into the ring slots of these frames wait for this launch
    MOVFE_CUDA(ctx, cudaEventRecord(w.consumed, ctx->stream));
    w.consumed_valid = true;
    movfe_ctx::ExtLaunch &el = ctx->ext_launches[ctx->ext_launch_head];
    el.first = first_frame;
    el.n = n_frames;
    MOVFE_CUDA(ctx, cudaEventRecord(el.done, ctx->stream));
    ctx->ext_launch_head = (ctx->ext_launch_head + 1) % movfe_ctx::N_EXT_LAUNCHES;
    ctx->ext_launch_count++;
    return MOVFE_OK;
}

// ----------------------------------------------------------------------------------------------- C-ABI --------
extern "C" int movfe_set_tracks(movfe_ctx *ctx, int stream, const movfe_track *tracks, int n, int32_t current_id) {
    if (!ctx) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    if (stream < 0 || stream >= c.n_streams || n < 0 || (n > 0 && !tracks)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "set_tracks: bad argument");
    if (n > c.max_tracks) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "set_tracks: %d tracks, capacity %d", n, c.max_tracks);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    const int64_t next = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    const int ts = tslot_of(ctx, next - 1);
    const int T = ctx->TSLOTS;
    // a pose chain still running on the pose stream may be reading the table of frame next-1 (or the frame that shares
    // its slot): the copy below waits for every pose launch covering either
    for (const auto &pl : ctx->pose_launches)
        if (pl.first >= 0 && ((pl.first <= next - 1 && next - 1 < pl.first + pl.n) || (pl.first <= next - 1 - T && next - 1 - T < pl.first + pl.n)))
            MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pl.done, 0));
    if (n > 0)
        MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_tracks + ((size_t)stream * T + ts) * c.max_tracks, tracks, (size_t)n * sizeof(movfe_track),
                                        cudaMemcpyHostToDevice, ctx->stream));
    const int32_t nn = n;
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_ntracks + stream * T + ts, &nn, 4, cudaMemcpyHostToDevice, ctx->stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_cur_id + stream * T + ts, &current_id, 4, cudaMemcpyHostToDevice, ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // nn / current_id live on this stack frame
    ExtScratch e = carve(ctx, nullptr);
    sort_only_kernel<<<1, FIN_THREADS, sort_smem(c.max_tracks), ctx->stream>>>(c.max_tracks, T, ts, stream, ctx->d_tracks,
                                                                              ctx->d_ntracks, e.order);
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}

extern "C" int movfe_extract(movfe_ctx *ctx, int64_t first_frame, int n_frames) {
    if (!ctx) return MOVFE_E_INVALID;
    const int64_t next = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    if (first_frame != next)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract: frames must be consumed in order (expected %lld, got %lld)", (long long)next, (long long)first_frame);
    const RasterBuf &w = ctx->rb[ctx->rb_cur];
    if (n_frames < 1 || w.first < 0 || first_frame < w.first || first_frame + n_frames > w.first + w.nout)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract: frames [%lld,%lld) are not inside the last raster window", (long long)first_frame,
                   (long long)(first_frame + n_frames));
    // the grey planes and frame flags of these frames must still be in the ring (a later push may have overwritten them:
    // push(k), raster(k), push(k+1), push(k+2), extract(k) passes every other check)
    if (ctx->pushed - first_frame > ctx->RING)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract: frame %lld has left the ring (pushed=%lld, ring=%d)", (long long)first_frame,
                   (long long)ctx->pushed, ctx->RING);
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int rc = movfe_extract_launch(ctx, first_frame, n_frames);
    if (rc) return rc;
    ctx->ext_first = first_frame;
    ctx->ext_n = n_frames;
    return MOVFE_OK;
}

static int track_slot(movfe_ctx *ctx, int stream, int64_t frame, int *ts) {
    if (!ctx) return MOVFE_E_INVALID;
    if (stream < 0 || stream >= ctx->cfg.n_streams) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "stream %d out of range", stream);
    const int64_t next = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    if (frame >= next || frame < next - ctx->TSLOTS)  // TSLOTS = 2F+1 tables per stream are resident (two windows + the seed)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "track table of frame %lld is not resident", (long long)frame);
    *ts = tslot_of(ctx, frame);
    return MOVFE_OK;
}

extern "C" int movfe_track_count(movfe_ctx *ctx, int stream, int64_t frame, int32_t *n_tracks, int32_t *current_id) {
    int ts;
    int rc = track_slot(ctx, stream, frame, &ts);
    if (rc) return rc;
    const int T = ctx->TSLOTS;
    if (n_tracks) MOVFE_CUDA(ctx, cudaMemcpyAsync(n_tracks, ctx->d_ntracks + stream * T + ts, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (current_id) MOVFE_CUDA(ctx, cudaMemcpyAsync(current_id, ctx->d_cur_id + stream * T + ts, 4, cudaMemcpyDeviceToHost, ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MOVFE_OK;
}

extern "C" int movfe_download_tracks(movfe_ctx *ctx, int stream, int64_t frame, movfe_track *out, int capacity) {
    int ts, n = 0;
    int rc = track_slot(ctx, stream, frame, &ts);
    if (rc) return rc;
    rc = movfe_track_count(ctx, stream, frame, &n, nullptr);
    if (rc) return rc;
    if (n > capacity) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "download_tracks: %d tracks, capacity %d", n, capacity);
    const int T = ctx->TSLOTS;
    if (n > 0) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_tracks + ((size_t)stream * T + ts) * ctx->cfg.max_tracks, (size_t)n * sizeof(movfe_track),
                                        cudaMemcpyDeviceToHost, ctx->stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return n;
}

extern "C" int movfe_set_lk_results(movfe_ctx *ctx, int stream, const uint8_t *status, const float *pts_xy, int n,
                                    const movfe_reloc_seed *reloc, int n_reloc) {
    if (!ctx) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    if (stream < 0 || stream >= c.n_streams || n < -1 || n_reloc < 0 || (n > 0 && (!status || !pts_xy)) || (n_reloc > 0 && !reloc))
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "set_lk_results: bad argument");
    if (n > c.max_tracks || n_reloc > c.max_tracks)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "set_lk_results: %d results / %d seeds, capacity %d", n, n_reloc, c.max_tracks);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    ExtScratch e = carve(ctx, nullptr);
    cudaStream_t st = ctx->stream;
    if (n > 0) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(e.lk.status + (size_t)stream * c.max_tracks, status, (size_t)n, cudaMemcpyHostToDevice, st));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(e.lk.pts + (size_t)stream * c.max_tracks, pts_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
    }
    if (n_reloc > 0)
        MOVFE_CUDA(ctx, cudaMemcpyAsync(e.lk.reloc + (size_t)stream * c.max_tracks, reloc, (size_t)n_reloc * sizeof(movfe_reloc_seed),
                                        cudaMemcpyHostToDevice, st));
    const int32_t nn = n, nr = n_reloc;
    MOVFE_CUDA(ctx, cudaMemcpyAsync(e.lk.n + stream, &nn, 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(e.lk.n_reloc + stream, &nr, 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));  // the arrays are the caller's, nn / nr live on this stack frame
    ctx->lk_pending = true;
    return MOVFE_OK;
}

extern "C" int64_t movfe_dropped_lk_tracks(movfe_ctx *ctx) {
    if (!ctx) return -1;
    ExtScratch e = carve(ctx, nullptr);
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, e.lk.dropped, sizeof v, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return (int64_t)v;
}

// Single-shot MOVExtractor::operator() for callers that hold one frame's raster results on the host (the drop-in shim
// when only MOVExtractor is replaced; the parity tests of propagation in isolation). One-stream contexts only; every call
// is one new frame, so it must not be mixed with the batched push / raster / extract calls on the same context.
int movfe_grey_upload(movfe_ctx *ctx, const uint8_t *d_src, int slot);  // raster.cu

extern "C" int movfe_extract_frame(movfe_ctx *ctx, uint32_t frame_flags, const uint8_t *grey, int grey_stride, const int32_t *grid,
                                   const movfe_hop *hops, int n_hops, const movfe_rect *kps, int n_kps, double coverage_area,
                                   const movfe_track *prev, int n_prev, const uint8_t *lk_status, const float *lk_pts, int n_lk,
                                   const movfe_reloc_seed *reloc, int n_reloc, int32_t *current_id, movfe_track *out, int capacity) {
    if (!ctx) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    if (grey && grey_stride == 0) grey_stride = c.width;
    if (grey && grey_stride < c.width) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "extract_frame: grey_stride %d below the frame width %d", grey_stride, c.width);
    if (c.n_streams != 1) MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract_frame: the context must have exactly one stream");
    if (ctx->fused) MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract_frame: takes a slot grid from the host; the context was created with MOVFE_CFG_NO_GRID");
    if (!grid || !current_id || !out || n_hops < 0 || n_kps < 0 || n_prev < 0 || (n_hops && !hops) || (n_kps && !kps) || (n_prev && !prev))
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "extract_frame: bad argument");
    if ((c.has_grey != 0) != (grey != nullptr))
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "extract_frame: grey plane %s but the context was created with has_grey=%d", grey ? "given" : "missing", c.has_grey);
    if (n_hops > ctx->max_hops || n_kps > ctx->max_kps)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "extract_frame: %d hops / %d kps exceed the context's capacity (%d / %d)", n_hops, n_kps, ctx->max_hops, ctx->max_kps);
    const int64_t next = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    if (next != ctx->pushed) MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract_frame: mixed with the batched calls on this context");
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    int rc = movfe_set_tracks(ctx, 0, prev, n_prev, *current_id);
    if (rc) return rc;
    if (n_lk >= 0 || n_reloc > 0) {
        rc = movfe_set_lk_results(ctx, 0, lk_status, lk_pts, n_lk, reloc, n_reloc);
        if (rc) return rc;
    }
    const int64_t a = ctx->pushed;
    const int slot = (int)(a % ctx->RING);
    const size_t plane = (size_t)c.width * c.height;
    cudaStream_t st = ctx->stream;
    RasterBuf &w = ctx->rb[ctx->rb_cur];  // the caller's raster results stand in for a movfe_raster call
    MOVFE_CUDA(ctx, cudaMemcpyAsync(w.d_grid, grid, plane * sizeof(int4), cudaMemcpyHostToDevice, st));
    if (n_hops) MOVFE_CUDA(ctx, cudaMemcpyAsync(w.d_hops, hops, (size_t)n_hops * sizeof(movfe_hop), cudaMemcpyHostToDevice, st));
    if (n_kps) MOVFE_CUDA(ctx, cudaMemcpyAsync(w.d_kps, kps, (size_t)n_kps * sizeof(movfe_rect), cudaMemcpyHostToDevice, st));
    const int32_t nh = n_hops, nk = n_kps;
    const uint8_t ff = (uint8_t)frame_flags;
    MOVFE_CUDA(ctx, cudaMemcpyAsync(w.d_nhops, &nh, 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(w.d_nkps, &nk, 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(w.d_cov, &coverage_area, 8, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_fflags + slot, &ff, 1, cudaMemcpyHostToDevice, st));
    if (grey) {
        if (ctx->stage_bytes[0] < plane + 16) {
            MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
            if (ctx->d_stage[0]) cudaFree(ctx->d_stage[0]);
            ctx->d_stage[0] = nullptr;
            ctx->stage_bytes[0] = 0;
            MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_stage[0], plane + 4096));
            ctx->stage_bytes[0] = plane + 4096;
        }
        // rows of `grey_stride` bytes (cv::Mat::step / AVFrame::linesize) are packed by the copy itself
        MOVFE_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_stage[0], (size_t)c.width, grey, (size_t)grey_stride, (size_t)c.width, (size_t)c.height,
                                          cudaMemcpyHostToDevice, st));
        rc = movfe_grey_upload(ctx, (const uint8_t *)ctx->d_stage[0], slot);
        if (rc) return rc;
    }
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));  // nh / nk / ff live on this stack frame
    ctx->pushed = a + 1;
    w.first = a;
    w.nout = 1;
    w.nin = 1;
    rc = movfe_extract(ctx, a, 1);
    if (rc) return rc;
    int32_t n = 0;
    rc = movfe_track_count(ctx, 0, a, &n, current_id);
    if (rc) return rc;
    return movfe_download_tracks(ctx, 0, a, out, capacity);
}
