// extract.cu — reference-chained track propagation + EXPRESS descriptors.
// Replaces MOVExtractor::operator() (src/MOVExtractor.cc:63-455) and include/EXPRESS.h:79-192, batched over
// streams; frames of a stream are processed in order (frame f's table is the input of frame f+1).
// Compiled with -fmad=false (positions are binary32 sums that must round like the reference).
//
// Per frame, three launches:
//   cand_kernel      one warp per previous track (in the reference's sorted order): slot-grid lookup, up to four
//                    candidate hops scored by the Hamming distance of warp-ballot EXPRESS descriptors, move,
//                    bounds, descriptor gate; in-bounds tracks claim their hop's kps entry with an integer
//                    atomicMin on the sorted rank (order-independent result == the reference's first-come rule).
//   birth_kernel     one warp per candidate-keypoint block: unclaimed + in bounds + compute_express -> descriptor.
//   finalize_kernel  one CTA per stream: ordered compaction of the survivors, births appended in kps order with
//                    ids ++mCurrentId, optional coverage back-fill / I-frame seeding on the 16-px lattice, then
//                    the stable (age desc, popcount desc) order of the new table for the next frame (bitonic sort
//                    of unique 64-bit keys in shared memory).
// LK-carried features (cv::calcOpticalFlowPyrLK; MOVExtractor.cc:81-120,161-243,337-377) are host work: coverage
// tracks and I-frame carry-over are dropped here, exactly like the oracle with lk_status == NULL.
#include <algorithm>
#include <cstdio>

#include "common.cuh"

namespace {

constexpr int CAND_WARPS = 8;
constexpr int FIN_THREADS = 1024;
constexpr int FIN_WARPS = FIN_THREADS / 32;
constexpr int MAX_TRACKS_CAP = 8192;

// Per previous track, produced by cand_kernel (index = sorted rank).
struct __align__(16) Cand {
    float pt_x, pt_y;
    movfe_rect mb;
    int32_t d_indx;
    uint32_t flags;  // bit0: in bounds (may claim), bit1: passes the descriptor gate
    uint32_t pad[2];
    uint32_t desc[8];
};
static_assert(sizeof(Cand) == 64, "Cand is 64 bytes");

struct ExtParams {
    int S, W, H, maxT, max_kps, max_hops, maxM, n_out, n_in, RING, TSLOTS;
    int fi;          // raster-window slot of this frame
    int gslot;       // ring slot of this frame (grey, flags)
    int tslot_prev, tslot_cur;
    int thr;         // EXPRESS threshold
    int has_grey;
    double cov_thr;
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

// ------------------------------------------------------------------------------------------- cand_kernel -----
__device__ __forceinline__ void cand_one(const ExtParams &p, int s, int i, int lane, const movfe_track *__restrict__ tracks,
                                         const uint16_t *__restrict__ order, const int4 *__restrict__ grid,
                                         const movfe_hop *__restrict__ hops, const uint8_t *__restrict__ grey,
                                         Cand *__restrict__ cand, int32_t *__restrict__ claim);

__global__ void __launch_bounds__(CAND_WARPS * 32)
cand_kernel(ExtParams p, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
            const uint16_t *__restrict__ order, const int4 *__restrict__ grid, const movfe_hop *__restrict__ hops,
            const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, Cand *__restrict__ cand,
            int32_t *__restrict__ claim) {
    const int s = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_prev = ntracks[s * p.TSLOTS + p.tslot_prev];
    if (!(fflags[s * p.RING + p.gslot] & MOVFE_FRAME_P)) return;  // I frame: nothing is propagated
    for (int i = blockIdx.x * CAND_WARPS + warp; i < n_prev; i += gridDim.x * CAND_WARPS)  // i = sorted rank
        cand_one(p, s, i, lane, tracks, order, grid, hops, grey, cand, claim);
}

__device__ __forceinline__ void cand_one(const ExtParams &p, int s, int i, int lane, const movfe_track *__restrict__ tracks,
                                         const uint16_t *__restrict__ order, const int4 *__restrict__ grid,
                                         const movfe_hop *__restrict__ hops, const uint8_t *__restrict__ grey,
                                         Cand *__restrict__ cand, int32_t *__restrict__ claim) {
    const movfe_track *prev = tracks + ((size_t)s * p.TSLOTS + p.tslot_prev) * p.maxT;
    const movfe_track pvf = prev[order[(size_t)s * p.maxT + i]];
    Cand out;
    out.pt_x = 0.f;
    out.pt_y = 0.f;
    out.mb = pvf.mb;
    out.d_indx = -1;
    out.flags = 0;
    out.pad[0] = out.pad[1] = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) out.desc[k] = 0;
    Cand *dst = cand + (size_t)s * p.maxT + i;

    const int4 *g = grid + ((size_t)s * p.n_out + p.fi) * ((size_t)p.W * p.H);
    const movfe_hop *hp = hops + ((size_t)s * p.n_out + p.fi) * p.max_hops;
    const uint8_t *img = p.has_grey ? grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.W * p.H) : nullptr;

    bool alive = !(pvf.flags & MOVFE_TRACK_COVERAGE);  // :258-262 coverage tracks go to the host LK step
    int4 sl = make_int4(-1, -1, -1, -1);
    if (alive) {
        const int x = (int)pvf.pt_x, y = (int)pvf.pt_y;  // :264
        if (x < 0 || y < 0 || x >= p.W || y >= p.H) alive = false;  // unchecked .at<>() in the reference (UB)
        else sl = __ldg(&g[(size_t)y * p.W + x]);
    }
    if (alive && sl.x == -1) alive = false;  // :265-268
    if (!alive) {
        if (lane == 0) *dst = out;
        return;
    }
    const int mw = pvf.mb.w, mh = pvf.mb.h;
    const float hw = (float)(mw / 2), hh = (float)(mh / 2);
    int indx = sl.x;  // :270
    uint32_t best_desc[8];
    bool have_best = false;
    if (sl.y >= 0) {  // :272 (MV-only mode: every distance is 0, so the first in-bounds candidate wins)
        int bestDesc = 256;
        const int sj[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (sj[j] == -1) break;  // :277-278
            const movfe_hop mv = hp[sj[j]];
            const float px = __fadd_rn(pvf.pt_x, mv.mv_x), py = __fadd_rn(pvf.pt_y, mv.mv_y);  // :283
            const int mx = (int)__fsub_rn(px, hw), my = (int)__fsub_rn(py, hh);                // :284
            if (rect_in_bounds(mx, my, mw, mh, p.W, p.H)) {                                    // :286
                uint32_t d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int dist = 0;
                if (img) {
                    const uint8_t *roi = img + (size_t)my * p.W + mx;
                    const Band bd = express_band(roi, p.W, mh, mw, p.thr);
                    express_mask(roi, p.W, mh, mw, bd, 1, false, d, lane);
                    dist = hamming256(pvf.desc, d);
                }
                if (dist < bestDesc) {  // :292-296
                    bestDesc = dist;
                    indx = sj[j];
                    have_best = img != nullptr;
#pragma unroll
                    for (int k = 0; k < 8; k++) best_desc[k] = d[k];
                }
            }
        }
    }
    const movfe_hop mv = hp[indx];  // :301
    const float px = __fadd_rn(pvf.pt_x, mv.mv_x), py = __fadd_rn(pvf.pt_y, mv.mv_y);
    const int mx = (int)__fsub_rn(px, hw), my = (int)__fsub_rn(py, hh);
    out.pt_x = px;
    out.pt_y = py;
    out.mb.x = (int16_t)mx;
    out.mb.y = (int16_t)my;
    out.d_indx = mv.d_indx;
    if (rect_in_bounds(mx, my, mw, mh, p.W, p.H)) {  // :306 (the claim test itself happens in finalize)
        out.flags |= 1u;
        if (img) {
            uint32_t d[8];
            if (have_best) {  // indx is the candidate whose descriptor won the comparison: same rectangle, reuse it
#pragma unroll
                for (int k = 0; k < 8; k++) d[k] = best_desc[k];
            } else {  // single-candidate pixel: no descriptor was evaluated yet
                const uint8_t *roi = img + (size_t)my * p.W + mx;
                const Band bd = express_band(roi, p.W, mh, mw, p.thr);
                express_mask(roi, p.W, mh, mw, bd, 1, false, d, lane);
            }
            const int dist = hamming256(pvf.desc, d);  // :311-316
            if (dist <= 40) out.flags |= 2u;
#pragma unroll
            for (int k = 0; k < 8; k++) out.desc[k] = d[k];
        } else {
            out.flags |= 2u;  // MV-only mode: flat image, every distance is 0
        }
        if (lane == 0 && mv.d_indx >= 0 && mv.d_indx < p.maxM)
            atomicMin(&claim[(size_t)s * p.maxM + mv.d_indx], i);  // first-come in sorted order (:306-309)
    }
    if (lane == 0) *dst = out;
}

// ------------------------------------------------------------------------------------------ birth_kernel -----
__global__ void __launch_bounds__(CAND_WARPS * 32)
birth_kernel(ExtParams p, const movfe_rect *__restrict__ kps, const int32_t *__restrict__ nkps,
             const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags, const int32_t *__restrict__ claim,
             uint8_t *__restrict__ birth_flag, uint32_t *__restrict__ birth_desc) {
    __shared__ uint32_t scratch[CAND_WARPS][8];
    const int s = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = nkps[s * p.n_in + p.fi];
    if (!(fflags[s * p.RING + p.gslot] & MOVFE_FRAME_P)) return;
    for (int i = blockIdx.x * CAND_WARPS + warp; i < n; i += gridDim.x * CAND_WARPS) {
    bool pass = false;
    const movfe_rect mb = kps[((size_t)s * p.n_out + p.fi) * p.max_kps + i];
    const bool claimed = i < p.maxM && claim[(size_t)s * p.maxM + i] != 0x7fffffff;  // lbFound[i] (:381)
    uint32_t d[8];
    if (!claimed && rect_in_bounds(mb.x, mb.y, mb.w, mb.h, p.W, p.H)) {  // :388
        const uint8_t *roi = grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.W * p.H) + (size_t)mb.y * p.W + mb.x;
        if (express_test(roi, p.W, mb.h, mb.w, p.thr, scratch[warp], lane)) {  // :391
            const Band bd = express_band(roi, p.W, mb.h, mb.w, p.thr);
            express_mask(roi, p.W, mb.h, mb.w, bd, 1, false, d, lane);
            pass = true;
        }
    }
    if (lane == 0) birth_flag[(size_t)s * p.max_kps + i] = pass ? 1 : 0;
    if (pass) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (lane == k) birth_desc[((size_t)s * p.max_kps + i) * 8 + k] = d[k];
    }
    }
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
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        wsum[lane] = w;  // inclusive
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

// Stable (age desc, popcount desc) order of a table: bitonic sort of unique 64-bit keys in shared memory.
__device__ void sort_table(const movfe_track *__restrict__ tab, int n, unsigned long long *keys, uint16_t *__restrict__ order_out) {
    int N = 1;
    while (N < n) N <<= 1;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const movfe_track &t = tab[i];
            const uint32_t a = 0x7fffffffu - (uint32_t)max(t.age, 0);
            k = ((unsigned long long)a << 32) | ((unsigned long long)(256 - popc256(t.desc)) << 16) | (unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) order_out[i] = (uint16_t)(keys[i] & 0xffffu);
    __syncthreads();
}

__global__ void __launch_bounds__(FIN_THREADS)
finalize_kernel(ExtParams p, movfe_track *__restrict__ tracks, int32_t *__restrict__ ntracks,
                int32_t *__restrict__ cur_id, uint16_t *__restrict__ order, const Cand *__restrict__ cand,
                int32_t *__restrict__ claim, const movfe_rect *__restrict__ kps, const int32_t *__restrict__ nkps,
                const double *__restrict__ cov, const uint8_t *__restrict__ birth_flag,
                const uint32_t *__restrict__ birth_desc, const int4 *__restrict__ grid,
                const uint8_t *__restrict__ grey, const uint8_t *__restrict__ fflags) {
    extern __shared__ unsigned long long keys[];
    __shared__ int wsum[FIN_WARPS];
    __shared__ uint32_t scratch[FIN_WARPS][8];
    __shared__ int lat_flag[FIN_WARPS];
    const int s = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const movfe_track *prev = tracks + ((size_t)s * p.TSLOTS + p.tslot_prev) * p.maxT;
    movfe_track *cur = tracks + ((size_t)s * p.TSLOTS + p.tslot_cur) * p.maxT;
    const uint16_t *ord = order + (size_t)s * p.maxT;
    const Cand *cd = cand + (size_t)s * p.maxT;
    int32_t *cl = claim + (size_t)s * p.maxM;
    const int n_prev = ntracks[s * p.TSLOTS + p.tslot_prev];
    const uint8_t ff = fflags[s * p.RING + p.gslot];
    const bool is_p = ff & MOVFE_FRAME_P;
    const int n_kps = nkps[s * p.n_in + p.fi];
    int id = cur_id[s * p.TSLOTS + p.tslot_prev];
    int n_out = 0;  // logical size of the new table (entries beyond maxT are dropped)
    const uint8_t *img = p.has_grey ? grey + ((size_t)s * p.RING + p.gslot) * ((size_t)p.W * p.H) : nullptr;
    const int4 *g = grid + ((size_t)s * p.n_out + p.fi) * ((size_t)p.W * p.H);

    if (is_p) {
        // survivors in sorted order (:254-334)
        for (int base = 0; base < n_prev; base += FIN_THREADS) {
            const int i = base + threadIdx.x;
            bool acc = false;
            Cand c;
            if (i < n_prev) {
                c = cd[i];
                const bool mine = c.d_indx < 0 || c.d_indx >= p.maxM || cl[c.d_indx] == i;  // !lbFound at my turn
                acc = (c.flags & 1u) && mine && (c.flags & 2u);
            }
            int tot;
            const int pos = n_out + block_excl_scan(acc ? 1 : 0, wsum, tot);
            if (acc && pos < p.maxT) {
                const movfe_track &pv = prev[ord[i]];
                movfe_track t;
                t.pt_x = c.pt_x;
                t.pt_y = c.pt_y;
                t.mb = c.mb;
                t.track_id = pv.track_id;
                t.age = pv.age + 1;
                t.q_indx = i;
                t.flags = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) t.desc[k] = c.desc[k];
                cur[pos] = t;
            }
            n_out += tot;
        }
        // births in kps order (:379-416)
        int mov_cnt = 0;
        if (img) {
            for (int base = 0; base < n_kps; base += FIN_THREADS) {
                const int i = base + threadIdx.x;
                const bool b = i < n_kps && birth_flag[(size_t)s * p.max_kps + i];
                int tot;
                const int r = block_excl_scan(b ? 1 : 0, wsum, tot);
                if (b && n_out + r < p.maxT) {
                    const movfe_rect mb = kps[((size_t)s * p.n_out + p.fi) * p.max_kps + i];
                    movfe_track t;
                    // (mb.br() + mb.tl()) * 0.5 on Point_<int>: saturate_cast<int>(double) rounds half to even (:385)
                    t.pt_x = (float)__double2int_rn((mb.x + mb.w + mb.x) * 0.5);
                    t.pt_y = (float)__double2int_rn((mb.y + mb.h + mb.y) * 0.5);
                    t.mb = mb;
                    t.track_id = id + r + 1;  // ++mCurrentId
                    t.age = 0;
                    t.q_indx = -1;
                    t.flags = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) t.desc[k] = birth_desc[((size_t)s * p.max_kps + i) * 8 + k];
                    cur[n_out + r] = t;
                }
                n_out += tot;
                id += tot;
                mov_cnt += tot;
            }
        }
        // reset the claims this frame used (own kps only) for the next frame
        for (int i = threadIdx.x; i < p.maxM; i += FIN_THREADS) cl[i] = 0x7fffffff;
        // lattice pass: coverage back-fill (:418-451); I-frame seeding below shares the walker
        const bool backfill = img && (cov[s * p.n_in + p.fi] < p.cov_thr || mov_cnt < 60);
        if (backfill) {
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
                        const uint8_t *roi = img + (size_t)(y - 8) * p.W + (x - 8);
                        if (express_test(roi, p.W, 16, 16, p.thr, scratch[warp], lane) && !(__ldg(&g[(size_t)y * p.W + x]).x >= 0)) {
                            const Band bd = express_band(roi, p.W, 16, 16, p.thr);
                            express_mask(roi, p.W, 16, 16, bd, 1, false, d, lane);
                            pass = true;
                        }
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
                    t.flags = MOVFE_TRACK_COVERAGE;
#pragma unroll
                    for (int k = 0; k < 8; k++) t.desc[k] = d[k];
                    cur[n_out + before] = t;
                }
                n_out += tot;
                id += tot;
                __syncthreads();
            }
        }
    } else if (n_prev == 0 && img) {
        // I frame without previous features: seeding on the 16-px lattice (:123-157). With previous features the
        // reference carries them by LK (:81-120) — host work, dropped here.
        const int gw = (p.W - 16 + 15) / 16, gh = (p.H - 16 + 15) / 16;
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
                    const uint8_t *roi = img + (size_t)(y - 8) * p.W + (x - 8);
                    if (express_test(roi, p.W, 16, 16, p.thr, scratch[warp], lane)) {
                        const Band bd = express_band(roi, p.W, 16, 16, p.thr);
                        express_mask(roi, p.W, 16, 16, bd, 1, false, d, lane);
                        pass = true;
                    }
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
                t.flags = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) t.desc[k] = d[k];
                cur[n_out + before] = t;
            }
            n_out += tot;
            id += tot;
            __syncthreads();
        }
    }
    const int n_new = min(n_out, p.maxT);
    if (threadIdx.x == 0) {
        ntracks[s * p.TSLOTS + p.tslot_cur] = n_new;
        cur_id[s * p.TSLOTS + p.tslot_cur] = id;
    }
    __syncthreads();  // cur[] writes visible to the whole CTA before the sort reads them
    sort_table(cur, n_new, keys, order + (size_t)s * p.maxT);
}

// Sorts a table that was installed from the host (movfe_set_tracks).
__global__ void __launch_bounds__(FIN_THREADS)
sort_only_kernel(int maxT, int TSLOTS, int tslot, int stream, const movfe_track *__restrict__ tracks,
                 const int32_t *__restrict__ ntracks, uint16_t *__restrict__ order) {
    extern __shared__ unsigned long long keys[];
    const int s = stream;
    sort_table(tracks + ((size_t)s * TSLOTS + tslot) * maxT, ntracks[s * TSLOTS + tslot], keys, order + (size_t)s * maxT);
}

__global__ void fill_i32(int32_t *p, size_t n, int32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct ExtScratch {
    Cand *cand;
    int32_t *claim;
    uint8_t *birth_flag;
    uint32_t *birth_desc;
    uint16_t *order;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

ExtScratch carve(const movfe_ctx *ctx, size_t *total) {
    const movfe_config &c = ctx->cfg;
    const size_t S = c.n_streams;
    uint8_t *base = (uint8_t *)ctx->d_ext_scratch;
    size_t off = 0;
    ExtScratch e;
    e.cand = (Cand *)(base + off);
    off += align256(S * c.max_tracks * sizeof(Cand));
    e.claim = (int32_t *)(base + off);
    off += align256(S * c.max_records_per_frame * sizeof(int32_t));
    e.birth_flag = (uint8_t *)(base + off);
    off += align256(S * (size_t)ctx->max_kps);
    e.birth_desc = (uint32_t *)(base + off);
    off += align256(S * (size_t)ctx->max_kps * 32);
    e.order = (uint16_t *)(base + off);
    off += align256(S * c.max_tracks * sizeof(uint16_t));
    if (total) *total = off;
    return e;
}

size_t sort_smem(int maxT) {
    int N = 1;
    while (N < maxT) N <<= 1;
    return (size_t)N * sizeof(unsigned long long);
}

}  // namespace

size_t movfe_extract_scratch_bytes(const movfe_ctx *ctx) {
    size_t total = 0;
    carve(ctx, &total);
    return total;
}

static int tslot_of(const movfe_ctx *ctx, int64_t frame) {
    const int T = ctx->cfg.window_frames + 1;
    return (int)(((frame % T) + T) % T);
}

int movfe_extract_init(movfe_ctx *ctx) {
    const movfe_config &c = ctx->cfg;
    if (c.max_tracks > MAX_TRACKS_CAP) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "max_tracks must be <= %d", MAX_TRACKS_CAP);
    ExtScratch e = carve(ctx, nullptr);
    const size_t n = (size_t)c.n_streams * c.max_records_per_frame;
    fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(e.claim, n, 0x7fffffff);
    MOVFE_CUDA(ctx, cudaMemsetAsync(e.order, 0, (size_t)c.n_streams * c.max_tracks * sizeof(uint16_t), ctx->stream));
    MOVFE_CUDA(ctx, cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem(c.max_tracks)));
    MOVFE_CUDA(ctx, cudaFuncSetAttribute(sort_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem(c.max_tracks)));
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}

int movfe_extract_launch(movfe_ctx *ctx, int64_t first_frame, int n_frames) {
    const movfe_config &c = ctx->cfg;
    ExtScratch e = carve(ctx, nullptr);
    ProfScope prof(ctx, MOVFE_STAGE_EXTRACT);
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
        p.n_out = ctx->win_nout;
        p.n_in = ctx->win_nin;
        p.RING = ctx->RING;
        p.TSLOTS = c.window_frames + 1;
        p.fi = (int)(a - ctx->win_first);
        p.gslot = (int)(a % ctx->RING);
        p.tslot_prev = tslot_of(ctx, a - 1);
        p.tslot_cur = tslot_of(ctx, a);
        p.thr = c.express_threshold;
        p.has_grey = c.has_grey;
        p.cov_thr = c.coverage_threshold;
        // grid-stride over tracks / kps: enough CTAs to fill the chip, never one CTA per (mostly empty) capacity slot
        const int bps = std::max(4, (8 * ctx->sm_count + c.n_streams - 1) / c.n_streams);
        dim3 gc(std::min((c.max_tracks + CAND_WARPS - 1) / CAND_WARPS, bps), c.n_streams);
        cand_kernel<<<gc, CAND_WARPS * 32, 0, ctx->stream>>>(p, ctx->d_tracks, ctx->d_ntracks, e.order, ctx->d_grid,
                                                            ctx->d_hops, ctx->d_grey, ctx->d_fflags, e.cand, e.claim);
        int nl = 2;
        if (c.has_grey) {
            dim3 gb(std::min((ctx->max_kps + CAND_WARPS - 1) / CAND_WARPS, bps), c.n_streams);
            birth_kernel<<<gb, CAND_WARPS * 32, 0, ctx->stream>>>(p, ctx->d_kps, ctx->d_nkps, ctx->d_grey, ctx->d_fflags,
                                                                 e.claim, e.birth_flag, e.birth_desc);
            nl = 3;
        }
        finalize_kernel<<<c.n_streams, FIN_THREADS, sort_smem(c.max_tracks), ctx->stream>>>(
            p, ctx->d_tracks, ctx->d_ntracks, ctx->d_cur_id, e.order, e.cand, e.claim, ctx->d_kps, ctx->d_nkps, ctx->d_cov,
            e.birth_flag, e.birth_desc, ctx->d_grid, ctx->d_grey, ctx->d_fflags);
        prof.launches(nl);
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
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
    const int T = c.window_frames + 1;
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
    if (n_frames < 1 || ctx->win_first < 0 || first_frame < ctx->win_first || first_frame + n_frames > ctx->win_first + ctx->win_nout)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "extract: frames [%lld,%lld) are not inside the last raster window", (long long)first_frame,
                   (long long)(first_frame + n_frames));
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
    if (frame >= next || frame < next - 1 - ctx->cfg.window_frames)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "track table of frame %lld is not resident", (long long)frame);
    *ts = tslot_of(ctx, frame);
    return MOVFE_OK;
}

extern "C" int movfe_track_count(movfe_ctx *ctx, int stream, int64_t frame, int32_t *n_tracks, int32_t *current_id) {
    int ts;
    int rc = track_slot(ctx, stream, frame, &ts);
    if (rc) return rc;
    const int T = ctx->cfg.window_frames + 1;
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
    const int T = ctx->cfg.window_frames + 1;
    if (n > 0) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_tracks + ((size_t)stream * T + ts) * ctx->cfg.max_tracks, (size_t)n * sizeof(movfe_track),
                                        cudaMemcpyDeviceToHost, ctx->stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return n;
}
