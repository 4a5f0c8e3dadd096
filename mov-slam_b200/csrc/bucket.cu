// bucket.cu — the frame's keypoint bucket grid: Frame::AssignFeaturesToGrid (src/Frame.cc:356-388), Frame::PosInGrid
// (:670-680) and the radius query Frame::GetFeaturesInArea (:602-668), batched over independent keypoint sets, and the
// descriptor-verified search by projection over that grid (movfe_search_by_projection; ORB-SLAM3 lineage, include/movfe.h).
// The reference builds this 64x48 grid for every frame and never queries it on the MOV path; it is
// provided as a single-shot operator so that a caller that does query it finds the same lists in the same order.
//
// Layout: CSR per problem, cell = ix*48 + iy (mGrid[ix][iy]); a cell's items are keypoint indices in insertion order,
// which is what `mGrid[x][y].push_back(i)` for i = 0..N-1 produces. Built as one stable sort per keypoint set: the key is
// (cell << 14 | index), so equal cells keep ascending index and no atomics are involved.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int BG_COLS = 64, BG_ROWS = 48;  // FRAME_GRID_COLS / FRAME_GRID_ROWS (include/Frame.h:40-41)
constexpr int BG_CELLS = BG_COLS * BG_ROWS;
constexpr int BG_THREADS = 1024;
constexpr int BG_MAX_N = 16384;            // keypoints per set (index field of the sort key: 14 bits)
constexpr uint32_t BG_INVALID = 0xffffffffu;

// Frame.cc:670-680 with mnMinX = mnMinY = 0 (undistorted mono frame, Frame.cc:739-745): round(), not floor()
__device__ __forceinline__ bool pos_in_grid(float x, float y, float w_inv, float h_inv, int &px, int &py) {
    px = (int)roundf(__fmul_rn(__fsub_rn(x, 0.0f), w_inv));
    py = (int)roundf(__fmul_rn(__fsub_rn(y, 0.0f), h_inv));
    return !(px < 0 || px >= BG_COLS || py < 0 || py >= BG_ROWS);
}

__global__ void __launch_bounds__(BG_THREADS)
assign_kernel(const float *__restrict__ pts, int stride, const int32_t *__restrict__ off, float w_inv, float h_inv, int N2,
              int32_t *__restrict__ cell_start, int32_t *__restrict__ cell_items) {  // point i = pts[i*stride], pts[i*stride + 1]
    extern __shared__ uint32_t keys[];  // [N2], N2 = power of two >= the largest set
    const int pidx = blockIdx.x;
    const int b = off[pidx], n = off[pidx + 1] - b;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) {
        uint32_t k = BG_INVALID;
        if (i < n) {
            const float2 q = *reinterpret_cast<const float2 *>(pts + (size_t)(b + i) * stride);
            int px, py;
            if (pos_in_grid(q.x, q.y, w_inv, h_inv, px, py)) k = ((uint32_t)(px * BG_ROWS + py) << 14) | (uint32_t)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    // bitonic sort, ascending; keys are distinct (the index is part of the key), so the result is the stable order
    for (int k = 2; k <= N2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N2 >> 1); t += blockDim.x) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
                const int hi = lo | j;
                const uint32_t a = keys[lo], c = keys[hi];
                const bool up = (lo & k) == 0;
                if ((a > c) == up) {
                    keys[lo] = c;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    int32_t *cs = cell_start + (size_t)pidx * (BG_CELLS + 1);
    for (int c = threadIdx.x; c <= BG_CELLS; c += blockDim.x) {  // lower bound of (c << 14): first item of cell c
        const uint32_t want = (uint32_t)c << 14;
        int lo = 0, hi = n;  // invalid keys sort last and are >= any (c << 14), c <= 3072
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid] < want) lo = mid + 1;
            else hi = mid;
        }
        cs[c] = lo;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (keys[i] != BG_INVALID) cell_items[b + i] = (int32_t)(keys[i] & 0x3fffu);
}

// Frame.cc:602-668 with minLevel = 0, maxLevel = -1 (every MOV keypoint is octave 0, Frame.cc:105-118): one thread per query
__global__ void area_kernel(const float2 *__restrict__ pts, const int32_t *__restrict__ off, const int32_t *__restrict__ cell_start,
                            const int32_t *__restrict__ cell_items, const movfe_area_query *__restrict__ queries, int n_queries,
                            float w_inv, float h_inv, int capacity, int32_t *__restrict__ out, int32_t *__restrict__ counts) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    const movfe_area_query Q = queries[q];
    const float x = Q.x, y = Q.y, r = Q.r;
    const int b = off[Q.problem];
    const int32_t *cs = cell_start + (size_t)Q.problem * (BG_CELLS + 1);
    int cnt = 0;
    const int min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, 0.0f), r), w_inv)));
    const int max_cx = min(BG_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, 0.0f), r), w_inv)));
    const int min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, 0.0f), r), h_inv)));
    const int max_cy = min(BG_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, 0.0f), r), h_inv)));
    if (min_cx < BG_COLS && max_cx >= 0 && min_cy < BG_ROWS && max_cy >= 0) {  // the four early returns of :611-625
        for (int ix = min_cx; ix <= max_cx; ix++)
            for (int iy = min_cy; iy <= max_cy; iy++) {
                const int c = ix * BG_ROWS + iy;
                for (int j = cs[c]; j < cs[c + 1]; j++) {
                    const int item = cell_items[b + j];
                    const float2 kp = pts[b + item];
                    if (fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r) {
                        if (cnt < capacity) out[(size_t)q * capacity + cnt] = item;
                        cnt++;
                    }
                }
            }
    }
    counts[q] = cnt;
}

// ------------------------------------------------------------------------------------------ search by projection -----
// A warp per map point: the cells of GetFeaturesInArea(u, v, r) are walked in the reference's order, a cell's keypoints 32 at a
// time; a candidate's position in that order rides in the low bits of its key, so the warp-wide minimum is the FIRST candidate
// with the smallest distance, as the sequential `if (dist < bestDist)` leaves it. The second smallest distance of the multiset
// (what `else if (dist < bestDist2)` leaves) is merged beside it. Keypoints chosen by several map points are settled by a 64-bit
// minimum per keypoint over (distance, point index): the keys are distinct, the result does not depend on the order of arrival.
constexpr int SP_WARPS = 4;
constexpr uint32_t SP_NONE = (256u << 20) | 0xfffffu;  // bestDist = 256 (never reached by `dist < bestDist`), no candidate

__device__ __forceinline__ int problem_of(const int32_t *__restrict__ off, int n_problems, int g) {
    int lo = 0, hi = n_problems;  // invariant off[lo] <= g < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&off[mid]) <= g) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(SP_WARPS * 32)
search_kernel(const movfe_track *__restrict__ feat, const uint8_t *__restrict__ taken, const int32_t *__restrict__ feat_off,
              const int32_t *__restrict__ cell_start, const int32_t *__restrict__ cell_items, const movfe_map_point *__restrict__ pts,
              const movfe_projection *__restrict__ proj, const uint4 *__restrict__ pt_desc, const int32_t *__restrict__ pt_off,
              int n_problems, int n_pts, movfe_projection_search_params prm, float w_inv, float h_inv,
              unsigned long long *__restrict__ winner, int32_t *__restrict__ prop, int32_t *__restrict__ pt_dist) {
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (g >= n_pts) return;  // warp-uniform
    const int p = problem_of(pt_off, n_problems, g);
    const int fb = __ldg(&feat_off[p]);
    const int32_t *cs = cell_start + (size_t)p * (BG_CELLS + 1);
    const movfe_projection pr = proj[g];
    const uint32_t flags = pts[g].flags;
    uint32_t key = SP_NONE;
    int idx = -1, second = 256;
    const bool live = pr.in_view && !(prm.far_points && pr.depth > prm.th_far) && !(flags & (MOVFE_MP_BAD | MOVFE_MP_SKIP | MOVFE_MP_NULL));
    if (live) {
        const float x = pr.u, y = pr.v;
        const float r = __fmul_rn(pr.view_cos > 0.998f ? 2.5f : 4.0f, prm.th);
        const int min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, 0.0f), r), w_inv)));
        const int max_cx = min(BG_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, 0.0f), r), w_inv)));
        const int min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, 0.0f), r), h_inv)));
        const int max_cy = min(BG_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, 0.0f), r), h_inv)));
        if (min_cx < BG_COLS && max_cx >= 0 && min_cy < BG_ROWS && max_cy >= 0) {
            const uint4 q0 = __ldg(&pt_desc[2 * (size_t)g]), q1 = __ldg(&pt_desc[2 * (size_t)g + 1]);
            int base = 0;  // keypoints of the cells walked so far: a candidate's place in the reference's order
            for (int ix = min_cx; ix <= max_cx; ix++) {
                const int c0 = ix * BG_ROWS + min_cy, c1 = ix * BG_ROWS + max_cy;
                // the cells (ix, min_cy..max_cy) are adjacent in the CSR: one run of items
                const int s0 = __ldg(&cs[c0]), s1 = __ldg(&cs[c1 + 1]);
                for (int j = s0 + lane; j < s1; j += 32) {
                    const int item = __ldg(&cell_items[fb + j]);
                    const uint4 *rec = reinterpret_cast<const uint4 *>(feat + fb + item);
                    const uint4 r0 = __ldg(rec);
                    const float kx = __uint_as_float(r0.x), ky = __uint_as_float(r0.y);
                    if (fabsf(__fsub_rn(kx, x)) < r && fabsf(__fsub_rn(ky, y)) < r && !(taken && taken[fb + item])) {
                        const uint4 d0 = __ldg(rec + 2), d1 = __ldg(rec + 3);
                        const int d = __popc(d0.x ^ q0.x) + __popc(d0.y ^ q0.y) + __popc(d0.z ^ q0.z) + __popc(d0.w ^ q0.w) +
                                      __popc(d1.x ^ q1.x) + __popc(d1.y ^ q1.y) + __popc(d1.z ^ q1.z) + __popc(d1.w ^ q1.w);
                        const int bd = (int)(key >> 20);
                        if (d < bd) {  // this lane's candidates arrive in order: strict '<' keeps the earlier of equals
                            second = bd;
                            key = ((uint32_t)d << 20) | (uint32_t)(base + j - s0);
                            idx = item;
                        } else if (d < second) {
                            second = d;
                        }
                    }
                }
                base += s1 - s0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, key, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        second = min(min(second, os), (int)max(key >> 20, ok >> 20));  // the larger of the two bests is a second best
        if (ok < key) {
            key = ok;
            idx = oi;
        }
    }
    const int best = (int)(key >> 20);
    bool accept = idx >= 0 && best <= prm.th_high;
    // bestLevel == bestLevel2 holds exactly when a second candidate below 256 was seen (all keypoints are octave 0)
    if (accept && second < 256 && (float)best > __fmul_rn(prm.nn_ratio, (float)second)) accept = false;
    if (lane == 0) {
        prop[g] = accept ? idx : -1;
        pt_dist[g] = accept ? best : -1;
        if (accept) atomicMin(&winner[fb + idx], ((unsigned long long)(unsigned)best << 32) | (unsigned)(g - __ldg(&pt_off[p])));
    }
}

__global__ void resolve_kernel(const int32_t *__restrict__ feat_off, const int32_t *__restrict__ pt_off, int n_problems, int n_pts,
                               const unsigned long long *__restrict__ winner, const int32_t *__restrict__ prop,
                               const int32_t *__restrict__ pt_dist, int32_t *__restrict__ feat_match, int32_t *__restrict__ pt_match,
                               int32_t *__restrict__ n_matches) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pts) return;
    const int f = prop[g];
    int m = -1;
    if (f >= 0) {
        const int p = problem_of(pt_off, n_problems, g);
        const int fb = feat_off[p], k = g - pt_off[p];
        if (winner[fb + f] == (((unsigned long long)(unsigned)pt_dist[g] << 32) | (unsigned)k)) {
            m = f;
            feat_match[fb + f] = k;
            atomicAdd(&n_matches[p], 1);  // a count: order-independent
        }
    }
    pt_match[g] = m;
}

}  // namespace

int movfe_ensure_op_scratch(movfe_ctx *ctx, size_t bytes);  // api.cu

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int movfe_assign_features_to_grid(movfe_ctx *ctx, int n_problems, const float *pts_xy, const int32_t *off,
                                             int32_t *cell_start, int32_t *cell_items) {
    if (!ctx) return MOVFE_E_INVALID;
    if (n_problems < 1 || !off || !cell_start) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "assign_features_to_grid: bad argument");
    const int n = off[n_problems];
    int largest = 0;
    for (int p = 0; p < n_problems; p++) {
        if (off[p + 1] < off[p]) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "assign_features_to_grid: offsets must not decrease");
        largest = std::max(largest, off[p + 1] - off[p]);
    }
    if (n < 0 || (n > 0 && (!pts_xy || !cell_items))) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "assign_features_to_grid: null array");
    if (largest > BG_MAX_N) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "assign_features_to_grid: %d keypoints in one set, limit %d", largest, BG_MAX_N);
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int N2 = 64;
    while (N2 < largest) N2 <<= 1;
    const size_t b_pts = a256((size_t)std::max(n, 1) * sizeof(float2)), b_off = a256((size_t)(n_problems + 1) * 4);
    const size_t b_cs = a256((size_t)n_problems * (BG_CELLS + 1) * 4), b_it = a256((size_t)std::max(n, 1) * 4);
    int rc = movfe_ensure_op_scratch(ctx, b_pts + b_off + b_cs + b_it);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_op;
    float2 *d_pts = (float2 *)base;
    int32_t *d_off = (int32_t *)(base + b_pts), *d_cs = (int32_t *)(base + b_pts + b_off), *d_it = (int32_t *)(base + b_pts + b_off + b_cs);
    cudaStream_t st = ctx->stream;
    if (n) MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pts, pts_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n) MOVFE_CUDA(ctx, cudaMemsetAsync(d_it, 0xff, (size_t)n * 4, st));  // entries past a set's valid count stay -1
    const float w_inv = (float)BG_COLS / (float)ctx->cfg.width, h_inv = (float)BG_ROWS / (float)ctx->cfg.height;  // Frame.cc:147-148
    const size_t smem = (size_t)N2 * sizeof(uint32_t);
    assign_kernel<<<n_problems, BG_THREADS, smem, st>>>(reinterpret_cast<const float *>(d_pts), 2, d_off, w_inv, h_inv, N2, d_cs, d_it);
    MOVFE_CUDA(ctx, cudaGetLastError());
    MOVFE_CUDA(ctx, cudaMemcpyAsync(cell_start, d_cs, (size_t)n_problems * (BG_CELLS + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (n) MOVFE_CUDA(ctx, cudaMemcpyAsync(cell_items, d_it, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
    return MOVFE_OK;
}

extern "C" int movfe_features_in_area(movfe_ctx *ctx, int n_problems, const float *pts_xy, const int32_t *off,
                                      const int32_t *cell_start, const int32_t *cell_items, int n_queries,
                                      const movfe_area_query *queries, int capacity, int32_t *out, int32_t *counts) {
    if (!ctx) return MOVFE_E_INVALID;
    if (n_problems < 1 || !off || !cell_start || n_queries < 0 || capacity < 0 || (n_queries && (!queries || !counts)) ||
        (n_queries && capacity && !out))
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "features_in_area: bad argument");
    const int n = off[n_problems];
    if (n < 0 || (n > 0 && (!pts_xy || !cell_items))) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "features_in_area: null array");
    for (int q = 0; q < n_queries; q++)
        if (queries[q].problem < 0 || queries[q].problem >= n_problems)
            MOVFE_FAIL(ctx, MOVFE_E_INVALID, "features_in_area: query %d names set %d of %d", q, queries[q].problem, n_problems);
    if (n_queries == 0) return MOVFE_OK;
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t b_pts = a256((size_t)std::max(n, 1) * sizeof(float2)), b_off = a256((size_t)(n_problems + 1) * 4);
    const size_t b_cs = a256((size_t)n_problems * (BG_CELLS + 1) * 4), b_it = a256((size_t)std::max(n, 1) * 4);
    const size_t b_q = a256((size_t)n_queries * sizeof(movfe_area_query)), b_out = a256((size_t)n_queries * std::max(capacity, 1) * 4);
    const size_t b_cnt = a256((size_t)n_queries * 4);
    int rc = movfe_ensure_op_scratch(ctx, b_pts + b_off + b_cs + b_it + b_q + b_out + b_cnt);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_op;
    float2 *d_pts = (float2 *)base;
    base += b_pts;
    int32_t *d_off = (int32_t *)base;
    base += b_off;
    int32_t *d_cs = (int32_t *)base;
    base += b_cs;
    int32_t *d_it = (int32_t *)base;
    base += b_it;
    movfe_area_query *d_q = (movfe_area_query *)base;
    base += b_q;
    int32_t *d_out = (int32_t *)base;
    base += b_out;
    int32_t *d_cnt = (int32_t *)base;
    cudaStream_t st = ctx->stream;
    if (n) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pts, pts_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_it, cell_items, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_cs, cell_start, (size_t)n_problems * (BG_CELLS + 1) * 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_q, queries, (size_t)n_queries * sizeof(movfe_area_query), cudaMemcpyHostToDevice, st));
    const float w_inv = (float)BG_COLS / (float)ctx->cfg.width, h_inv = (float)BG_ROWS / (float)ctx->cfg.height;
    area_kernel<<<(n_queries + 127) / 128, 128, 0, st>>>(d_pts, d_off, d_cs, d_it, d_q, n_queries, w_inv, h_inv, capacity, d_out, d_cnt);
    MOVFE_CUDA(ctx, cudaGetLastError());
    if (capacity) MOVFE_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n_queries * capacity * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(counts, d_cnt, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
    return MOVFE_OK;
}

// The bucket grid of a RESIDENT track table (the frame's keypoints never leave the device): Frame::AssignFeaturesToGrid as
// the Frame constructors call it (src/Frame.cc:118,216), for frame `frame` of stream `stream` after movfe_extract.
extern "C" int movfe_track_feature_grid(movfe_ctx *ctx, int stream, int64_t frame, int32_t *cell_start, int32_t *cell_items,
                                        int capacity) {
    if (!ctx) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    if (stream < 0 || stream >= c.n_streams || !cell_start || capacity < 0 || (capacity && !cell_items))
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "track_feature_grid: bad argument");
    const int64_t next = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    if (frame >= next || frame < next - 1 - c.window_frames)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "track table of frame %lld is not resident", (long long)frame);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    const int T = ctx->TSLOTS;
    const int ts = (int)(((frame % T) + T) % T);
    int32_t n = 0;
    cudaStream_t st = ctx->stream;
    MOVFE_CUDA(ctx, cudaMemcpyAsync(&n, ctx->d_ntracks + stream * T + ts, 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
    if (n > capacity) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "track_feature_grid: %d keypoints, capacity %d", n, capacity);
    if (n > BG_MAX_N) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "track_feature_grid: %d keypoints, limit %d", n, BG_MAX_N);
    const size_t b_off = a256(8), b_cs = a256((size_t)(BG_CELLS + 1) * 4), b_it = a256((size_t)std::max(n, 1) * 4);
    int rc = movfe_ensure_op_scratch(ctx, b_off + b_cs + b_it);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_op;
    int32_t *d_off = (int32_t *)base, *d_cs = (int32_t *)(base + b_off), *d_it = (int32_t *)(base + b_off + b_cs);
    const int32_t off[2] = {0, n};
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_off, off, 8, cudaMemcpyHostToDevice, st));
    if (n) MOVFE_CUDA(ctx, cudaMemsetAsync(d_it, 0xff, (size_t)n * 4, st));
    int N2 = 64;
    while (N2 < n) N2 <<= 1;
    const float w_inv = (float)BG_COLS / (float)c.width, h_inv = (float)BG_ROWS / (float)c.height;
    const size_t smem = (size_t)N2 * sizeof(uint32_t);
    const movfe_track *tab = ctx->d_tracks + ((size_t)stream * T + ts) * c.max_tracks;  // pt_x, pt_y lead the 64-byte record
    assign_kernel<<<1, BG_THREADS, smem, st>>>(reinterpret_cast<const float *>(tab), (int)(sizeof(movfe_track) / 4), d_off, w_inv, h_inv,
                                               N2, d_cs, d_it);
    MOVFE_CUDA(ctx, cudaGetLastError());
    MOVFE_CUDA(ctx, cudaMemcpyAsync(cell_start, d_cs, (size_t)(BG_CELLS + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (n) MOVFE_CUDA(ctx, cudaMemcpyAsync(cell_items, d_it, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));  // `off` and `n` live on this stack frame
    return n;
}

extern "C" int movfe_search_by_projection(movfe_ctx *ctx, int n_problems, const movfe_track *feat, const uint8_t *feat_taken,
                                          const int32_t *feat_off, const movfe_map_point *pts, const movfe_projection *proj,
                                          const uint32_t *pt_desc, const int32_t *pt_off, const movfe_projection_search_params *prm,
                                          int32_t *feat_match, int32_t *pt_match, int32_t *pt_dist, int32_t *n_matches) {
    if (!ctx) return MOVFE_E_INVALID;
    if (n_problems < 1 || !feat_off || !pt_off || !prm || !n_matches) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "search_by_projection: bad argument");
    const int n = feat_off[n_problems], m = pt_off[n_problems];
    int largest = 0;
    for (int p = 0; p < n_problems; p++) {
        if (feat_off[p + 1] < feat_off[p] || pt_off[p + 1] < pt_off[p])
            MOVFE_FAIL(ctx, MOVFE_E_INVALID, "search_by_projection: offsets must not decrease");
        largest = std::max(largest, feat_off[p + 1] - feat_off[p]);
    }
    if (feat_off[0] != 0 || pt_off[0] != 0) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "search_by_projection: offsets must start at 0");
    if ((n > 0 && (!feat || !feat_match)) || (m > 0 && (!pts || !proj || !pt_desc || !pt_match || !pt_dist)))
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "search_by_projection: null array");
    if (largest > BG_MAX_N) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "search_by_projection: %d keypoints in one frame, limit %d", largest, BG_MAX_N);
    for (int p = 0; p < n_problems; p++) n_matches[p] = 0;
    for (int i = 0; i < n; i++) feat_match[i] = -1;
    for (int k = 0; k < m; k++) pt_match[k] = pt_dist[k] = -1;
    if (n == 0 || m == 0) return MOVFE_OK;
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int N2 = 64;
    while (N2 < largest) N2 <<= 1;
    const size_t b_feat = a256((size_t)n * sizeof(movfe_track)), b_tk = a256((size_t)n), b_fo = a256((size_t)(n_problems + 1) * 4);
    const size_t b_cs = a256((size_t)n_problems * (BG_CELLS + 1) * 4), b_it = a256((size_t)n * 4), b_win = a256((size_t)n * 8);
    const size_t b_fm = a256((size_t)n * 4), b_pts = a256((size_t)m * sizeof(movfe_map_point)), b_pr = a256((size_t)m * sizeof(movfe_projection));
    const size_t b_pd = a256((size_t)m * 32), b_po = b_fo, b_m4 = a256((size_t)m * 4), b_nm = a256((size_t)n_problems * 4);
    int rc = movfe_ensure_op_scratch(ctx, b_feat + b_tk + b_fo + b_cs + b_it + b_win + b_fm + b_pts + b_pr + b_pd + b_po + 3 * b_m4 + b_nm);
    if (rc) return rc;
    uint8_t *base = (uint8_t *)ctx->d_op;
    auto take = [&](size_t bytes) {
        uint8_t *q = base;
        base += bytes;
        return q;
    };
    movfe_track *d_feat = (movfe_track *)take(b_feat);
    uint8_t *d_tk = take(b_tk);
    int32_t *d_fo = (int32_t *)take(b_fo), *d_cs = (int32_t *)take(b_cs), *d_it = (int32_t *)take(b_it);
    unsigned long long *d_win = (unsigned long long *)take(b_win);
    int32_t *d_fm = (int32_t *)take(b_fm);
    movfe_map_point *d_pts = (movfe_map_point *)take(b_pts);
    movfe_projection *d_pr = (movfe_projection *)take(b_pr);
    uint4 *d_pd = (uint4 *)take(b_pd);
    int32_t *d_po = (int32_t *)take(b_po), *d_prop = (int32_t *)take(b_m4), *d_pm = (int32_t *)take(b_m4), *d_dist = (int32_t *)take(b_m4);
    int32_t *d_nm = (int32_t *)take(b_nm);
    cudaStream_t st = ctx->stream;
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_feat, feat, (size_t)n * sizeof(movfe_track), cudaMemcpyHostToDevice, st));
    if (feat_taken) MOVFE_CUDA(ctx, cudaMemcpyAsync(d_tk, feat_taken, (size_t)n, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_fo, feat_off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_po, pt_off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pts, pts, (size_t)m * sizeof(movfe_map_point), cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pr, proj, (size_t)m * sizeof(movfe_projection), cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pd, pt_desc, (size_t)m * 32, cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemsetAsync(d_it, 0xff, (size_t)n * 4, st));
    MOVFE_CUDA(ctx, cudaMemsetAsync(d_win, 0xff, (size_t)n * 8, st));
    MOVFE_CUDA(ctx, cudaMemsetAsync(d_fm, 0xff, (size_t)n * 4, st));
    MOVFE_CUDA(ctx, cudaMemsetAsync(d_nm, 0, (size_t)n_problems * 4, st));
    const float w_inv = (float)BG_COLS / (float)ctx->cfg.width, h_inv = (float)BG_ROWS / (float)ctx->cfg.height;  // Frame.cc:147-148
    assign_kernel<<<n_problems, BG_THREADS, (size_t)N2 * sizeof(uint32_t), st>>>(reinterpret_cast<const float *>(d_feat),
                                                                                  (int)(sizeof(movfe_track) / 4), d_fo, w_inv, h_inv, N2, d_cs, d_it);
    search_kernel<<<(m + SP_WARPS - 1) / SP_WARPS, SP_WARPS * 32, 0, st>>>(d_feat, feat_taken ? d_tk : nullptr, d_fo, d_cs, d_it, d_pts, d_pr, d_pd,
                                                                            d_po, n_problems, m, *prm, w_inv, h_inv, d_win, d_prop, d_dist);
    resolve_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_fo, d_po, n_problems, m, d_win, d_prop, d_dist, d_fm, d_pm, d_nm);
    MOVFE_CUDA(ctx, cudaGetLastError());
    MOVFE_CUDA(ctx, cudaMemcpyAsync(feat_match, d_fm, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(pt_match, d_pm, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(pt_dist, d_dist, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(n_matches, d_nm, (size_t)n_problems * 4, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
    return MOVFE_OK;
}

int movfe_bucket_init(movfe_ctx *ctx) {
    MOVFE_CUDA(ctx, optin_dynamic_smem(assign_kernel, ctx->smem_optin));
    return MOVFE_OK;
}
