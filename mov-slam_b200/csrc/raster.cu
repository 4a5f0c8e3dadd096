// raster.cu — motion-vector side data -> hop lists, candidate-keypoint lists, per-pixel slot grid.
// Replaces the MV loop of VideoDecoder::NextImage (src/VideoDecoder.cc:211-350), batched over
// (stream x frame-window). Compiled with -fmad=false: the float expressions below must round exactly like the
// reference's non-contracted binary32 arithmetic ("float bit-exactness").
//
// Kernels:
//   ingest_kernel   40-byte AVMotionVector records -> 16-byte Rec16 in the per-stream frame ring.
//                   Coalesced 128-bit loads of the packed records, staged through shared memory.
//   count_kernel    per frame: how many records fall in each order class (hop segment k, own kps, chained kps r).
//   bases_kernel    per target frame: start of every segment of its hop / kps list (exclusive sums over the
//                   look-ahead frames) -> every output index is known without atomics.
//   emit_kernel     per record: hop j = ref+1..1 written at its final index in frame f-(j-1)'s list, kps rectangle
//                   written at its final index; order = the reference's push_back order (ballot ranks).
//   bbox_kernel     y- and x-extent of every 32-hop chunk (lets the grid kernel skip chunks without reading them).
//   grid_kernel     (grid.cu) one CTA owns a 32-row band of one frame's grid: the hops touching the band go into
//                   per-tile queues in shared memory in list order, then one warp per 32x32 tile resolves the slots
//                   once per (column run, row run) cell. Every pixel is written exactly once with one 128-bit
//                   streaming store (no atomics, no read-modify-write, the -1 fill is implicit).
// All of these run on the context's raster stream, beside the propagation of the previous window.
#include <algorithm>
#include <cstdio>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ ingest -----
constexpr int INGEST_WARPS = 8;
constexpr int INGEST_REC_PER_WARP = 64;  // 64 records = 2560 B = 160 x 16 B: five coalesced 128-bit loads per lane

__global__ void __launch_bounds__(INGEST_WARPS * 32)
ingest_kernel(const uint4 *__restrict__ recs16, int64_t n_records, const int64_t *__restrict__ rec_off, int n_seg,
              int n_frames, int64_t first_abs, int RING, int maxM, Rec16 *__restrict__ d_rec,
              unsigned long long *__restrict__ rejected) {
    __shared__ uint4 stage[INGEST_WARPS][INGEST_REC_PER_WARP * 40 / 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * INGEST_WARPS + warp) * INGEST_REC_PER_WARP;  // first record of this warp
    if (w0 >= n_records) return;
    const int64_t full16 = n_records * 40 / 16;  // 128-bit words that lie wholly inside the record array
    const int64_t base16 = w0 * 40 / 16;         // 2560*k/16: exact
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const int q = i * 32 + lane;
        if (base16 + q < full16) {
            stage[warp][q] = __ldg(&recs16[base16 + q]);
        } else if (base16 + q == full16 && (n_records & 1)) {
            // an odd record count ends in the middle of a word: only its first 8 bytes belong to the caller's array
            const uint2 h = __ldg(reinterpret_cast<const uint2 *>(&recs16[base16 + q]));
            stage[warp][q] = make_uint4(h.x, h.y, 0u, 0u);
        }
    }
    __syncwarp();
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage[warp]);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int li = j * 32 + lane;
        const int64_t i = w0 + li;
        if (i >= n_records) continue;
        const uint32_t *r = sw + li * 10;  // 40-byte record = 10 words (layout: include/movfe_types.h)
        Rec16 o;
        const int32_t source = (int32_t)r[0];
        o.w = (uint8_t)(r[1] & 0xff);
        o.h = (uint8_t)((r[1] >> 8) & 0xff);
        o.sx = (int16_t)(r[1] >> 16);
        o.sy = (int16_t)(r[2] & 0xffff);
        o.dx = (int16_t)(r[2] >> 16);
        o.dy = (int16_t)(r[3] & 0xffff);
        o.src_sign = source < 0 ? -1 : (source > 0 ? 1 : 0);
        o.pad = 0;
        o.ref = (int32_t)r[9];
        // segment (stream, frame) of record i: last seg with rec_off[seg] <= i
        int lo = 0, hi = n_seg;  // invariant rec_off[lo] <= i < rec_off[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(&rec_off[mid]) <= i) lo = mid; else hi = mid;
        }
        const int s = lo / n_frames, f = lo - s * n_frames;
        const int64_t pos = i - __ldg(&rec_off[lo]);
        if (pos < maxM) {
            const int slot = (int)((first_abs + f) % RING);
            d_rec[((size_t)s * RING + slot) * maxM + pos] = o;
        } else {
            atomicAdd(rejected, 1ull);  // error counter only; never on the data path
        }
    }
}

// 16-byte records as the host packed them (movfe_push_frames_packed): one 128-bit load and store per record.
__global__ void __launch_bounds__(256)
ingest_packed_kernel(const uint4 *__restrict__ recs, int64_t n_records, const int64_t *__restrict__ rec_off, int n_seg, int n_frames,
                     int64_t first_abs, int RING, int maxM, Rec16 *__restrict__ d_rec, unsigned long long *__restrict__ rejected) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_records) return;
    uint4 r = __ldg(&recs[i]);
    r.z &= 0x00ffffffu;  // the reserved byte is not part of the record
    int lo = 0, hi = n_seg;  // segment (stream, frame) of record i: invariant rec_off[lo] <= i < rec_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&rec_off[mid]) <= i) lo = mid; else hi = mid;
    }
    const int s = lo / n_frames, f = lo - s * n_frames;
    const int64_t pos = i - __ldg(&rec_off[lo]);
    if (pos < maxM) {
        const int slot = (int)((first_abs + f) % RING);
        reinterpret_cast<uint4 *>(d_rec)[((size_t)s * RING + slot) * maxM + pos] = r;
    } else {
        atomicAdd(rejected, 1ull);  // error counter only; never on the data path
    }
}

__global__ void ingest_meta_kernel(const int64_t *__restrict__ rec_off, const uint8_t *__restrict__ flags, int n_seg,
                                   int n_frames, int64_t first_abs, int RING, int maxM, int32_t *__restrict__ rec_cnt,
                                   uint8_t *__restrict__ fflags) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    const int s = seg / n_frames, f = seg - s * n_frames;
    const int slot = (int)((first_abs + f) % RING);
    const int64_t n = rec_off[seg + 1] - rec_off[seg];
    rec_cnt[s * RING + slot] = (int32_t)(n < maxM ? n : maxM);
    fflags[s * RING + slot] = flags[seg];
}

// Grey planes [segment][H][W] (packed, as the decoder hands them over) -> the ring [S][slot][H][pitch]. The power-of-two
// pitch makes every row offset inside a patch a compile-time constant for the propagation kernels (extract.cu).
__global__ void grey_ingest_kernel(const uint8_t *__restrict__ src, int n_seg, int n_frames, int W, int H, int pitch, int vec,
                                   int64_t first_abs, int RING, uint8_t *__restrict__ ring) {
    const int wv = W / vec;
    const int64_t total = (int64_t)n_seg * H * wv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % wv);
        const int64_t row = i / wv;
        const int y = (int)(row % H);
        const int seg = (int)(row / H);
        const int s = seg / n_frames, f = seg - s * n_frames;
        const int slot = (int)((first_abs + f) % RING);
        const size_t so = ((size_t)seg * H + y) * W + (size_t)x * vec;
        const size_t dof = (((size_t)s * RING + slot) * H + y) * pitch + (size_t)x * vec;
        if (vec == 16) *reinterpret_cast<uint4 *>(ring + dof) = __ldg(reinterpret_cast<const uint4 *>(src + so));
        else ring[dof] = src[so];
    }
}

// ------------------------------------------------------------------------------------------ record classes ---
struct RecInfo {
    bool valid;    // passes the skip rule (VideoDecoder.cc:236-241) and ref <= K
    bool pbranch;  // source <= 0: hops + coverage (VideoDecoder.cc:287)
    bool chained;  // ref > 0 && source < 0: block goes to frame N-1-ref's kps (VideoDecoder.cc:245)
    bool bad;      // ref > K
    int  nh;       // number of hops = ref+1 for the P branch
    float mv_x, mv_y, half_w, half_h;
    int  kx, ky;   // top-left of dMB
};

__device__ __forceinline__ RecInfo classify(const Rec16 &r, int W, int H, int K) {
    RecInfo c;
    c.bad = r.ref > K;
    const float mb_w = (float)r.w, mb_h = (float)r.h;          // :215-216
    c.half_w = mb_w / 2;                                        // :217-218
    c.half_h = mb_h / 2;
    const float dxs = (float)((int)r.dx - (int)r.sx);           // :220-221
    const float dys = (float)((int)r.dy - (int)r.sy);
    const float den = (float)(r.ref + 1);
    c.mv_x = __fdiv_rn(dxs, den);                               // :223-224
    c.mv_y = __fdiv_rn(dys, den);
    c.chained = r.ref > 0 && r.src_sign < 0;
    const float dst_x = c.chained ? (float)r.sx : (float)r.dx;  // :227-228
    const float dst_y = c.chained ? (float)r.sy : (float)r.dy;
    float d_x_top = __fsub_rn(dst_x, c.half_w);                 // :230-235
    if (d_x_top < 0) d_x_top = 0;
    float d_y_top = __fsub_rn(dst_y, c.half_h);
    if (d_y_top < 0) d_y_top = 0;
    const float d_x_bottom = __fadd_rn(dst_x, c.half_w);        // :236-241
    const float d_y_bottom = __fadd_rn(dst_y, c.half_h);
    c.valid = !c.bad && !(d_x_bottom >= (float)W) && !(d_y_bottom >= (float)H);
    c.kx = (int)d_x_top;                                        // :244 cv::Rect(float...) truncates
    c.ky = (int)d_y_top;
    c.pbranch = r.src_sign <= 0;
    c.nh = c.pbranch ? (r.ref + 1 > 0 ? r.ref + 1 : 0) : 0;
    return c;
}

// Source rectangle of hop j (VideoDecoder.cc:289-306 + loop bounds :330-333), inclusive, or the empty encoding.
__device__ __forceinline__ HopRect hop_rect(const Rec16 &r, const RecInfo &c, int j, int W, int H) {
    const float fj = (float)j;
    const float src_x = __fadd_rn((float)r.dx, __fmul_rn(__fmul_rn(c.mv_x, fj), -1.0f));  // :291-292
    const float src_y = __fadd_rn((float)r.dy, __fmul_rn(__fmul_rn(c.mv_y, fj), -1.0f));
    float s_x_top = __fsub_rn(src_x, c.half_w);
    if (s_x_top < 0) s_x_top = 0;
    float s_y_top = __fsub_rn(src_y, c.half_h);
    if (s_y_top < 0) s_y_top = 0;
    float s_x_bottom = __fadd_rn(src_x, c.half_w);
    if (s_x_bottom >= (float)W) s_x_bottom = (float)(W - 1);
    float s_y_bottom = __fadd_rn(src_y, c.half_h);
    if (s_y_bottom >= (float)H) s_y_bottom = (float)(H - 1);
    HopRect o = {0, 32767, -1, -32768};
    // for (int h = s_y_top; h <= s_y_bottom; h++): first value (int)s_y_top, runs while (float)h <= bound.
    // NaN / inf displacements (ref = -1) never reach here because nh == 0.
    if (!(s_x_top <= 40000.0f) || !(s_y_top <= 40000.0f)) return o;  // far outside: loop is empty, avoid int overflow
    const int x0 = (int)s_x_top, y0 = (int)s_y_top;
    if (!((float)x0 <= s_x_bottom) || !((float)y0 <= s_y_bottom)) return o;
    o.x0 = (int16_t)x0;
    o.y0 = (int16_t)y0;
    o.x1 = (int16_t)(int)s_x_bottom;  // bound >= x0 >= 0: truncation == floor
    o.y1 = (int16_t)(int)s_y_bottom;
    return o;
}

// ------------------------------------------------------------------------------------------------- count -----
constexpr int CNT_THREADS = 512;
constexpr int CNT_WARPS = CNT_THREADS / 32;
constexpr int MAXCLS = MOVFE_NCLS(MOVFE_MAX_K);



// A frame's records are cut into segments of `rseg` records (a multiple of CNT_THREADS); count and emit run one CTA per
// (stream, frame, segment), so a dense frame (129 600 records at 1920x1080 / 4x4) is spread over many CTAs. seg_cnt holds
// the per-segment class counts [S][n_in][n_rseg][SEG_WORDS]: classes, then area, then rejected records.
constexpr int SEG_WORDS = MAXCLS + 2;

__global__ void __launch_bounds__(CNT_THREADS)
count_kernel(WinParams p, int rseg, const Rec16 *__restrict__ d_rec, const int32_t *__restrict__ rec_cnt,
             const uint8_t *__restrict__ fflags, int32_t *__restrict__ seg_cnt) {
    __shared__ int32_t part[CNT_WARPS][MAXCLS + 2];
    const int sf = blockIdx.x;  // s*n_in + fi
    const int seg = blockIdx.y, n_rseg = gridDim.y;
    const int s = sf / p.n_in, fi = sf - s * p.n_in;
    const int slot = (int)((p.first + fi) % p.RING);
    const int ncls = MOVFE_NCLS(p.K);
    const bool mv_on = fflags[s * p.RING + slot] & MOVFE_FRAME_MV;
    const int M = mv_on ? rec_cnt[s * p.RING + slot] : 0;
    const Rec16 *recs = d_rec + ((size_t)s * p.RING + slot) * p.maxM;

    int cnt[MAXCLS];
#pragma unroll
    for (int c = 0; c < MAXCLS; c++) cnt[c] = 0;
    int ar = 0, bad = 0;
    const int i_end = min(M, (seg + 1) * rseg);
    for (int i = seg * rseg + threadIdx.x; i < i_end; i += CNT_THREADS) {
        const Rec16 r = recs[i];
        const RecInfo c = classify(r, p.W, p.H, p.K);
        bad += c.bad;
        if (!c.valid) continue;
#pragma unroll
        for (int k = 0; k <= MOVFE_MAX_K; k++)
            if (k <= p.K && c.nh > k) cnt[k]++;
        if (!c.chained) cnt[p.K + 1]++;
#pragma unroll
        for (int rr = 1; rr <= MOVFE_MAX_K; rr++)
            if (rr <= p.K && c.chained && r.ref == rr) cnt[p.K + 1 + rr]++;
        if (c.pbranch) ar += (int)r.w * (int)r.h;  // :347 coverage += dMB.area()
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < MAXCLS; c++) {
        int v = cnt[c];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) part[warp][c] = v;
    }
    for (int o = 16; o; o >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane == 0) {
        part[warp][MAXCLS] = ar;
        part[warp][MAXCLS + 1] = bad;
    }
    __syncthreads();
    if (threadIdx.x < SEG_WORDS) {
        int v = 0;  // a segment's area is at most rseg * 255 * 255 < 2^31
        for (int w = 0; w < CNT_WARPS; w++) v += part[w][threadIdx.x];
        seg_cnt[((size_t)sf * n_rseg + seg) * SEG_WORDS + threadIdx.x] = v;
    }
    (void)ncls;
    (void)fi;
}

// per-frame totals of the segment counts
__global__ void seg_sum_kernel(WinParams p, int n_rseg, const int32_t *__restrict__ seg_cnt, int32_t *__restrict__ cls_cnt,
                               int64_t *__restrict__ area, unsigned long long *__restrict__ rejected) {
    const int sf = blockIdx.x, c = threadIdx.x;
    if (c >= SEG_WORDS) return;
    long long v = 0;
    for (int seg = 0; seg < n_rseg; seg++) v += seg_cnt[((size_t)sf * n_rseg + seg) * SEG_WORDS + c];
    if (c < MAXCLS) cls_cnt[(size_t)sf * MAXCLS + c] = (int32_t)v;
    if (c == MAXCLS) area[sf] = v;
    // a look-ahead frame is counted again when it becomes an output frame: charge rejects once
    if (c == MAXCLS + 1 && v && sf % p.n_in < p.n_out) atomicAdd(rejected, (unsigned long long)v);
}

// ------------------------------------------------------------------------------------------------- bases -----
__global__ void bases_kernel(WinParams p, const int32_t *__restrict__ cls_cnt, const int64_t *__restrict__ area,
                             int32_t *__restrict__ hop_base, int32_t *__restrict__ kps_base,
                             int32_t *__restrict__ nhops, int32_t *__restrict__ nkps, double *__restrict__ cov,
                             unsigned long long *__restrict__ stats) {
    const int sg = blockIdx.x * blockDim.x + threadIdx.x;
    if (sg >= p.S * p.n_in) return;
    const int s = sg / p.n_in, g = sg - s * p.n_in;
    const int K = p.K;
    int hb = 0;
    for (int k = 0; k <= K; k++) {  // hop segment k of frame g = frame g+k's records with ref >= k (j = k+1)
        hop_base[(size_t)sg * (K + 2) + k] = hb;
        if (g + k < p.n_in) hb += cls_cnt[((size_t)s * p.n_in + g + k) * MAXCLS + k];
    }
    hop_base[(size_t)sg * (K + 2) + K + 1] = hb;
    nhops[sg] = hb < p.max_hops ? hb : p.max_hops;
    if (g < p.n_out) {  // workload counters (diagnostic)
        atomicAdd(&stats[5], (unsigned long long)(hb < p.max_hops ? hb : p.max_hops));
        atomicAdd(&stats[6], 1ull);
    }
    int kb = 0;
    kps_base[(size_t)sg * (K + 2) + 0] = 0;
    kb = cls_cnt[(size_t)sg * MAXCLS + K + 1];  // own
    for (int r = 1; r <= K; r++) {              // kps segment r = frame g+1+r's chained records with ref == r
        kps_base[(size_t)sg * (K + 2) + r] = kb;
        if (g + 1 + r < p.n_in) kb += cls_cnt[((size_t)s * p.n_in + g + 1 + r) * MAXCLS + K + 1 + r];
    }
    kps_base[(size_t)sg * (K + 2) + K + 1] = kb;
    nkps[sg] = kb < p.max_kps ? kb : p.max_kps;
    // :204,:347,:350 — float accumulation of integer areas is exact below 2^24, then / (double)(W*H)
    cov[sg] = (double)(float)area[sg] / (double)(p.W * p.H);
}

// -------------------------------------------------------------------------------------------------- emit -----
__global__ void __launch_bounds__(CNT_THREADS)
emit_kernel(WinParams p, int rseg, const int32_t *__restrict__ seg_cnt, const Rec16 *__restrict__ d_rec, const int32_t *__restrict__ rec_cnt,
            const uint8_t *__restrict__ fflags, const int32_t *__restrict__ hop_base,
            const int32_t *__restrict__ kps_base, movfe_hop *__restrict__ hops, HopRect *__restrict__ hop_rects,
            movfe_rect *__restrict__ kps) {
    __shared__ int32_t wtot[MAXCLS][CNT_WARPS];  // per-chunk: exclusive prefix over warps (after the scan step)
    __shared__ int32_t cbase[MAXCLS];            // records of each class seen in earlier chunks (and earlier segments)
    const int sf = blockIdx.x;
    const int seg = blockIdx.y, n_rseg = gridDim.y;
    const int s = sf / p.n_in, fi = sf - s * p.n_in;
    const int slot = (int)((p.first + fi) % p.RING);
    const int K = p.K, ncls = MOVFE_NCLS(K);
    const bool mv_on = fflags[s * p.RING + slot] & MOVFE_FRAME_MV;
    const int M = mv_on ? rec_cnt[s * p.RING + slot] : 0;
    const Rec16 *recs = d_rec + ((size_t)s * p.RING + slot) * p.maxM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    if (threadIdx.x < MAXCLS) {
        int v = 0;
        for (int q = 0; q < seg; q++) v += seg_cnt[((size_t)sf * n_rseg + q) * SEG_WORDS + threadIdx.x];
        cbase[threadIdx.x] = v;
    }
    __syncthreads();

    const int i_end = min(M, (seg + 1) * rseg);
    for (int base = seg * rseg; base < i_end; base += CNT_THREADS) {
        const int i = base + threadIdx.x;
        Rec16 r = {};
        RecInfo c = {};
        if (i < i_end) {
            r = recs[i];
            c = classify(r, p.W, p.H, K);
        }
        int rank_hop[MOVFE_MAX_K + 1];
#pragma unroll
        for (int k = 0; k <= MOVFE_MAX_K; k++) {
            rank_hop[k] = 0;
            if (k <= K) {
                const unsigned b = __ballot_sync(0xffffffffu, c.valid && c.nh > k);
                rank_hop[k] = __popc(b & lt);
                if (lane == 0) wtot[k][warp] = __popc(b);
            }
        }
        int rank_kps;
        {
            const unsigned b = __ballot_sync(0xffffffffu, c.valid && !c.chained);
            rank_kps = __popc(b & lt);
            if (lane == 0) wtot[K + 1][warp] = __popc(b);
        }
#pragma unroll
        for (int rr = 1; rr <= MOVFE_MAX_K; rr++) {
            if (rr <= K) {
                const bool mine = c.valid && c.chained && r.ref == rr;
                const unsigned b = __ballot_sync(0xffffffffu, mine);
                if (mine) rank_kps = __popc(b & lt);
                if (lane == 0) wtot[K + 1 + rr][warp] = __popc(b);
            }
        }
        __syncthreads();
        if (threadIdx.x < ncls) {  // exclusive scan over warps, offset by the chunks before
            int run = cbase[threadIdx.x];
            for (int w = 0; w < CNT_WARPS; w++) {
                const int t = wtot[threadIdx.x][w];
                wtot[threadIdx.x][w] = run;
                run += t;
            }
            cbase[threadIdx.x] = run;
        }
        __syncthreads();
        if (c.valid) {
            // candidate-keypoint rectangle dMB (:243-253)
            const movfe_rect dMB = {(int16_t)c.kx, (int16_t)c.ky, (int16_t)r.w, (int16_t)r.h};
            int d_indx = -1;
            if (!c.chained) {
                d_indx = wtot[K + 1][warp] + rank_kps;  // == smv->kps.size()-1 after the push
                if (fi < p.n_out && d_indx < p.max_kps)
                    kps[((size_t)s * p.n_out + fi) * p.max_kps + d_indx] = dMB;
            } else {
                const int g = fi - 1 - r.ref;  // vqueue[(size-1)-ref]
                if (g >= 0 && g < p.n_out) {
                    const int idx = kps_base[((size_t)s * p.n_in + g) * (K + 2) + r.ref] + wtot[K + 1 + r.ref][warp] + rank_kps;
                    if (idx < p.max_kps) kps[((size_t)s * p.n_out + g) * p.max_kps + idx] = dMB;
                }
            }
            // hops j = ref+1 .. 1 (:289-346): hop j lands in frame fi-(j-1), segment k = j-1
#pragma unroll
            for (int k = 0; k <= MOVFE_MAX_K; k++) {
                if (k <= K && k < c.nh) {
                    const int g = fi - k;
                    if (g >= 0 && g < p.n_out) {
                        const int idx = hop_base[((size_t)s * p.n_in + g) * (K + 2) + k] + wtot[k][warp] + rank_hop[k];
                        if (idx < p.max_hops) {
                            const size_t o = ((size_t)s * p.n_out + g) * p.max_hops + idx;
                            movfe_hop hv;
                            hv.mv_x = c.mv_x;
                            hv.mv_y = c.mv_y;
                            hv.d_indx = d_indx;
                            hv._pad = 0;
                            hops[o] = hv;
                            hop_rects[o] = hop_rect(r, c, k + 1, p.W, p.H);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------------- bbox -----
__global__ void bbox_kernel(WinParams p, const HopRect *__restrict__ hop_rects, const int32_t *__restrict__ nhops,
                            int2 *__restrict__ chunk_bbox) {
    const int sg = blockIdx.y;  // s*n_out + g
    const int s = sg / p.n_out, g = sg - s * p.n_out;
    const int n = nhops[s * p.n_in + g];
    const int nchunks = (n + 31) >> 5;
    const int lane = threadIdx.x & 31;
    const int wpg = (gridDim.x * blockDim.x) >> 5;  // warps per (stream, frame): a warp strides over the chunks
    for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += wpg) {
        const int h = chunk * 32 + lane;
        int ymin = 32767, ymax = -32768, xmin = 32767, xmax = -32768;
        if (h < n) {
            const HopRect r = hop_rects[(size_t)sg * p.max_hops + h];
            ymin = r.y0;
            ymax = r.y1;
            if (r.y1 >= r.y0) {  // empty rectangles carry an inverted extent
                xmin = r.x0;
                xmax = r.x1;
            }
        }
        for (int o = 16; o; o >>= 1) {
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        }
        if (lane == 0) chunk_bbox[(size_t)sg * p.max_chunks + chunk] = make_int2((ymin & 0xffff) | (ymax << 16), (xmin & 0xffff) | (xmax << 16));
    }
}

}  // namespace

// ----------------------------------------------------------------------------------------------- launchers -----
int movfe_ingest_launch(movfe_ctx *ctx, int n_frames, const void *d_recs, bool packed, const int64_t *d_rec_off,
                        int64_t n_records, const uint8_t *d_flags, const uint8_t *d_grey) {
    const movfe_config &c = ctx->cfg;
    const int n_seg = c.n_streams * n_frames;
    ProfScope prof(ctx, MOVFE_STAGE_INGEST, ctx->ingest_stream);
    prof.launches(n_records > 0 ? 2 : 1);
    if (n_records > 0 && packed) {
        ingest_packed_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, ctx->ingest_stream>>>(
            reinterpret_cast<const uint4 *>(d_recs), n_records, d_rec_off, n_seg, n_frames, ctx->pushed, ctx->RING, c.max_records_per_frame,
            ctx->d_rec, ctx->d_rejected);
    } else if (n_records > 0) {
        const int64_t warps = (n_records + INGEST_REC_PER_WARP - 1) / INGEST_REC_PER_WARP;
        const int blocks = (int)((warps + INGEST_WARPS - 1) / INGEST_WARPS);
        ingest_kernel<<<blocks, INGEST_WARPS * 32, 0, ctx->ingest_stream>>>(
            reinterpret_cast<const uint4 *>(d_recs), n_records, d_rec_off, n_seg, n_frames, ctx->pushed, ctx->RING,
            c.max_records_per_frame, ctx->d_rec, ctx->d_rejected);
    }
    ingest_meta_kernel<<<(n_seg + 255) / 256, 256, 0, ctx->ingest_stream>>>(d_rec_off, d_flags, n_seg, n_frames, ctx->pushed,
                                                                     ctx->RING, c.max_records_per_frame,
                                                                     ctx->d_rec_cnt, ctx->d_fflags);
    if (c.has_grey && d_grey) {
        const int vec = (c.width % 16 == 0 && ((uintptr_t)d_grey & 15) == 0) ? 16 : 1;
        const int64_t total = (int64_t)n_seg * c.height * (c.width / vec);
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
        grey_ingest_kernel<<<blocks, 256, 0, ctx->ingest_stream>>>(d_grey, n_seg, n_frames, c.width, c.height, ctx->grey_pitch, vec, ctx->pushed,
                                                          ctx->RING, ctx->d_grey);
        prof.launches(1);
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}

// one packed plane (device memory) -> ring slot `slot` of stream 0 (movfe_extract_frame)
int movfe_grey_upload(movfe_ctx *ctx, const uint8_t *d_src, int slot) {
    const movfe_config &c = ctx->cfg;
    const int vec = (c.width % 16 == 0 && ((uintptr_t)d_src & 15) == 0) ? 16 : 1;
    const int64_t total = (int64_t)c.height * (c.width / vec);
    grey_ingest_kernel<<<(int)((total + 255) / 256), 256, 0, ctx->stream>>>(d_src, 1, 1, c.width, c.height, ctx->grey_pitch, vec, slot, ctx->RING,
                                                                      ctx->d_grey);
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}

int movfe_raster_launch(movfe_ctx *ctx, RasterBuf &w, int64_t first_frame, int n_out, int n_in) {
    const movfe_config &c = ctx->cfg;
    WinParams p;
    p.S = c.n_streams;
    p.n_in = n_in;
    p.n_out = n_out;
    p.K = ctx->K;
    p.RING = ctx->RING;
    p.maxM = c.max_records_per_frame;
    p.W = c.width;
    p.H = c.height;
    p.first = first_frame;
    p.max_hops = ctx->max_hops;
    p.max_kps = ctx->max_kps;
    p.max_chunks = ctx->max_chunks;
    const int SF = p.S * n_in;
    // the hop-list kernels: on the raster stream, or (MOVFE_HOPS_PRIO) on a high-priority stream between two events
    cudaStream_t hs = ctx->hops_stream ? ctx->hops_stream : ctx->raster_stream;
    if (ctx->hops_stream) {
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_hops, ctx->raster_stream));
        MOVFE_CUDA(ctx, cudaStreamWaitEvent(hs, ctx->ev_hops, 0));
    }
    {
    ProfScope prof(ctx, MOVFE_STAGE_HOPS, hs);
    prof.launches(5);
    const dim3 gseg(SF, ctx->n_rseg);
    count_kernel<<<gseg, CNT_THREADS, 0, hs>>>(p, ctx->rseg, ctx->d_rec, ctx->d_rec_cnt, ctx->d_fflags, w.d_seg_cnt);
    seg_sum_kernel<<<SF, 32, 0, hs>>>(p, ctx->n_rseg, w.d_seg_cnt, w.d_cls_cnt, w.d_area, ctx->d_rejected);
    bases_kernel<<<(SF + 127) / 128, 128, 0, hs>>>(p, w.d_cls_cnt, w.d_area, w.d_hop_base,
                                                            w.d_kps_base, w.d_nhops, w.d_nkps, w.d_cov, ctx->d_stats);
    emit_kernel<<<gseg, CNT_THREADS, 0, hs>>>(p, ctx->rseg, w.d_seg_cnt, ctx->d_rec, ctx->d_rec_cnt, ctx->d_fflags,
                                                       w.d_hop_base, w.d_kps_base, w.d_hops, w.d_hop_rect, w.d_kps);
    {
        // capacity would be max_chunks warps per (stream, frame); frames hold a fraction of it, so a few CTAs stride instead
        dim3 g(std::min((ctx->max_chunks * 32 + 255) / 256, 16), p.S * n_out);
        bbox_kernel<<<g, 256, 0, hs>>>(p, w.d_hop_rect, w.d_nhops, w.d_chunk_bbox);
    }
    }
    if (ctx->hops_stream) {
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_hops, hs));
        MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->raster_stream, ctx->ev_hops, 0));
    }
    if (int rc = movfe_grid_launch(ctx, p, w)) return rc;
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}
