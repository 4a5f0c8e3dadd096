// lk.cu — sparse pyramidal Lucas-Kanade on the GPU: cv::calcOpticalFlowPyrLK as the reference calls it for its carry-over
// branches (src/MOVExtractor.cc:91-92 I-frame carry-over of every track, :196-197 lost relocalisation, :347-348 coverage tracks:
// winSize 31 x 31, maxLevel 3, 20 iterations / eps 0.01, OPTFLOW_LK_GET_MIN_EIGENVALS, minEigThreshold 1e-4; src/Frame.cc:305: 21 x 21).
// SURVEY.md 8f item 3. The arithmetic is OpenCV's (modules/video/src/lkpyramid.cpp, modules/imgproc/src/pyramids.cpp of OpenCV
// 4.6.0, un-vendored): restated in oracle/lk.py, which is pinned to OpenCV's own outputs; this file follows the same steps -
//   pyr_down_kernel   [1 4 6 4 1] x [1 4 6 4 1], BORDER_REFLECT_101, exact integer sum, (sum + 128) >> 8
//   scharr_kernel     int16 derivatives of every level of the first image (3/10/3 across, central difference along)
//   lk_track_kernel   a WARP per point walks the levels coarse to fine: the window's I, Ix, Iy (14-bit fixed-point bilinear
//                     weights, image border reflected, derivative border zero) stay in shared memory as int16, the 2x2 normal
//                     matrix and every iteration's right-hand side are warp-shuffle sums of per-lane float partials
// so pyramids and derivatives are bit-exact and positions agree with OpenCV to ~1e-3 px (the window sums are formed in a
// different order than OpenCV's SIMD path: last float bits). Compiled with -fmad=false like the rest of the library.
#include <algorithm>
#include <cstdio>

#include "common.cuh"

int movfe_ensure_op_scratch(movfe_ctx *ctx, size_t bytes);

namespace {

constexpr int LK_MAX_LEVELS = 8;
constexpr int LK_W_BITS = 14;

struct LkLevels {
    int n_levels;                  // usable levels (cv::buildOpticalFlowPyramid stops before a level no larger than the window)
    int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
    size_t off[LK_MAX_LEVELS];     // pixel offset of a level inside one image's pyramid
    size_t px;                     // pixels of one pyramid
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    int m = i % period;
    if (m < 0) m += period;
    return m >= n ? period - m : m;
}

// level l of every image from level l-1: one thread per output pixel
__global__ void pyr_down_kernel(uint8_t *__restrict__ pyr, size_t img_px, size_t src_off, size_t dst_off, int sw, int sh, int dw, int dh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const uint8_t *src = pyr + (size_t)blockIdx.z * img_px + src_off;
    const int k[5] = {1, 4, 6, 4, 1};
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint8_t *row = src + (size_t)reflect101(2 * y + j - 2, sh) * sw;
        int r = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) r += k[i] * (int)row[reflect101(2 * x + i - 2, sw)];
        sum += k[j] * r;
    }
    pyr[(size_t)blockIdx.z * img_px + dst_off + (size_t)y * dw + x] = (uint8_t)((sum + 128) >> 8);
}

// calcSharrDeriv of one level of the FIRST image of every pair: dxy[2 * pixel] = dx, [2 * pixel + 1] = dy
__global__ void scharr_kernel(const uint8_t *__restrict__ pyr, size_t img_stride_px, size_t off, int w, int h, int16_t *__restrict__ dxy, size_t d_stride_px) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t *a = pyr + (size_t)blockIdx.z * img_stride_px + off;
    const int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h), xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    auto at = [&](int yy, int xx) { return (int)a[(size_t)yy * w + xx]; };
    const int t0p = (at(ym, xp) + at(yp, xp)) * 3 + at(y, xp) * 10, t0m = (at(ym, xm) + at(yp, xm)) * 3 + at(y, xm) * 10;
    const int t1p = at(yp, xp) - at(ym, xp), t1m = at(yp, xm) - at(ym, xm), t1c = at(yp, x) - at(ym, x);
    int16_t *o = dxy + ((size_t)blockIdx.z * d_stride_px + off + (size_t)y * w + x) * 2;
    o[0] = (int16_t)(t0p - t0m);
    o[1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void lk_weights(float a, float b, int &w00, int &w01, int &w10, int &w11) {
    const float s = (float)(1 << LK_W_BITS);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), s));  // cvRound: round half to even
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), s));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), s));
    w11 = (1 << LK_W_BITS) - w00 - w01 - w10;
}

constexpr int LK_WARPS = 4;

// cv::detail::LKTrackerInvoker::operator() for one point per warp. prev_pyr / next_pyr: [problem][pyramid]; dxy: [problem][pyramid][2].
// Points: packed with offsets (off != nullptr: movfe_lk), or `slots` per problem of which the first counts[problem] are used
// (blockIdx.y = problem: movfe_lk_carry, where a problem is a stream and the arrays are the LK hand-over buffers).
template <int WIN>
__global__ void __launch_bounds__(LK_WARPS * 32)
lk_track_kernel(LkLevels lv, const uint8_t *__restrict__ prev_pyr, const uint8_t *__restrict__ next_pyr, const int16_t *__restrict__ dxy,
                const float2 *__restrict__ pts, const int32_t *__restrict__ off, int n_problems, const int32_t *__restrict__ counts, int slots,
                int max_count, double eps2, float min_eig_thr, float2 *__restrict__ out, uint8_t *__restrict__ status, float *__restrict__ err) {
    constexpr int NPX = WIN * WIN;
    __shared__ int16_t sI[LK_WARPS][NPX], sIx[LK_WARPS][NPX], sIy[LK_WARPS][NPX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int pi = blockIdx.x * LK_WARPS + warp;
    int prob = 0;
    if (off) {  // problem of this point: last problem with off[prob] <= pi
        if (pi >= off[n_problems]) return;
        int lo = 0, hi = n_problems;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (off[mid] <= pi) lo = mid; else hi = mid;
        }
        prob = lo;
    } else {
        prob = blockIdx.y;
        if (pi >= counts[prob]) return;
        pi += prob * slots;
    }
    const uint8_t *P = prev_pyr + (size_t)prob * lv.px, *N = next_pyr + (size_t)prob * lv.px;
    const int16_t *D = dxy + (size_t)prob * lv.px * 2;
    const float2 pt = pts[pi];
    const float half = (float)((WIN - 1) * 0.5);
    float nx = 0.f, ny = 0.f, ox = 0.f, oy = 0.f;
    bool ok = true;
    float e0 = 0.f;
    for (int level = lv.n_levels - 1; level >= 0; level--) {
        const int w = lv.w[level], h = lv.h[level];
        const uint8_t *I = P + lv.off[level], *J = N + lv.off[level];
        const int16_t *dI = D + lv.off[level] * 2;
        const float sc = 1.f / (float)(1 << level);
        float px = __fmul_rn(pt.x, sc), py = __fmul_rn(pt.y, sc);
        if (level == lv.n_levels - 1) {
            nx = px;
            ny = py;
        } else {
            nx = __fmul_rn(nx, 2.f);
            ny = __fmul_rn(ny, 2.f);
        }
        ox = nx;
        oy = ny;
        px = px - half;
        py = py - half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -WIN || ipx >= w || ipy < -WIN || ipy >= h) {
            if (level == 0) {
                ok = false;
                e0 = 0.f;
            }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(px - (float)ipx, py - (float)ipy, w00, w01, w10, w11);
        float a11 = 0.f, a12 = 0.f, a22 = 0.f;
        for (int q = lane; q < NPX; q += 32) {
            const int y = q / WIN, x = q - y * WIN;
            const int gy0 = ipy + y, gx0 = ipx + x;
            const int y0 = reflect101(gy0, h), y1 = reflect101(gy0 + 1, h), x0 = reflect101(gx0, w), x1 = reflect101(gx0 + 1, w);
            const int iv = ((int)I[(size_t)y0 * w + x0] * w00 + (int)I[(size_t)y0 * w + x1] * w01 + (int)I[(size_t)y1 * w + x0] * w10 +
                            (int)I[(size_t)y1 * w + x1] * w11 + (1 << (LK_W_BITS - 5 - 1))) >> (LK_W_BITS - 5);
            // the derivative image has a ZERO border (copyMakeBorder BORDER_CONSTANT), the image a reflected one
            const bool iy0 = gy0 >= 0 && gy0 < h, iy1 = gy0 + 1 >= 0 && gy0 + 1 < h, ix0 = gx0 >= 0 && gx0 < w, ix1 = gx0 + 1 >= 0 && gx0 + 1 < w;
            const int2 zero = make_int2(0, 0);
            auto dv = [&](bool in, int yy, int xx) {
                if (!in) return zero;
                const short2 v = *reinterpret_cast<const short2 *>(dI + ((size_t)yy * w + xx) * 2);
                return make_int2((int)v.x, (int)v.y);
            };
            const int2 d00 = dv(iy0 && ix0, y0, x0), d01 = dv(iy0 && ix1, y0, x1), d10 = dv(iy1 && ix0, y1, x0), d11 = dv(iy1 && ix1, y1, x1);
            const int ixv = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
            const int iyv = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
            sI[warp][q] = (int16_t)iv;
            sIx[warp][q] = (int16_t)ixv;
            sIy[warp][q] = (int16_t)iyv;
            a11 += (float)(ixv * ixv);
            a12 += (float)(ixv * iyv);
            a22 += (float)(iyv * iyv);
        }
        const float FLT_SCALE = 1.f / (float)(1 << 20);
        const float A11 = __fmul_rn(warp_sum(a11), FLT_SCALE), A12 = __fmul_rn(warp_sum(a12), FLT_SCALE), A22 = __fmul_rn(warp_sum(a22), FLT_SCALE);
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dif = A11 - A22;
        const float min_eig = __fdiv_rn(A22 + A11 - __fsqrt_rn(__fadd_rn(__fmul_rn(dif, dif), __fmul_rn(__fmul_rn(4.f, A12), A12))), (float)(2 * WIN * WIN));
        if (level == 0) e0 = min_eig;
        if (min_eig < min_eig_thr || Dt < 1.1920929e-07f) {
            if (level == 0) ok = false;
            continue;
        }
        Dt = __fdiv_rn(1.f, Dt);
        nx = nx - half;
        ny = ny - half;
        float pdx = 0.f, pdy = 0.f;
        __syncwarp();
        for (int j = 0; j < max_count; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
                if (level == 0) ok = false;
                break;
            }
            lk_weights(nx - (float)inx, ny - (float)iny, w00, w01, w10, w11);
            float b1 = 0.f, b2 = 0.f;
            for (int q = lane; q < NPX; q += 32) {
                const int y = q / WIN, x = q - y * WIN;
                const int y0 = reflect101(iny + y, h), y1 = reflect101(iny + y + 1, h), x0 = reflect101(inx + x, w), x1 = reflect101(inx + x + 1, w);
                const int jv = ((int)J[(size_t)y0 * w + x0] * w00 + (int)J[(size_t)y0 * w + x1] * w01 + (int)J[(size_t)y1 * w + x0] * w10 +
                                (int)J[(size_t)y1 * w + x1] * w11 + (1 << (LK_W_BITS - 5 - 1))) >> (LK_W_BITS - 5);
                const int diff = jv - (int)sI[warp][q];
                b1 += (float)(diff * (int)sIx[warp][q]);
                b2 += (float)(diff * (int)sIy[warp][q]);
            }
            const float B1 = __fmul_rn(warp_sum(b1), FLT_SCALE), B2 = __fmul_rn(warp_sum(b2), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, B2), __fmul_rn(A22, B1)), Dt);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, B1), __fmul_rn(A11, B2)), Dt);
            nx = nx + dx;
            ny = ny + dy;
            ox = nx + half;
            oy = ny + half;
            if ((double)dx * (double)dx + (double)dy * (double)dy <= eps2) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                ox = ox - __fmul_rn(dx, 0.5f);
                oy = oy - __fmul_rn(dy, 0.5f);
                break;
            }
            pdx = dx;
            pdy = dy;
        }
        nx = ox;
        ny = oy;
        __syncwarp();  // the window arrays are rewritten at the next level
    }
    if (lane == 0) {
        out[pi] = make_float2(ox, oy);
        status[pi] = ok ? 1 : 0;
        if (err) err[pi] = e0;
    }
}

// ---- device-resident carry-over (movfe_lk_carry) -----------------------------------------------------------------------------
// The points the reference hands to calcOpticalFlowPyrLK before frame f: at an intra picture every track of the previous table in
// TABLE order (src/MOVExtractor.cc:85-90), at a P picture the coverage tracks in SORTED order (:337-346; the i-th one owns result i).
constexpr int LKP_THREADS = 1024;
__global__ void __launch_bounds__(LKP_THREADS)
lk_points_kernel(const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks, const uint16_t *__restrict__ order,
                 const uint8_t *__restrict__ fflags, int TSLOTS, int tslot_prev, int RING, int gslot, int maxT, float2 *__restrict__ pts,
                 int32_t *__restrict__ n_out) {
    __shared__ int wsum[LKP_THREADS / 32];
    __shared__ int s_total;
    const int s = blockIdx.x;
    const movfe_track *prev = tracks + ((size_t)s * TSLOTS + tslot_prev) * maxT;
    const int n_prev = ntracks[s * TSLOTS + tslot_prev];
    const uint16_t *ord = order + (size_t)s * maxT;
    float2 *out = pts + (size_t)s * maxT;
    const bool is_p = fflags[s * RING + gslot] & MOVFE_FRAME_P;
    if (!is_p) {
        for (int i = threadIdx.x; i < n_prev; i += LKP_THREADS) out[i] = *reinterpret_cast<const float2 *>(&prev[i].pt_x);
        if (threadIdx.x == 0) n_out[s] = n_prev;
        return;
    }
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int base = 0; base < n_prev; base += LKP_THREADS) {
        const int r = base + threadIdx.x;
        int t = 0;
        bool cov = false;
        if (r < n_prev) {
            t = ord[r];
            cov = (prev[t].flags & MOVFE_TRACK_COVERAGE) != 0;
        }
        const unsigned b = __ballot_sync(0xffffffffu, cov);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int before = s_total, tot = 0;
        for (int w = 0; w < LKP_THREADS / 32; w++) {
            const int c = wsum[w];
            before += w < warp ? c : 0;
            tot += c;
        }
        if (cov) out[before + __popc(b & ((1u << lane) - 1u))] = *reinterpret_cast<const float2 *>(&prev[t].pt_x);
        __syncthreads();
        if (threadIdx.x == 0) s_total += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_out[s] = s_total;
}

// level 0 of the pyramids of two ring frames of every stream: rows of the pitched grey ring -> dense rows
__global__ void lk_level0_kernel(const uint8_t *__restrict__ ring, int RING, int slot_prev, int slot_next, int W, int H, int pitch, int S,
                                 uint8_t *__restrict__ pyr, size_t img_px) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, z = blockIdx.z;  // z: [prev of all streams][next of all streams]
    if (x >= W) return;
    const int s = z % S, slot = z < S ? slot_prev : slot_next;
    const uint8_t *src = ring + (((size_t)s * RING + slot) * H + y) * pitch + x;
    uint8_t *dst = pyr + (size_t)z * img_px + (size_t)y * W + x;
    if (x + 4 <= W && (W & 3) == 0) *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
    else
        for (int k = 0; k < 4 && x + k < W; k++) dst[k] = src[k];
}

struct LkCarve {
    uint8_t *base;
    size_t off = 0;
    template <typename T>
    T *take(size_t n) {
        T *p = (T *)(base + off);
        off += (n * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
};

}  // namespace

extern "C" int movfe_lk(movfe_ctx *ctx, int n_problems, const uint8_t *prev, const uint8_t *next, int stride, const float *pts_xy, const int32_t *off,
                        int win_size, int max_level, int max_count, double epsilon, double min_eig_threshold, float *out_xy, uint8_t *status,
                        float *err) {
    if (!ctx || n_problems < 1 || !prev || !next || !off || !out_xy || !status || !err) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    const int W = c.width, H = c.height;
    if (stride == 0) stride = W;
    const int n = off[n_problems];
    if (n < 0 || (n > 0 && !pts_xy) || stride < W) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "lk: bad argument");
    if (win_size != 21 && win_size != 31) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "lk: window %d (the reference uses 31 x 31 and 21 x 21)", win_size);
    if (max_level < 0 || max_level >= LK_MAX_LEVELS) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "lk: max_level %d outside [0, %d)", max_level, LK_MAX_LEVELS);
    max_count = std::min(std::max(max_count, 0), 100);  // as cv::calcOpticalFlowPyrLK clamps its criteria
    epsilon = std::min(std::max(epsilon, 0.), 10.);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    LkLevels lv = {};
    lv.w[0] = W;
    lv.h[0] = H;
    lv.off[0] = 0;
    lv.n_levels = 1;
    size_t px = (size_t)W * H;
    for (int l = 1; l <= max_level; l++) {
        const int w = (lv.w[l - 1] + 1) / 2, h = (lv.h[l - 1] + 1) / 2;
        if (w <= win_size || h <= win_size) break;  // cv::buildOpticalFlowPyramid
        lv.w[l] = w;
        lv.h[l] = h;
        lv.off[l] = px;
        px += (size_t)w * h;
        lv.n_levels = l + 1;
    }
    lv.px = px;
    const size_t need = 2 * (size_t)n_problems * px + (size_t)n_problems * px * 4 + (size_t)n * 13 + (size_t)(n_problems + 1) * 4 + 8 * 256;
    int rc = movfe_ensure_op_scratch(ctx, need);
    if (rc) return rc;
    LkCarve cv{(uint8_t *)ctx->d_op};
    uint8_t *d_pyr = cv.take<uint8_t>(2 * (size_t)n_problems * px);  // [prev pyramids of all problems][next pyramids of all problems]
    int16_t *d_dxy = cv.take<int16_t>((size_t)n_problems * px * 2);
    float2 *d_pts = cv.take<float2>((size_t)std::max(n, 1)), *d_out = cv.take<float2>((size_t)std::max(n, 1));
    int32_t *d_off = cv.take<int32_t>((size_t)n_problems + 1);
    uint8_t *d_status = cv.take<uint8_t>((size_t)std::max(n, 1));
    float *d_err = cv.take<float>((size_t)std::max(n, 1));
    cudaStream_t st = ctx->stream;
    // level 0 of both pyramids: the images as they are (rows `stride` bytes apart on the host)
    for (int i = 0; i < n_problems; i++) {
        MOVFE_CUDA(ctx, cudaMemcpy2DAsync(d_pyr + (size_t)i * px, (size_t)W, prev + (size_t)i * stride * H, (size_t)stride, (size_t)W, (size_t)H,
                                          cudaMemcpyHostToDevice, st));
        MOVFE_CUDA(ctx, cudaMemcpy2DAsync(d_pyr + ((size_t)n_problems + i) * px, (size_t)W, next + (size_t)i * stride * H, (size_t)stride, (size_t)W,
                                          (size_t)H, cudaMemcpyHostToDevice, st));
    }
    if (n > 0) MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pts, pts_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    for (int l = 1; l < lv.n_levels; l++)
        pyr_down_kernel<<<dim3((lv.w[l] + 127) / 128, lv.h[l], 2 * n_problems), 128, 0, st>>>(d_pyr, px, lv.off[l - 1], lv.off[l], lv.w[l - 1], lv.h[l - 1],
                                                                                                lv.w[l], lv.h[l]);
    for (int l = 0; l < lv.n_levels; l++)
        scharr_kernel<<<dim3((lv.w[l] + 127) / 128, lv.h[l], n_problems), 128, 0, st>>>(d_pyr, px, lv.off[l], lv.w[l], lv.h[l], d_dxy, px);
    if (n > 0) {
        const dim3 grid((n + LK_WARPS - 1) / LK_WARPS);
        const double eps2 = epsilon * epsilon;
        if (win_size == 31)
            lk_track_kernel<31><<<grid, LK_WARPS * 32, 0, st>>>(lv, d_pyr, d_pyr + (size_t)n_problems * px, d_dxy, d_pts, d_off, n_problems, nullptr, 0, max_count, eps2,
                                                                 (float)min_eig_threshold, d_out, d_status, d_err);
        else
            lk_track_kernel<21><<<grid, LK_WARPS * 32, 0, st>>>(lv, d_pyr, d_pyr + (size_t)n_problems * px, d_dxy, d_pts, d_off, n_problems, nullptr, 0, max_count, eps2,
                                                                 (float)min_eig_threshold, d_out, d_status, d_err);
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
    if (n > 0) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(out_xy, d_out, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, st));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(status, d_status, (size_t)n, cudaMemcpyDeviceToHost, st));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(err, d_err, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
    return MOVFE_OK;
}

// Device-resident carry-over for the batched path: Lucas-Kanade results for frame `frame` of every stream, from the grey planes of
// frame - 1 and frame in the ring and the track table of frame - 1, installed as movfe_set_lk_results would install them. Call it
// between movfe_extract(.., frame - 1) and movfe_extract(frame, ..) - at an intra picture in mid-stream (every track is carried,
// src/MOVExtractor.cc:81-120) or when coverage tracks are to be followed (:337-377). Nothing crosses PCIe.
extern "C" int movfe_lk_carry(movfe_ctx *ctx, int64_t frame) {
    if (!ctx) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    if (!c.has_grey) MOVFE_FAIL(ctx, MOVFE_E_STATE, "lk_carry: the context has no grey planes");
    const int64_t next = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    if (frame != next || frame < 1) MOVFE_FAIL(ctx, MOVFE_E_STATE, "lk_carry: frame %lld is not the next frame to be propagated (%lld)", (long long)frame, (long long)next);
    if (frame >= ctx->pushed || ctx->pushed - (frame - 1) > ctx->RING)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "lk_carry: frames %lld and %lld are not both in the ring (pushed=%lld, ring=%d)", (long long)(frame - 1), (long long)frame,
                   (long long)ctx->pushed, ctx->RING);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    const int W = c.width, H = c.height, S = c.n_streams, win = 31, max_level = 3;
    LkLevels lv = {};
    lv.w[0] = W;
    lv.h[0] = H;
    lv.n_levels = 1;
    size_t px = (size_t)W * H;
    for (int l = 1; l <= max_level; l++) {
        const int w = (lv.w[l - 1] + 1) / 2, h = (lv.h[l - 1] + 1) / 2;
        if (w <= win || h <= win) break;
        lv.w[l] = w;
        lv.h[l] = h;
        lv.off[l] = px;
        px += (size_t)w * h;
        lv.n_levels = l + 1;
    }
    lv.px = px;
    const size_t need = 2 * (size_t)S * px + (size_t)S * px * 4 + (size_t)S * c.max_tracks * sizeof(float2) + 4 * 256;
    cudaStream_t st = ctx->stream;
    if (ctx->lk_scratch_bytes < need) {
        MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->d_lk_scratch) cudaFree(ctx->d_lk_scratch);
        ctx->d_lk_scratch = nullptr;
        ctx->lk_scratch_bytes = 0;
        MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_lk_scratch, need));
        ctx->lk_scratch_bytes = need;
    }
    LkCarve cv{(uint8_t *)ctx->d_lk_scratch};
    uint8_t *d_pyr = cv.take<uint8_t>(2 * (size_t)S * px);
    int16_t *d_dxy = cv.take<int16_t>((size_t)S * px * 2);
    float2 *d_in = cv.take<float2>((size_t)S * c.max_tracks);
    movfe_lk_handover hb;
    movfe_lk_buffers(ctx, &hb);
    // the raster stream wrote the ring (ingest): the newest push must have landed before the planes are read
    MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_tables, ctx->ingest_stream));
    MOVFE_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_tables, 0));
    const int T = ctx->TSLOTS, ts_prev = (int)((((frame - 1) % T) + T) % T);
    const int slot_prev = (int)((frame - 1) % ctx->RING), slot_next = (int)(frame % ctx->RING);
    lk_points_kernel<<<S, LKP_THREADS, 0, st>>>(ctx->d_tracks, ctx->d_ntracks, hb.order, ctx->d_fflags, T, ts_prev, ctx->RING, slot_next, c.max_tracks, d_in, hb.n);
    lk_level0_kernel<<<dim3((W / 4 + 127) / 128 + 1, H, 2 * S), 128, 0, st>>>(ctx->d_grey, ctx->RING, slot_prev, slot_next, W, H, ctx->grey_pitch, S, d_pyr, px);
    for (int l = 1; l < lv.n_levels; l++)
        pyr_down_kernel<<<dim3((lv.w[l] + 127) / 128, lv.h[l], 2 * S), 128, 0, st>>>(d_pyr, px, lv.off[l - 1], lv.off[l], lv.w[l - 1], lv.h[l - 1], lv.w[l], lv.h[l]);
    for (int l = 0; l < lv.n_levels; l++)
        scharr_kernel<<<dim3((lv.w[l] + 127) / 128, lv.h[l], S), 128, 0, st>>>(d_pyr, px, lv.off[l], lv.w[l], lv.h[l], d_dxy, px);
    lk_track_kernel<31><<<dim3((c.max_tracks + LK_WARPS - 1) / LK_WARPS, S), LK_WARPS * 32, 0, st>>>(
        lv, d_pyr, d_pyr + (size_t)S * px, d_dxy, d_in, nullptr, S, hb.n, c.max_tracks, 20, 0.01 * 0.01, 1e-4f, reinterpret_cast<float2 *>(hb.pts), hb.status,
        nullptr);
    MOVFE_CUDA(ctx, cudaGetLastError());
    ctx->lk_pending = true;
    return MOVFE_OK;
}
