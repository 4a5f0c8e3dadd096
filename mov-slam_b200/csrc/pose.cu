// pose.cu — frustum projection, track-id joins and the pose-only Gauss-Newton / Huber solver.
// Replaces Frame::isInFrustum (src/Frame.cc:456-519) + Pinhole::project (src/CameraModels/Pinhole.cpp:45-52),
// MOVMatcher::SearchByVideoFeature x2 / SearchForInitialization (include/MOVMatcher.h:35-137) and
// Optimizer::PoseOptimization (include/Optimizer.h:55) with the residual / Jacobian of
// EdgeSE3ProjectXYZOnlyPose (include/OptimizableTypes.h:41-46, src/OptimizableTypes.cpp:54-69).
// Compiled with -fmad=false: the frustum's binary32 arithmetic must round like the reference's.
//
// Streams are independent, and inside a stream the chain  join(KF) -> pose -> frustum -> join(local) -> pose
// (Tracking.cc:796-811, 890-905, 1109-1158) is sequential over frames, so the batched driver is ONE persistent
// kernel with one CTA per stream that walks the window's frames: no grid-wide sync, no collective, no launch per
// step. The 27 normal-equation sums (21 JtJ + 6 Jtr) are reduced with a fixed warp-shuffle tree + a fixed
// shared-memory order, so results are run-to-run identical; no atomics touch floating-point data.
#include <algorithm>
#include <cstdio>

#include <cooperative_groups.h>

#include "common.cuh"

namespace {

#ifndef MOVFE_TP_THREADS
#define MOVFE_TP_THREADS 256
#endif
constexpr int TP_THREADS = MOVFE_TP_THREADS;  // 256 x 128 registers: half an SM's register file, so propagation CTAs co-reside
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int HASH_EMPTY = (int)0x80000000;

// ------------------------------------------------------------------------------------------------ camera -----
struct CamD {
    int model;
    double fx, fy, cx, cy, k0, k1, k2, k3;
};

__device__ __forceinline__ CamD widen(const movfe_camera &c) {
    CamD r;
    r.model = c.model;
    r.fx = c.fx;
    r.fy = c.fy;
    r.cx = c.cx;
    r.cy = c.cy;
    r.k0 = c.k[0];
    r.k1 = c.k[1];
    r.k2 = c.k[2];
    r.k3 = c.k[3];
    return r;
}

// Pinhole.cpp:37-43 / KannalaBrandt8 (ORB-SLAM3 lineage), double
__device__ __forceinline__ void project_d(const CamD &c, double x, double y, double z, double &u, double &v) {
    if (c.model == MOVFE_CAM_FISHEYE) {
        const double r = sqrt(x * x + y * y);
        const double theta = atan2(r, z);
        const double t2 = theta * theta;
        const double thetad = theta * (1.0 + t2 * (c.k0 + t2 * (c.k1 + t2 * (c.k2 + t2 * c.k3))));
        const double s = r > 1e-12 ? thetad / r : 1.0;
        u = c.fx * s * x + c.cx;
        v = c.fy * s * y + c.cy;
    } else {
        u = c.fx * x / z + c.cx;
        v = c.fy * y / z + c.cy;
    }
}

// Pinhole.cpp:77-88 / KB8 Jacobian, 2x3 row-major
__device__ __forceinline__ void project_jac_d(const CamD &c, double x, double y, double z, double J[6]) {
    if (c.model == MOVFE_CAM_FISHEYE) {
        const double r2 = x * x + y * y;
        const double r = sqrt(r2);
        if (r >= 1e-8) {
            const double theta = atan2(r, z);
            const double t2 = theta * theta;
            const double f = theta * (1.0 + t2 * (c.k0 + t2 * (c.k1 + t2 * (c.k2 + t2 * c.k3))));
            const double fd = 1.0 + t2 * (3 * c.k0 + t2 * (5 * c.k1 + t2 * (7 * c.k2 + t2 * 9 * c.k3)));
            const double D = r2 + z * z;
            const double r3 = r2 * r;
            J[0] = c.fx * (fd * z * x * x / (r2 * D) + f * y * y / r3);
            J[1] = c.fx * (fd * z * x * y / (r2 * D) - f * x * y / r3);
            J[2] = -c.fx * fd * x / D;
            J[3] = c.fy * (fd * z * x * y / (r2 * D) - f * x * y / r3);
            J[4] = c.fy * (fd * z * y * y / (r2 * D) + f * x * x / r3);
            J[5] = -c.fy * fd * y / D;
            return;
        }
    }
    J[0] = c.fx / z;
    J[1] = 0;
    J[2] = -c.fx * x / (z * z);
    J[3] = 0;
    J[4] = c.fy / z;
    J[5] = -c.fy * y / (z * z);
}

// project_d + project_jac_d of the fisheye model in one go: radius, angle and polynomial are shared (sqrt and atan2 are most
// of the per-point cost of a KannalaBrandt8 solve). The same expressions, so the same bits as the two separate calls.
__device__ __forceinline__ void project_with_jac_kb8(const CamD &c, double x, double y, double z, double &u, double &v, double J[6]) {
    const double r2 = x * x + y * y;
    const double r = sqrt(r2);
    const double theta = atan2(r, z);
    const double t2 = theta * theta;
    const double f = theta * (1.0 + t2 * (c.k0 + t2 * (c.k1 + t2 * (c.k2 + t2 * c.k3))));
    const double s = r > 1e-12 ? f / r : 1.0;
    u = c.fx * s * x + c.cx;
    v = c.fy * s * y + c.cy;
    if (r >= 1e-8) {
        const double fd = 1.0 + t2 * (3 * c.k0 + t2 * (5 * c.k1 + t2 * (7 * c.k2 + t2 * 9 * c.k3)));
        const double D = r2 + z * z;
        const double r3 = r2 * r;
        J[0] = c.fx * (fd * z * x * x / (r2 * D) + f * y * y / r3);
        J[1] = c.fx * (fd * z * x * y / (r2 * D) - f * x * y / r3);
        J[2] = -c.fx * fd * x / D;
        J[3] = c.fy * (fd * z * x * y / (r2 * D) - f * x * y / r3);
        J[4] = c.fy * (fd * z * y * y / (r2 * D) + f * x * x / r3);
        J[5] = -c.fy * fd * y / D;
    } else {  // on the optical axis: the pinhole Jacobian (project_jac_d)
        J[0] = c.fx / z;
        J[1] = 0;
        J[2] = -c.fx * x / (z * z);
        J[3] = 0;
        J[4] = c.fy / z;
        J[5] = -c.fy * y / (z * z);
    }
}

// ------------------------------------------------------------------------------------------------ frustum ----
__device__ __forceinline__ float dot3f(const float a[3], const float b[3]) {
    // Eigen 3.4's unrolled 3-vector reduction: c0 + (c1 + c2) (see oracle/match.cc)
    return __fadd_rn(__fmul_rn(a[0], b[0]), __fadd_rn(__fmul_rn(a[1], b[1]), __fmul_rn(a[2], b[2])));
}

struct FrustumPose {
    float R[9], t[3], Ow[3];
};

__device__ __forceinline__ FrustumPose frustum_pose(const movfe_pose &T) {
    FrustumPose f;
    for (int i = 0; i < 9; i++) f.R[i] = (float)T.R[i];
    for (int i = 0; i < 3; i++) f.t[i] = (float)T.t[i];
    for (int i = 0; i < 3; i++)
        f.Ow[i] = (float)(-(T.R[0 * 3 + i] * T.t[0] + T.R[1 * 3 + i] * T.t[1] + T.R[2 * 3 + i] * T.t[2]));
    return f;
}

// Frame::isInFrustum, mono branch (Frame.cc:458-519); `skip` = mnLastFrameSeen == current frame (Tracking.cc:1136)
__device__ __forceinline__ movfe_projection frustum_point(const FrustumPose &fp, const movfe_camera &cam, int W, int H,
                                                          float cosLimit, const movfe_map_point &mp, bool skip) {
    movfe_projection o;
    o.in_view = 0;
    o.u = -1.f;
    o.v = -1.f;
    o.depth = 0.f;
    o.view_cos = 0.f;
    if (skip || (mp.flags & (MOVFE_MP_BAD | MOVFE_MP_SKIP | MOVFE_MP_NULL))) return o;
    float Pc[3];
    for (int i = 0; i < 3; i++) Pc[i] = __fadd_rn(dot3f(&fp.R[3 * i], mp.pos), fp.t[i]);  // :468
    const float Pc_dist = __fsqrt_rn(dot3f(Pc, Pc));                                       // :469
    if (Pc[2] < 0.0f) return o;                                                            // :474
    float u, v;
    if (cam.model == MOVFE_CAM_FISHEYE) {
        double ud, vd;
        project_d(widen(cam), (double)Pc[0], (double)Pc[1], (double)Pc[2], ud, vd);
        u = (float)ud;
        v = (float)vd;
    } else {  // Pinhole.cpp:48-49
        u = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fx, Pc[0]), Pc[2]), cam.cx);
        v = __fadd_rn(__fdiv_rn(__fmul_rn(cam.fy, Pc[1]), Pc[2]), cam.cy);
    }
    if (u < 0.0f || u > (float)W) return o;  // :479-482 (mnMinX = 0, mnMaxX = cols)
    if (v < 0.0f || v > (float)H) return o;
    o.u = u;
    o.v = v;
    const float maxD = __fmul_rn(1.2f, mp.max_dist), minD = __fmul_rn(0.8f, mp.min_dist);  // MapPoint.cc:443-453
    const float PO[3] = {__fsub_rn(mp.pos[0], fp.Ow[0]), __fsub_rn(mp.pos[1], fp.Ow[1]), __fsub_rn(mp.pos[2], fp.Ow[2])};
    const float dist = __fsqrt_rn(dot3f(PO, PO));
    if (dist < minD || dist > maxD) return o;  // :493
    const float viewCos = __fdiv_rn(dot3f(PO, mp.normal), dist);  // :499
    if (viewCos < cosLimit) return o;
    o.in_view = 1;
    o.depth = Pc_dist;
    o.view_cos = viewCos;
    return o;
}

// ------------------------------------------------------------------------------------------------ joins ------
__device__ __forceinline__ unsigned hash_id(int id) { return (unsigned)id * 2654435761u; }

// F.mvVFMap: first index per track id (std::map::insert never overwrites, MOVExtractor.cc:330).
// keys/vals: cap entries of shared memory, cap = power of two >= 2n. CTA-cooperative.
template <typename IdFn>
__device__ void hash_build(IdFn id_of, int n, int *keys, int *vals, int cap) {
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        keys[i] = HASH_EMPTY;
        vals[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int id = id_of(t);
        unsigned h = hash_id(id) & (cap - 1);
        while (true) {
            const int prev = atomicCAS(&keys[h], HASH_EMPTY, id);
            if (prev == HASH_EMPTY || prev == id) {
                atomicMin(&vals[h], t);  // integer min: first index wins whatever the thread order
                break;
            }
            h = (h + 1) & (cap - 1);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int hash_find(int id, const int *keys, const int *vals, int cap) {
    if (id == HASH_EMPTY) return -1;
    unsigned h = hash_id(id) & (cap - 1);
    while (true) {
        const int k = keys[h];
        if (k == id) return vals[h];
        if (k == HASH_EMPTY) return -1;
        h = (h + 1) & (cap - 1);
    }
}

// ------------------------------------------------------------------------------------------------ solver -----
// g2o SE3Quat::exp, update = [omega, upsilon]
__device__ void se3_exp_d(const double dx[6], double R[9], double t[3]) {
    const double wx = dx[0], wy = dx[1], wz = dx[2];
    const double theta = sqrt(wx * wx + wy * wy + wz * wz);
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    double a, b, c;
    if (theta < 0.00001) {
        a = 1.0;
        b = 0.5;
        c = 1.0 / 6.0;
    } else {
        double sn, cs;
        sincos(theta, &sn, &cs);
        const double it = 1.0 / theta;
        a = sn * it;
        b = (1 - cs) * it * it;
        c = (theta - sn) * it * it * it;
    }
    double V[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = I + a * O[i] + b * O2[i];
        V[i] = I + b * O[i] + c * O2[i];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = V[i * 3] * dx[3] + V[i * 3 + 1] * dx[4] + V[i * 3 + 2] * dx[5];
}

// 6x6 Cholesky solve; H given as its 21 upper-triangle entries in row order. Same pivot rule as the oracle. Fully
// unrolled (everything stays in registers) with one reciprocal per pivot: this runs on one thread between two barriers,
// so its dependent-operation chain is on the critical path of every Gauss-Newton iteration.
__device__ __forceinline__ bool solve6_d(const double *Hu, const double *b, double (&x)[6]) {
    double H[6][6];
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = a; c < 6; c++) {
                H[a][c] = Hu[k];
                H[c][a] = Hu[k];
                k++;
            }
    }
    double maxd = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) maxd = fmax(maxd, fabs(H[i][i]));
    if (!(maxd > 0)) return false;
    const double tiny = 1e-13 * maxd;
    double L[6][6], inv[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = H[j][j];
#pragma unroll
        for (int q = 0; q < j; q++) d -= L[j][q] * L[j][q];
        if (!(d > tiny)) ok = false;
        const double r = rsqrt(d);
        inv[j] = r;
        L[j][j] = d * r;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = H[i][j];
#pragma unroll
            for (int q = 0; q < j; q++) s -= L[i][q] * L[j][q];
            L[i][j] = s * r;
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int q = 0; q < i; q++) s -= L[i][q] * y[q];
        y[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = y[i];
#pragma unroll
        for (int q = i + 1; q < 6; q++) s -= L[q][i] * x[q];
        x[i] = s * inv[i];
    }
    return true;
}

struct SolverShared {
    double R[9], t[3];
    double part[TP_WARPS][28];  // per-warp partial sums: 21 JtJ + 6 Jtr (+1 pad)
    int    ipart[TP_WARPS];
    int    flag;                // 0 continue, 1 converged, 2 solver failure
    int    stats[4];
};

// One correspondence: residual, Huber weight, 2x6 Jacobian -> 27 sums. Explicit fma(): this translation unit is compiled
// with -fmad=false for the frustum's binary32 parity; the solver's tolerance is 1e-5 relative, not bit-exactness.
__device__ __forceinline__ void accumulate_point(const CamD &cam, const double *R, const double *t, float X0, float X1, float X2,
                                                 float ou, float ov, bool robust, double delta, double (&acc)[27]) {
    const double X[3] = {X0, X1, X2};
    double Xc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) Xc[r] = fma(R[r * 3], X[0], fma(R[r * 3 + 1], X[1], fma(R[r * 3 + 2], X[2], t[r])));
    if (!(Xc[2] > 0.0)) return;  // isDepthPositive (OptimizableTypes.h:48-52)
    const double x = Xc[0], y = Xc[1], z = Xc[2];
    double u, v, J0[6], J1[6];
    if (cam.model == MOVFE_CAM_FISHEYE) {
        double Jp[6];
        project_with_jac_kb8(cam, x, y, z, u, v, Jp);
        // J = -Jp * [ -[Xc]x | I ]  (OptimizableTypes.cpp:63-68)
        const double D[3][6] = {{0, z, -y, 1, 0, 0}, {-z, 0, x, 0, 1, 0}, {y, -x, 0, 0, 0, 1}};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            J0[k] = -(Jp[0] * D[0][k] + Jp[1] * D[1][k] + Jp[2] * D[2][k]);
            J1[k] = -(Jp[3] * D[0][k] + Jp[4] * D[1][k] + Jp[5] * D[2][k]);
        }
    } else {
        // Pinhole (Pinhole.cpp:37-43, 77-88) with the zero entries of Jp folded away
        const double iz = 1.0 / z;
        const double a = cam.fx * iz, b = cam.fy * iz;  // Jp[0], Jp[4]
        const double xz = x * iz, yz = y * iz;
        u = fma(cam.fx, xz, cam.cx);
        v = fma(cam.fy, yz, cam.cy);
        const double c = -a * xz, d = -b * yz;  // Jp[2], Jp[5]
        J0[0] = -(c * y);
        J0[1] = -(a * z - c * x);
        J0[2] = a * y;
        J0[3] = -a;
        J0[4] = 0.0;
        J0[5] = -c;
        J1[0] = -(d * y - b * z);
        J1[1] = d * x;
        J1[2] = -(b * x);
        J1[3] = 0.0;
        J1[4] = -b;
        J1[5] = -d;
    }
    const double e0 = (double)ou - u, e1 = (double)ov - v;  // OptimizableTypes.h:41-46
    const double chi2 = fma(e0, e0, e1 * e1);
    const double w = (robust && chi2 > delta * delta) ? delta * rsqrt(chi2) : 1.0;  // RobustKernelHuber rho'
    double wJ0[6], wJ1[6];
#pragma unroll
    for (int a = 0; a < 6; a++) {
        wJ0[a] = w * J0[a];
        wJ1[a] = w * J1[a];
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int c = a; c < 6; c++) {
            acc[q] = fma(wJ0[a], J0[c], fma(wJ1[a], J1[c], acc[q]));
            q++;
        }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) acc[21 + a] -= fma(wJ0[a], e0, wJ1[a] * e1);
}

__device__ __forceinline__ bool classify_point(const CamD &cam, const double *R, const double *t, float X0, float X1, float X2,
                                               float ou, float ov, double chi2thr) {
    const double X[3] = {X0, X1, X2};
    double Xc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) Xc[r] = fma(R[r * 3], X[0], fma(R[r * 3 + 1], X[1], fma(R[r * 3 + 2], X[2], t[r])));
    if (!(Xc[2] > 0.0)) return true;
    double u, v;
    project_d(cam, Xc[0], Xc[1], Xc[2], u, v);
    const double e0 = (double)ou - u, e1 = (double)ov - v;
    return (e0 * e0 + e1 * e1) > chi2thr;
}

// Sum of 27 per-lane doubles over the warp by recursive halving: at every step a lane keeps one half of its values and
// trades the other half with its partner, so 16+8+4+2+1 values cross the warp instead of 27 x 5. On return lane l holds
// the warp total of value slot_of(l) (bit-reversed index), or nothing for slots >= 27.
__device__ __forceinline__ double shfl_xor_d(double v, int o) {
    return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), o), __shfl_xor_sync(0xffffffffu, __double2loint(v), o));
}

template <int HALF>
__device__ __forceinline__ void halve(double (&v)[32], int lane) {
    const bool upper = lane & HALF;
#pragma unroll
    for (int i = 0; i < HALF; i++) {
        const double keep = upper ? v[i + HALF] : v[i];
        const double send = upper ? v[i] : v[i + HALF];
        v[i] = keep + shfl_xor_d(send, HALF);
    }
}

__device__ __forceinline__ double warp_reduce27(const double (&acc)[27], int lane) {
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = i < 27 ? acc[i] : 0.0;
    halve<16>(v, lane);
    halve<8>(v, lane);
    halve<4>(v, lane);
    halve<2>(v, lane);
    halve<1>(v, lane);
    return v[0];  // slot (lane&16) + (lane&8) + ... : the lane's own index
}

// CTA-cooperative PoseOptimization. `Src` yields correspondence i of n: X(i), obs(i); every one is valid.
// outlier[i]: 1 = outlier, 0 = inlier. pose (global) is read and, unless fewer than 4 correspondences exist, overwritten.
// Only ceil(n/32) warps (at most the CTA) take part in the per-point passes. Returns the inlier count.
template <typename Src>
__device__ int pose_solve(const Src &src, int n, const movfe_camera &cam_, const movfe_pose_params &pp, movfe_pose *pose,
                          uint8_t *outlier, int32_t *stats_out, SolverShared &sh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CamD cam = widen(cam_);
    const float repErrorF = pp.is_lost ? (float)pp.reprojection_error_lost : (float)pp.reprojection_error;  // Optimizer.cc:423-427
    const double delta = repErrorF, chi2thr = delta * delta;
    const int its = pp.iteration_count / 4 > 1 ? pp.iteration_count / 4 : 1;
    const int nw = min((n + 31) >> 5, (int)(blockDim.x >> 5));  // warps that own points
    const int nthr = nw * 32;

    for (int i = threadIdx.x; i < n; i += blockDim.x) outlier[i] = 0;
    if (threadIdx.x < 9) sh.R[threadIdx.x] = pose->R[threadIdx.x];
    if (threadIdx.x < 3) sh.t[threadIdx.x] = pose->t[threadIdx.x];
    if (threadIdx.x < 4) sh.stats[threadIdx.x] = 0;
    __syncthreads();
    if (n < 4) {  // Optimizer.cc:415-418
        if (stats_out && threadIdx.x < 4) stats_out[threadIdx.x] = 0;
        return 0;
    }
    int n_bad = 0;
    for (int round = 0; round < 4; round++) {
        const bool robust = round < 3;
        for (int it = 0; it < its; it++) {
            if (warp < nw) {
                double acc[27];
#pragma unroll
                for (int q = 0; q < 27; q++) acc[q] = 0.0;
                for (int i = threadIdx.x; i < n; i += nthr) {
                    if (outlier[i]) continue;
                    float X0, X1, X2, ou, ov;
                    src.get(i, X0, X1, X2, ou, ov);
                    accumulate_point(cam, sh.R, sh.t, X0, X1, X2, ou, ov, robust, delta, acc);
                }
                const double v = warp_reduce27(acc, lane);
                if (lane < 27) sh.part[warp][lane] = v;
            }
            __syncthreads();
            if (threadIdx.x < 27) {
                double v = 0;
                for (int w = 0; w < nw; w++) v += sh.part[w][threadIdx.x];
                sh.part[0][threadIdx.x] = v;  // only thread q touches column q
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                sh.stats[0]++;
                sh.stats[2]++;
                double dx[6];
                if (!solve6_d(sh.part[0], &sh.part[0][21], dx)) {
                    sh.flag = 2;
                    sh.stats[3]++;
                } else {
                    double dR[9], dt[3], Rn[9], tn[3];
                    se3_exp_d(dx, dR, dt);
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int j = 0; j < 3; j++)
                            Rn[i * 3 + j] = dR[i * 3] * sh.R[j] + dR[i * 3 + 1] * sh.R[3 + j] + dR[i * 3 + 2] * sh.R[6 + j];
#pragma unroll
                    for (int r = 0; r < 3; r++) tn[r] = dR[r * 3] * sh.t[0] + dR[r * 3 + 1] * sh.t[1] + dR[r * 3 + 2] * sh.t[2] + dt[r];
#pragma unroll
                    for (int i = 0; i < 9; i++) sh.R[i] = Rn[i];
#pragma unroll
                    for (int i = 0; i < 3; i++) sh.t[i] = tn[i];
                    double m = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) m = fmax(m, fabs(dx[a]));
                    sh.flag = m < 1e-10 ? 1 : 0;
                }
            }
            __syncthreads();
            const int flag = sh.flag;
            if (flag) break;
        }
        // re-classification of every correspondence
        int bad = 0;
        if (warp < nw) {
            for (int i = threadIdx.x; i < n; i += nthr) {
                float X0, X1, X2, ou, ov;
                src.get(i, X0, X1, X2, ou, ov);
                const bool b = classify_point(cam, sh.R, sh.t, X0, X1, X2, ou, ov, chi2thr);
                outlier[i] = b ? 1 : 0;
                bad += b;
            }
            for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        }
        __syncthreads();  // everyone is past the flag read / previous ipart use
        if (lane == 0 && warp < nw) sh.ipart[warp] = bad;
        if (threadIdx.x == 0) {
            sh.stats[1]++;
            sh.stats[2]++;
        }
        __syncthreads();
        n_bad = 0;
        for (int w = 0; w < nw; w++) n_bad += sh.ipart[w];
        if (n - n_bad < 3) break;
    }
    __syncthreads();
    if (threadIdx.x < 9) pose->R[threadIdx.x] = sh.R[threadIdx.x];
    if (threadIdx.x < 3) pose->t[threadIdx.x] = sh.t[threadIdx.x];
    if (stats_out && threadIdx.x < 4) stats_out[threadIdx.x] = sh.stats[threadIdx.x];
    __syncthreads();
    return n - n_bad;
}

// correspondence sources
struct DirectSrc {  // packed global arrays (movfe_pose_optimize)
    const float *pts, *obs;
    __device__ void get(int i, float &X0, float &X1, float &X2, float &u, float &v) const {
        X0 = pts[3 * i];
        X1 = pts[3 * i + 1];
        X2 = pts[3 * i + 2];
        u = obs[2 * i];
        v = obs[2 * i + 1];
    }
};

// Optimizer.cc:404-413 gathers, for every keypoint with a map point, (GetWorldPos(), mvKeys[i].pt). The persistent
// driver compacts them once per solve into shared memory (keypoint order), so the passes never touch global memory.
struct CompactSrc {
    const float *x, *y, *z, *u, *v;
    __device__ void get(int i, float &X0, float &X1, float &X2, float &ou, float &ov) const {
        X0 = x[i];
        X1 = y[i];
        X2 = z[i];
        ou = u[i];
        ov = v[i];
    }
};

// ----------------------------------------------------------------------------------- persistent driver ------
struct TrackPoseParams {
    int S, W, H, maxT, maxMap, TSLOTS, F, hash_cap;
    int n_frames;
    int tslot0;       // track-table slot of the first frame; slots advance modulo TSLOTS
    int out0;         // output slot (frame - pose window start) of the first frame
    float view_cos;
    movfe_camera cam;
    movfe_pose_params pp;
};

// Exclusive scan of one int per thread over the CTA (TP_THREADS); wsum: TP_WARPS ints of shared memory.
__device__ __forceinline__ int tp_excl_scan(int v, int *wsum, int &total) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < TP_WARPS; w++) {
        const int c = wsum[w];
        before += w < warp ? c : 0;
        tot += c;
    }
    total = tot;
    __syncthreads();
    return before + x - v;
}

// Optimizer.cc:404-413: the frame's (map point, keypoint) pairs in keypoint order -> shared memory. Returns the count.
__device__ int gather_pairs(const movfe_track *__restrict__ tr, int n, const int32_t *__restrict__ match,
                            const movfe_map_point *__restrict__ mp, float *cx, float *cy, float *cz, float *cu, float *cv, int *cidx,
                            int cap, int *wsum) {
    int n_pairs = 0;
    for (int base = 0; base < n; base += TP_THREADS) {
        const int t = base + threadIdx.x;
        const int m = t < n ? match[t] : -1;
        int tot;
        const int pos = n_pairs + tp_excl_scan(m >= 0 ? 1 : 0, wsum, tot);
        if (m >= 0 && pos < cap) {
            cx[pos] = mp[m].pos[0];
            cy[pos] = mp[m].pos[1];
            cz[pos] = mp[m].pos[2];
            const float2 pt = *reinterpret_cast<const float2 *>(&tr[t].pt_x);
            cu[pos] = pt.x;
            cv[pos] = pt.y;
            cidx[pos] = t;
        }
        n_pairs += tot;
    }
    __syncthreads();
    return min(n_pairs, cap);
}

// One join of the frame (MOVMatcher.h:35-103) with the hash on the PROBE side, so shared memory scales with the map
// points (hundreds) and not with the track table (thousands): key = track id, value = the last eligible map point with
// that id (list order, last wins). Every track then looks its id up; of several tracks with one id only the first owns
// the match (F.mvVFMap is first-wins, MOVExtractor.cc:330) - resolved by an integer atomicMin per map point.
// reset: SearchByVideoFeature(KF, F, out) starts from an all-NULL vector; the Frame overload keeps earlier matches.
template <typename Elig>
__device__ void join_frame(const movfe_track *__restrict__ tr, int n, const movfe_map_point *__restrict__ mp, int n_pts, Elig elig,
                           bool reset, int *keys, int *vals, int cap, int *first, int32_t *__restrict__ match) {
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        keys[i] = HASH_EMPTY;
        vals[i] = -1;
    }
    for (int i = threadIdx.x; i < n_pts; i += blockDim.x) first[i] = 0x7fffffff;
    __syncthreads();
    for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
        if (!elig(i)) continue;
        const int id = mp[i].track_id;
        if (id == HASH_EMPTY) continue;
        unsigned h = hash_id(id) & (cap - 1);
        while (true) {
            const int prev = atomicCAS(&keys[h], HASH_EMPTY, id);
            if (prev == HASH_EMPTY || prev == id) {
                atomicMax(&vals[h], i);
                break;
            }
            h = (h + 1) & (cap - 1);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int m = hash_find(tr[t].track_id, keys, vals, cap);
        if (m >= 0) atomicMin(&first[m], t);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int m = hash_find(tr[t].track_id, keys, vals, cap);
        if (m >= 0 && first[m] == t) match[t] = m;
        else if (reset) match[t] = -1;
    }
    __syncthreads();
}

#ifndef MOVFE_TP_MINB
#define MOVFE_TP_MINB (512 / MOVFE_TP_THREADS)  // 128 registers per thread
#endif
__global__ void __launch_bounds__(TP_THREADS, MOVFE_TP_MINB)
track_poses_kernel(TrackPoseParams p, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
                   const movfe_map_point *__restrict__ map, const int32_t *__restrict__ nmap, const int32_t *__restrict__ nkf,
                   movfe_pose *__restrict__ pose_cur, movfe_pose *__restrict__ poses, int32_t *__restrict__ ninl,
                   int32_t *__restrict__ match_out, uint8_t *__restrict__ outlier_out, int32_t *__restrict__ skip_tag,
                   unsigned long long *__restrict__ stats) {
    extern __shared__ int hsm[];  // keys[cap] vals[cap] first[maxMap] | pairs: x y z u v idx [maxMap] out[maxMap]
    __shared__ SolverShared sh;
    __shared__ int wsum[TP_WARPS];
    int *keys = hsm, *vals = hsm + p.hash_cap, *first = hsm + 2 * p.hash_cap;
    float *cx = reinterpret_cast<float *>(first + p.maxMap), *cy = cx + p.maxMap, *cz = cy + p.maxMap, *cu = cz + p.maxMap, *cv = cu + p.maxMap;
    int *cidx = reinterpret_cast<int *>(cv + p.maxMap);
    uint8_t *cout = reinterpret_cast<uint8_t *>(cidx + p.maxMap);
    const int s = blockIdx.x;
    const movfe_map_point *mp = map + (size_t)s * p.maxMap;
    const int n_map = nmap[s], n_kf = min(nkf[s], n_map);
    int32_t *tag = skip_tag + (size_t)s * p.maxMap;
    movfe_pose *pc = pose_cur + s;
    const CompactSrc src{cx, cy, cz, cu, cv};
    int cap = 2;
    while (cap < 2 * n_map) cap <<= 1;

    for (int k = 0; k < p.n_frames; k++) {
        const int ts = (p.tslot0 + k) % p.TSLOTS;
        const movfe_track *tr = tracks + ((size_t)s * p.TSLOTS + ts) * p.maxT;
        const int n = ntracks[s * p.TSLOTS + ts];
        int32_t *match = match_out + ((size_t)s * p.F + p.out0 + k) * p.maxT;
        uint8_t *outl = outlier_out + ((size_t)s * p.F + p.out0 + k) * p.maxT;
        int n_inl = 0;
        if (n > 0 && n_map > 0) {
            // --- TrackReferenceKeyFrame: SearchByVideoFeature(KF, F, matches) (MOVMatcher.h:70-103)
            join_frame(tr, n, mp, n_kf, [&](int i) { return !(mp[i].flags & (MOVFE_MP_NULL | MOVFE_MP_BAD)); }, true, keys, vals, cap,
                       first, match);
            int np = gather_pairs(tr, n, match, mp, cx, cy, cz, cu, cv, cidx, p.maxMap, wsum);
            pose_solve(src, np, p.cam, p.pp, pc, cout, nullptr, sh);  // pose := last frame's pose is already in pc (Tracking.cc:807)
            if (threadIdx.x == 0) {  // workload counters (diagnostic)
                atomicAdd(&stats[2], 1ull);
                atomicAdd(&stats[3], (unsigned long long)np);
                atomicAdd(&stats[4], (unsigned long long)sh.stats[2]);
            }
            // --- TrackLocalMap / SearchLocalPoints (Tracking.cc:1109-1158)
            const int frame_tag = 1;  // tags are cleared again below, so one value is enough
            for (int i = threadIdx.x; i < np; i += blockDim.x) tag[match[cidx[i]]] = frame_tag;  // mnLastFrameSeen = current frame
            __syncthreads();
            const FrustumPose fp = frustum_pose(*pc);
            // isInFrustum for every local point (:1136-1151); in view and not bad -> eligible (MOVMatcher.h:43-49, far filter off)
            for (int i = threadIdx.x; i < n_map; i += blockDim.x) {
                const movfe_map_point m = mp[i];
                const movfe_projection pr = frustum_point(fp, p.cam, p.W, p.H, p.view_cos, m, tag[i] == frame_tag);
                cout[i] = pr.in_view && !(m.flags & MOVFE_MP_BAD);
            }
            __syncthreads();
            join_frame(tr, n, mp, n_map, [&](int i) { return cout[i] != 0; }, false, keys, vals, cap, first, match);
            np = gather_pairs(tr, n, match, mp, cx, cy, cz, cu, cv, cidx, p.maxMap, wsum);
            n_inl = pose_solve(src, np, p.cam, p.pp, pc, cout, nullptr, sh);
            if (threadIdx.x == 0) {
                atomicAdd(&stats[2], 1ull);
                atomicAdd(&stats[3], (unsigned long long)np);
                atomicAdd(&stats[4], (unsigned long long)sh.stats[2]);
            }
            // Frame::mvbOutlier (Optimizer.cc:452-456): true everywhere, false for the inliers
            for (int t = threadIdx.x; t < n; t += blockDim.x) outl[t] = 1;
            __syncthreads();
            if (np >= 4)
                for (int i = threadIdx.x; i < np; i += blockDim.x) outl[cidx[i]] = cout[i];
            else
                for (int i = threadIdx.x; i < np; i += blockDim.x) outl[cidx[i]] = 0;  // <4 pairs: the frame is left untouched
            for (int i = threadIdx.x; i < n_map; i += blockDim.x) tag[i] = 0;  // leave the tags clean for the next launch
            __syncthreads();
        } else {
            for (int t = threadIdx.x; t < n; t += blockDim.x) {
                match[t] = -1;
                outl[t] = 1;
            }
        }
        if (threadIdx.x == 0) {
            poses[(size_t)s * p.F + p.out0 + k] = *pc;
            ninl[(size_t)s * p.F + p.out0 + k] = n_inl;
        }
        __syncthreads();
    }
}

// ================================================================================== pose chain, second form ======
// The first form above (kept behind MOVFE_POSE_V1=1 for A/B runs) walks the whole track table of a frame - thousands of
// 64-byte records - three times per join and seventeen block scans per gather, and every Gauss-Newton pass parks the CTA at
// three barriers around a one-thread 6x6 solve. It runs beside propagation but was the longer of the two chains.
// Second form:
//  * tp_prep_kernel (wide: one CTA per frame and stream, all frames of a window at once, nothing pose-dependent): for every
//    local map point the FIRST track carrying its id (F.mvVFMap, MOVExtractor.cc:330) and the first map point of its id group.
//    What remains of a join is then work on the few hundred map points only: the last eligible point of an id group owns
//    the group's track (MOVMatcher.h:35-103 walk the points in order and overwrite).
//  * the pairs are put in keypoint order (Optimizer.cc:404-413) through a bitmap of the matched tracks: one popcount prefix.
//  * pose_solve2: every pass ends at ONE barrier; each warp then adds the per-warp partial sums in the fixed order and solves
//    the 6x6 system itself (identical arithmetic in every thread, so the poses stay bit-identical across the CTA), and the
//    re-classification that closes a round is fused into the first pass of the next round (same pose, same residuals).
//    (One solver warp between two barriers instead - a third fewer instructions for the chain - measured 1.5 % slower for the
//    step: the chain's latency, not its instruction count, is what the step sees.)
// Sums are formed in the same order as in the first form, so both give the same bits.
#ifndef MOVFE_SOLVE_LANE0
#define MOVFE_SOLVE_LANE0 1
#endif
struct Solver2Shared {
    double part[2][TP_WARPS][32];  // per-warp partial sums (21 JtJ + 6 Jtr + outlier count), double-buffered by pass parity
    double tot[TP_WARPS][28];      // each warp's own copy of the totals
    double Rt[TP_WARPS][12];       // each warp's own copy of the pose
    int ipart[2][TP_WARPS];
};

template <typename Src>
__device__ int pose_solve2(const Src &src, int n, const movfe_camera &cam_, const movfe_pose_params &pp, movfe_pose *pose,
                           uint8_t *outlier, int (&stats)[4], int &executed, Solver2Shared &sh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CamD cam = widen(cam_);
    const float repErrorF = pp.is_lost ? (float)pp.reprojection_error_lost : (float)pp.reprojection_error;  // Optimizer.cc:423-427
    const double delta = repErrorF, chi2thr = delta * delta;
    const int its = pp.iteration_count / 4 > 1 ? pp.iteration_count / 4 : 1;
    const int nw = min((n + 31) >> 5, (int)(blockDim.x >> 5));  // warps that own points
    const int nthr = nw * 32;
    double *Rt = sh.Rt[warp];
    stats[0] = stats[1] = stats[2] = stats[3] = 0;  // iterations, classifications, their sum, solver failures (the oracle's counters)
    executed = 0;                                   // passes over the correspondences actually made (a fused pass counts once)
    for (int i = threadIdx.x; i < n; i += blockDim.x) outlier[i] = 0;
    if (lane < 9) Rt[lane] = pose->R[lane];
    else if (lane < 12) Rt[lane] = pose->t[lane - 9];
    __syncthreads();
    if (n < 4) return 0;  // Optimizer.cc:415-418
    int n_bad = 0, pass = 0;
    bool pending = false;  // a round has ended: every correspondence is to be re-classified at the current pose
    bool stop = false;
    for (int round = 0; round < 4 && !stop; round++) {
        const bool robust = round < 3;
        for (int it = 0; it < its; it++) {
            double acc[27];
#pragma unroll
            for (int q = 0; q < 27; q++) acc[q] = 0.0;
            int bad = 0;
            if (warp < nw) {
                for (int i = threadIdx.x; i < n; i += nthr) {
                    float X0, X1, X2, ou, ov;
                    src.get(i, X0, X1, X2, ou, ov);
                    if (pending) {
                        const bool b = classify_point(cam, Rt, Rt + 9, X0, X1, X2, ou, ov, chi2thr);
                        outlier[i] = b ? 1 : 0;  // only this thread reads it back in later passes
                        bad += b;
                        if (b) continue;
                    } else if (outlier[i]) {
                        continue;
                    }
                    accumulate_point(cam, Rt, Rt + 9, X0, X1, X2, ou, ov, robust, delta, acc);
                }
                const double v = warp_reduce27(acc, lane);
                sh.part[pass & 1][warp][lane] = v;
                if (pending) {
                    for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
                    if (lane == 0) sh.ipart[pass & 1][warp] = bad;
                }
            }
            __syncthreads();  // the only barrier of a pass (the partials of the pass after next go to this buffer again: by then
                              // every warp has passed the next barrier, i.e. has finished reading these)
            if (lane < 27) {
                double v = 0;
                for (int w = 0; w < nw; w++) v += sh.part[pass & 1][w][lane];
                sh.tot[warp][lane] = v;
            }
            executed++;
            if (pending) {
                n_bad = 0;
                for (int w = 0; w < nw; w++) n_bad += sh.ipart[pass & 1][w];
                stats[1]++;
                pending = false;
                if (n - n_bad < 3) {  // the first form leaves the rounds here, before another iteration
                    stop = true;
                    pass++;
                    break;
                }
            }
            pass++;
            __syncwarp();
            stats[0]++;
            int flag = 0;
#if MOVFE_SOLVE_LANE0
            if (lane == 0)  // one lane per warp solves (the warp's own copy of the pose): the FP64 pipe sees one lane, not 32
#endif
            {
                double dx[6];
                if (!solve6_d(sh.tot[warp], &sh.tot[warp][21], dx)) {
                    flag = 2;
                } else {
                    double dR[9], dt[3], Rn[9], tn[3];
                    se3_exp_d(dx, dR, dt);
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int j = 0; j < 3; j++) Rn[i * 3 + j] = dR[i * 3] * Rt[j] + dR[i * 3 + 1] * Rt[3 + j] + dR[i * 3 + 2] * Rt[6 + j];
#pragma unroll
                    for (int r = 0; r < 3; r++) tn[r] = dR[r * 3] * Rt[9] + dR[r * 3 + 1] * Rt[10] + dR[r * 3 + 2] * Rt[11] + dt[r];
#if !MOVFE_SOLVE_LANE0
                    __syncwarp();  // every lane has read the old pose
#endif
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 9; i++) Rt[i] = Rn[i];
#pragma unroll
                        for (int i = 0; i < 3; i++) Rt[9 + i] = tn[i];
                    }
                    double m = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) m = fmax(m, fabs(dx[a]));
                    flag = m < 1e-10 ? 1 : 0;
                }
            }
            __syncwarp();
#if MOVFE_SOLVE_LANE0
            flag = __shfl_sync(0xffffffffu, flag, 0);
#endif
            if (flag == 2) stats[3]++;
            if (flag) break;
        }
        if (!stop) pending = true;
    }
    if (pending) {  // the classification that closes the last round
        int bad = 0;
        if (warp < nw) {
            for (int i = threadIdx.x; i < n; i += nthr) {
                float X0, X1, X2, ou, ov;
                src.get(i, X0, X1, X2, ou, ov);
                const bool b = classify_point(cam, Rt, Rt + 9, X0, X1, X2, ou, ov, chi2thr);
                outlier[i] = b ? 1 : 0;
                bad += b;
            }
            for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
            if (lane == 0) sh.ipart[pass & 1][warp] = bad;
        }
        __syncthreads();
        n_bad = 0;
        for (int w = 0; w < nw; w++) n_bad += sh.ipart[pass & 1][w];
        stats[1]++;
        executed++;
    }
    stats[2] = stats[0] + stats[1];
    if (threadIdx.x < 9) pose->R[threadIdx.x] = Rt[threadIdx.x];
    else if (threadIdx.x < 12) pose->t[threadIdx.x - 9] = Rt[threadIdx.x];
    __syncthreads();  // outlier[] and the pose are visible to the whole CTA
    return n - n_bad;
}

// first track / group leader of every map point of one (frame, stream): (first track index + 1) | leader point << 16
__global__ void __launch_bounds__(TP_THREADS)
tp_prep_kernel(TrackPoseParams p, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
               const movfe_map_point *__restrict__ map, const int32_t *__restrict__ nmap, uint32_t *__restrict__ first_track,
               int32_t *__restrict__ match_out, uint8_t *__restrict__ outlier_out) {
    extern __shared__ int hsm[];  // keys[cap] ftrack[cap] leader[cap]
    int *keys = hsm, *ftrack = hsm + p.hash_cap, *leader = hsm + 2 * p.hash_cap;
    const int k = blockIdx.x, s = blockIdx.y;
    const movfe_map_point *mp = map + (size_t)s * p.maxMap;
    const int n_map = nmap[s];
    const int ts = (p.tslot0 + k) % p.TSLOTS;
    const movfe_track *tr = tracks + ((size_t)s * p.TSLOTS + ts) * p.maxT;
    const int n = ntracks[s * p.TSLOTS + ts];
    int32_t *match = match_out + ((size_t)s * p.F + p.out0 + k) * p.maxT;
    uint8_t *outl = outlier_out + ((size_t)s * p.F + p.out0 + k) * p.maxT;
    uint32_t *ft = first_track + ((size_t)s * p.F + p.out0 + k) * p.maxMap;
    // SearchByVideoFeature(KF, F, matches) starts from an all-NULL vector (MOVMatcher.h:76), mvbOutlier from all true (Optimizer.cc:452)
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        match[t] = -1;
        outl[t] = 1;
    }
    int cap = 2;
    while (cap < 2 * n_map) cap <<= 1;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        keys[i] = HASH_EMPTY;
        ftrack[i] = 0x7fffffff;
        leader[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_map; i += blockDim.x) {
        const int id = mp[i].track_id;
        if (id == HASH_EMPTY) continue;
        unsigned h = hash_id(id) & (cap - 1);
        while (true) {
            const int prev = atomicCAS(&keys[h], HASH_EMPTY, id);
            if (prev == HASH_EMPTY || prev == id) {
                atomicMin(&leader[h], i);
                break;
            }
            h = (h + 1) & (cap - 1);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int id = tr[t].track_id;
        if (id == HASH_EMPTY) continue;
        unsigned h = hash_id(id) & (cap - 1);
        while (true) {
            const int kk = keys[h];
            if (kk == id) {
                atomicMin(&ftrack[h], t);  // integer min: the first index wins whatever the thread order
                break;
            }
            if (kk == HASH_EMPTY) break;
            h = (h + 1) & (cap - 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_map; i += blockDim.x) {
        const int id = mp[i].track_id;
        uint32_t v = 0;
        if (id != HASH_EMPTY) {
            unsigned h = hash_id(id) & (cap - 1);
            while (keys[h] != id) h = (h + 1) & (cap - 1);
            const int f = ftrack[h];
            v = (f == 0x7fffffff ? 0u : (uint32_t)(f + 1)) | ((uint32_t)leader[h] << 16);
        }
        ft[i] = v;
    }
}

// keypoint-ordered (map point, keypoint) pairs of a join result: win[g] = the map point that owns the track of id group g
// (set at the group's leader only), -1 elsewhere. Returns the pair count. bits/pref: maxT/32 words each.
__device__ int gather_pairs2(const movfe_track *__restrict__ tr, const movfe_map_point *__restrict__ mp, int n_map, const uint32_t *pk,
                             const int *win, uint32_t *bits, int *pref, int n_words, float *cx, float *cy, float *cz, float *cu, float *cv,
                             int *cidx, int *cm, int cap, int *wsum, int32_t *__restrict__ match) {
    for (int w = threadIdx.x; w < n_words; w += blockDim.x) bits[w] = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < n_map; g += blockDim.x)
        if (win[g] >= 0) {
            const int t = (int)(pk[g] & 0xffffu) - 1;
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
    __syncthreads();
    int n_pairs = 0;
    for (int base = 0; base < n_words; base += TP_THREADS) {
        const int w = base + threadIdx.x;
        const int c = w < n_words ? __popc(bits[w]) : 0;
        int tot;
        const int e = n_pairs + tp_excl_scan(c, wsum, tot);
        if (w < n_words) pref[w] = e;
        n_pairs += tot;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < n_map; g += blockDim.x) {
        const int m = win[g];
        if (m < 0) continue;
        const int t = (int)(pk[g] & 0xffffu) - 1;
        const int pos = pref[t >> 5] + __popc(bits[t >> 5] & ((1u << (t & 31)) - 1u));
        match[t] = m;
        if (pos < cap) {
            cx[pos] = mp[m].pos[0];
            cy[pos] = mp[m].pos[1];
            cz[pos] = mp[m].pos[2];
            const float2 pt = *reinterpret_cast<const float2 *>(&tr[t].pt_x);
            cu[pos] = pt.x;
            cv[pos] = pt.y;
            cidx[pos] = t;
            cm[pos] = m;
        }
    }
    __syncthreads();
    return min(n_pairs, cap);
}

__global__ void __launch_bounds__(TP_THREADS, MOVFE_TP_MINB)
track_poses2_kernel(TrackPoseParams p, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
                    const movfe_map_point *__restrict__ map, const int32_t *__restrict__ nmap, const int32_t *__restrict__ nkf,
                    const uint32_t *__restrict__ first_track, movfe_pose *__restrict__ pose_cur, movfe_pose *__restrict__ poses,
                    int32_t *__restrict__ ninl, int32_t *__restrict__ match_out, uint8_t *__restrict__ outlier_out,
                    unsigned long long *__restrict__ stats) {
    extern __shared__ int hsm[];  // pk win1 win2 cidx cm [maxMap] | x y z u v [maxMap] | bits pref [maxT/32] | cout tag [maxMap bytes]
    __shared__ Solver2Shared sh;
    __shared__ int wsum[TP_WARPS];
    const int M = p.maxMap, NWORDS = (p.maxT + 31) / 32;
    uint32_t *pk = reinterpret_cast<uint32_t *>(hsm);
    int *win1 = hsm + M, *win2 = hsm + 2 * M, *cidx = hsm + 3 * M, *cm = hsm + 4 * M;
    float *cx = reinterpret_cast<float *>(hsm + 5 * M), *cy = cx + M, *cz = cy + M, *cu = cz + M, *cv = cu + M;
    uint32_t *bits = reinterpret_cast<uint32_t *>(cv + M);
    int *pref = reinterpret_cast<int *>(bits + NWORDS);
    uint8_t *cout = reinterpret_cast<uint8_t *>(pref + NWORDS), *tag = cout + M;
    const int s = blockIdx.x;
    const movfe_map_point *mp = map + (size_t)s * p.maxMap;
    const int n_map = nmap[s], n_kf = min(nkf[s], n_map);
    movfe_pose *pc = pose_cur + s;
    const CompactSrc src{cx, cy, cz, cu, cv};

    for (int k = 0; k < p.n_frames; k++) {
        const int ts = (p.tslot0 + k) % p.TSLOTS;
        const movfe_track *tr = tracks + ((size_t)s * p.TSLOTS + ts) * p.maxT;
        const int n = ntracks[s * p.TSLOTS + ts];
        int32_t *match = match_out + ((size_t)s * p.F + p.out0 + k) * p.maxT;
        uint8_t *outl = outlier_out + ((size_t)s * p.F + p.out0 + k) * p.maxT;
        const uint32_t *ft = first_track + ((size_t)s * p.F + p.out0 + k) * p.maxMap;
        int n_inl = 0;
        if (n > 0 && n_map > 0) {
            const int n_words = (n + 31) >> 5;
            int st[4], passes;
            // --- TrackReferenceKeyFrame: SearchByVideoFeature(KF, F, matches) (MOVMatcher.h:70-103)
            for (int i = threadIdx.x; i < n_map; i += blockDim.x) {
                pk[i] = ft[i];
                win1[i] = -1;
                win2[i] = -1;
                tag[i] = 0;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n_kf; i += blockDim.x)
                if (!(mp[i].flags & (MOVFE_MP_NULL | MOVFE_MP_BAD)) && (pk[i] & 0xffffu)) atomicMax(&win1[pk[i] >> 16], i);  // the last point of a group wins
            __syncthreads();
            int np = gather_pairs2(tr, mp, n_map, pk, win1, bits, pref, n_words, cx, cy, cz, cu, cv, cidx, cm, M, wsum, match);
            pose_solve2(src, np, p.cam, p.pp, pc, cout, st, passes, sh);  // pose := last frame's pose is already in pc (Tracking.cc:807)
            if (threadIdx.x == 0) {  // workload counters (diagnostic)
                atomicAdd(&stats[2], 1ull);
                atomicAdd(&stats[3], (unsigned long long)np);
                atomicAdd(&stats[4], (unsigned long long)passes);
            }
            // --- TrackLocalMap / SearchLocalPoints (Tracking.cc:1109-1158)
            for (int i = threadIdx.x; i < np; i += blockDim.x) tag[cm[i]] = 1;  // mnLastFrameSeen = current frame
            __syncthreads();
            const FrustumPose fp = frustum_pose(*pc);
            // isInFrustum for every local point (:1136-1151); in view and not bad -> eligible (MOVMatcher.h:43-49, far filter off).
            // The Frame overload keeps the earlier matches: a group without an eligible point keeps its owner.
            for (int i = threadIdx.x; i < n_map; i += blockDim.x) {
                const movfe_map_point m = mp[i];
                const movfe_projection pr = frustum_point(fp, p.cam, p.W, p.H, p.view_cos, m, tag[i] != 0);
                if (pr.in_view && !(m.flags & MOVFE_MP_BAD) && (pk[i] & 0xffffu)) atomicMax(&win2[pk[i] >> 16], i);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n_map; i += blockDim.x)
                if (win2[i] >= 0) win1[i] = win2[i];
            __syncthreads();
            np = gather_pairs2(tr, mp, n_map, pk, win1, bits, pref, n_words, cx, cy, cz, cu, cv, cidx, cm, M, wsum, match);
            n_inl = pose_solve2(src, np, p.cam, p.pp, pc, cout, st, passes, sh);
            if (threadIdx.x == 0) {
                atomicAdd(&stats[2], 1ull);
                atomicAdd(&stats[3], (unsigned long long)np);
                atomicAdd(&stats[4], (unsigned long long)passes);
            }
            // Frame::mvbOutlier (Optimizer.cc:452-456): true everywhere (tp_prep_kernel), false for the inliers;
            // < 4 pairs: the frame is left untouched
            for (int i = threadIdx.x; i < np; i += blockDim.x) outl[cidx[i]] = np >= 4 ? cout[i] : 0;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            poses[(size_t)s * p.F + p.out0 + k] = *pc;
            ninl[(size_t)s * p.F + p.out0 + k] = n_inl;
        }
        __syncthreads();
    }
}

// ---- the same per-frame chain as four launches (movfe_track_poses_launch, split mode) ------------------------------
// The fused kernel above keeps 256 threads x 128 registers resident per stream for the whole frame, although the two
// solves - most of its time - are a one-warp job between barriers; those registers are what the propagation CTAs on the
// same SM are waiting for. Split: the joins run wide and short, the solves run in small CTAs (a quarter or less of the
// registers), and the correspondences travel between them through a per-stream buffer in global memory.
struct PairBuf {
    float *x, *y, *z, *u, *v;
    int *idx;
};
__device__ __forceinline__ PairBuf pair_buf(float *base, int s, int maxMap) {
    float *b = base + (size_t)s * 6 * maxMap;
    return PairBuf{b, b + maxMap, b + 2 * maxMap, b + 3 * maxMap, b + 4 * maxMap, reinterpret_cast<int *>(b + 5 * maxMap)};
}

// phase 0: TrackReferenceKeyFrame's SearchByVideoFeature(KF, F, matches) + the gather of Optimizer.cc:404-413
// phase 1: SearchLocalPoints (tags, isInFrustum, SearchByVideoFeature(F, local points)) + gather; mvbOutlier prefilled
__global__ void __launch_bounds__(TP_THREADS)
tp_join_kernel(TrackPoseParams p, int phase, const movfe_track *__restrict__ tracks, const int32_t *__restrict__ ntracks,
               const movfe_map_point *__restrict__ map, const int32_t *__restrict__ nmap, const int32_t *__restrict__ nkf,
               const movfe_pose *__restrict__ pose_cur, int32_t *__restrict__ match_out, uint8_t *__restrict__ outlier_out,
               int32_t *__restrict__ skip_tag, float *__restrict__ pairs, int32_t *__restrict__ npairs) {
    extern __shared__ int hsm[];  // keys[cap] vals[cap] first[maxMap] elig[maxMap bytes]
    __shared__ int wsum[TP_WARPS];
    int *keys = hsm, *vals = hsm + p.hash_cap, *first = hsm + 2 * p.hash_cap;
    uint8_t *elig = reinterpret_cast<uint8_t *>(first + p.maxMap);
    const int s = blockIdx.x;
    const movfe_map_point *mp = map + (size_t)s * p.maxMap;
    const int n_map = nmap[s], n_kf = min(nkf[s], n_map);
    int32_t *tag = skip_tag + (size_t)s * p.maxMap;
    const PairBuf pb = pair_buf(pairs, s, p.maxMap);
    int cap = 2;
    while (cap < 2 * n_map) cap <<= 1;
    const int ts = p.tslot0 % p.TSLOTS;
    const movfe_track *tr = tracks + ((size_t)s * p.TSLOTS + ts) * p.maxT;
    const int n = ntracks[s * p.TSLOTS + ts];
    int32_t *match = match_out + ((size_t)s * p.F + p.out0) * p.maxT;
    uint8_t *outl = outlier_out + ((size_t)s * p.F + p.out0) * p.maxT;
    if (!(n > 0 && n_map > 0)) {
        if (phase == 0) {
            for (int t = threadIdx.x; t < n; t += blockDim.x) {
                match[t] = -1;
                outl[t] = 1;
            }
            if (threadIdx.x == 0) npairs[s] = 0;
        }
        return;
    }
    if (phase == 0) {
        join_frame(tr, n, mp, n_kf, [&](int i) { return !(mp[i].flags & (MOVFE_MP_NULL | MOVFE_MP_BAD)); }, true, keys, vals, cap, first, match);
    } else {
        const int np0 = npairs[s];
        for (int i = threadIdx.x; i < np0; i += blockDim.x) tag[match[pb.idx[i]]] = 1;  // mnLastFrameSeen = current frame
        for (int t = threadIdx.x; t < n; t += blockDim.x) outl[t] = 1;                  // Optimizer.cc:452: true everywhere ...
        __syncthreads();
        const FrustumPose fp = frustum_pose(pose_cur[s]);
        for (int i = threadIdx.x; i < n_map; i += blockDim.x) {
            const movfe_map_point m = mp[i];
            const movfe_projection pr = frustum_point(fp, p.cam, p.W, p.H, p.view_cos, m, tag[i] == 1);
            elig[i] = pr.in_view && !(m.flags & MOVFE_MP_BAD);
        }
        __syncthreads();
        join_frame(tr, n, mp, n_map, [&](int i) { return elig[i] != 0; }, false, keys, vals, cap, first, match);
    }
    const int np = gather_pairs(tr, n, match, mp, pb.x, pb.y, pb.z, pb.u, pb.v, pb.idx, p.maxMap, wsum);
    if (threadIdx.x == 0) npairs[s] = np;
}

// Optimizer::PoseOptimization on the gathered pairs of one stream; phase 1 also writes the frame's results.
// At most 128 threads x 128 registers = a quarter of an SM's register file: a CTA fits wherever one propagation CTA left.
constexpr int TP_SOLVE_THREADS = 128;
__global__ void __launch_bounds__(TP_SOLVE_THREADS, 4)
tp_solve_kernel(TrackPoseParams p, int phase, int np_lo, int np_hi, const int32_t *__restrict__ ntracks, const int32_t *__restrict__ nmap,
                movfe_pose *__restrict__ pose_cur, movfe_pose *__restrict__ poses, int32_t *__restrict__ ninl,
                uint8_t *__restrict__ outlier_out, int32_t *__restrict__ skip_tag, float *__restrict__ pairs,
                const int32_t *__restrict__ npairs) {
    extern __shared__ int hsm[];  // x y z u v [maxMap] out[maxMap bytes]
    __shared__ SolverShared sh;
    float *cx = reinterpret_cast<float *>(hsm), *cy = cx + p.maxMap, *cz = cy + p.maxMap, *cu = cz + p.maxMap, *cv = cu + p.maxMap;
    uint8_t *cout = reinterpret_cast<uint8_t *>(cv + p.maxMap);
    const int s = blockIdx.x;
    const int n_map = nmap[s];
    const int ts = p.tslot0 % p.TSLOTS;
    const int n = ntracks[s * p.TSLOTS + ts];
    const bool active = n > 0 && n_map > 0;
    const int np = active ? npairs[s] : 0;
    // the launch is repeated with CTA sizes fitted to the correspondence count; a CTA of the wrong size leaves at once
    if (np < np_lo || np > np_hi) return;
    movfe_pose *pc = pose_cur + s;
    int n_inl = 0;
    if (active) {
        const PairBuf pb = pair_buf(pairs, s, p.maxMap);
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            cx[i] = pb.x[i];
            cy[i] = pb.y[i];
            cz[i] = pb.z[i];
            cu[i] = pb.u[i];
            cv[i] = pb.v[i];
        }
        __syncthreads();
        const CompactSrc src{cx, cy, cz, cu, cv};
        n_inl = pose_solve(src, np, p.cam, p.pp, pc, cout, nullptr, sh);
        if (phase == 1) {
            uint8_t *outl = outlier_out + ((size_t)s * p.F + p.out0) * p.maxT;
            // Frame::mvbOutlier (Optimizer.cc:452-456): ... false for the inliers; <4 pairs: the frame is left untouched
            for (int i = threadIdx.x; i < np; i += blockDim.x) outl[pb.idx[i]] = np >= 4 ? cout[i] : 0;
            int32_t *tag = skip_tag + (size_t)s * p.maxMap;
            for (int i = threadIdx.x; i < n_map; i += blockDim.x) tag[i] = 0;  // leave the tags clean for the next frame
        }
    }
    if (phase == 1 && threadIdx.x == 0) {
        poses[(size_t)s * p.F + p.out0] = *pc;
        ninl[(size_t)s * p.F + p.out0] = n_inl;
    }
}

// ------------------------------------------------------------------------------------ single-shot kernels ----
__global__ void frustum_kernel(const movfe_pose *__restrict__ poses, const movfe_map_point *__restrict__ pts,
                               const int32_t *__restrict__ off, movfe_camera cam, int W, int H, float cosLimit,
                               movfe_projection *__restrict__ out) {
    const int pidx = blockIdx.y;
    const FrustumPose fp = frustum_pose(poses[pidx]);
    const int b = off[pidx], e = off[pidx + 1];
    for (int i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x)
        out[i] = frustum_point(fp, cam, W, H, cosLimit, pts[i], false);
}

__global__ void __launch_bounds__(TP_THREADS)
join_kernel(const int32_t *__restrict__ track_ids, const int32_t *__restrict__ track_off, const int32_t *__restrict__ probe_ids,
            const uint8_t *__restrict__ probe_valid, const int32_t *__restrict__ probe_off, int32_t *__restrict__ match,
            int32_t *__restrict__ n_matches, int hash_cap_max) {
    extern __shared__ int hsm[];
    __shared__ int red[TP_WARPS];
    const int pidx = blockIdx.x;
    const int tb = track_off[pidx], n = track_off[pidx + 1] - tb;
    const int pb = probe_off[pidx], m = probe_off[pidx + 1] - pb;
    int cap = 2;
    while (cap < 2 * n) cap <<= 1;
    int *keys = hsm, *vals = hsm + hash_cap_max, *hit = hsm + 2 * hash_cap_max;
    hash_build([&](int t) { return track_ids[tb + t]; }, n, keys, vals, cap);
    for (int t = threadIdx.x; t < n; t += blockDim.x) hit[t] = -1;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        if (!probe_valid[pb + i]) continue;
        const int t = hash_find(probe_ids[pb + i], keys, vals, cap);
        if (t >= 0) {
            atomicMax(&hit[t], i);
            cnt++;
        }
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x)
        if (hit[t] >= 0) match[tb + t] = hit[t];
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < TP_WARPS; w++) tot += red[w];
        n_matches[pidx] = tot;
    }
}

__global__ void __launch_bounds__(TP_THREADS)
pose_kernel(const float *__restrict__ pts, const float *__restrict__ obs, const int32_t *__restrict__ off, movfe_camera cam,
            movfe_pose_params pp, movfe_pose *__restrict__ poses, uint8_t *__restrict__ outlier, int32_t *__restrict__ n_inl,
            int32_t *__restrict__ stats) {
    __shared__ Solver2Shared sh;
    const int pidx = blockIdx.x;
    const int b = off[pidx], n = off[pidx + 1] - b;
    DirectSrc src{pts + 3 * (size_t)b, obs + 2 * (size_t)b};
    int st[4], passes;
    const int r = pose_solve2(src, n, cam, pp, poses + pidx, outlier + b, st, passes, sh);  // outlier flags live in global memory here
    if (threadIdx.x == 0) {
        n_inl[pidx] = r;
        if (stats) {
#pragma unroll
            for (int k = 0; k < 4; k++) stats[4 * pidx + k] = st[k];
        }
    }
}

// ------------------------------------------------------------------ PoseOptimization by a thread-block cluster -----
// Large problems (20 000 correspondences per frame) with one CTA each leave the FP64 pipes of an SM to
// eight warps and 20 SMs of the chip to nobody (128 problems per GPU). Here a CLUSTER of CTAs owns a problem: every CTA
// accumulates its share of the correspondences, the per-warp partial sums stay in each CTA's shared memory, and after ONE
// cluster barrier per pass every warp of every CTA adds all of them - its own CTA's and, through distributed shared memory,
// the other CTAs' - in the same fixed order and solves the 6x6 system itself, so all CTAs walk the same poses without a
// word of global traffic. Same pass schedule as pose_solve2 (classification fused into the next round's first pass).
template <typename Src>
__device__ int pose_solve_cluster(const Src &src, int n, const movfe_camera &cam_, const movfe_pose_params &pp, movfe_pose *pose,
                                  uint8_t *outlier, int (&stats)[4], Solver2Shared &sh) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = TP_WARPS;
    const CamD cam = widen(cam_);
    const float repErrorF = pp.is_lost ? (float)pp.reprojection_error_lost : (float)pp.reprojection_error;  // Optimizer.cc:423-427
    const double delta = repErrorF, chi2thr = delta * delta;
    const int its = pp.iteration_count / 4 > 1 ? pp.iteration_count / 4 : 1;
    const int first = rank * TP_THREADS + threadIdx.x, stride = CL * TP_THREADS;  // this thread's correspondences
    double *Rt = sh.Rt[warp];
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    for (int i = first; i < n; i += stride) outlier[i] = 0;
    if (lane < 9) Rt[lane] = pose->R[lane];
    else if (lane < 12) Rt[lane] = pose->t[lane - 9];
    // the other CTAs' partial-sum buffers (distributed shared memory)
    const double *rpart[8];
    const int *ripart[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        rpart[r] = r < CL ? cluster.map_shared_rank(&sh.part[0][0][0], r) : nullptr;
        ripart[r] = r < CL ? cluster.map_shared_rank(&sh.ipart[0][0], r) : nullptr;
    }
    cluster.sync();  // every CTA has read the initial pose before the last pass's writer (rank 0) can overwrite it
    if (n < 4) return 0;  // Optimizer.cc:415-418
    int n_bad = 0, pass = 0;
    bool pending = false, stop = false;
    for (int round = 0; round < 4 && !stop; round++) {
        const bool robust = round < 3;
        for (int it = 0; it < its; it++) {
            double acc[27];
#pragma unroll
            for (int q = 0; q < 27; q++) acc[q] = 0.0;
            int bad = 0;
            for (int i = first; i < n; i += stride) {
                float X0, X1, X2, ou, ov;
                src.get(i, X0, X1, X2, ou, ov);
                if (pending) {
                    const bool b = classify_point(cam, Rt, Rt + 9, X0, X1, X2, ou, ov, chi2thr);
                    outlier[i] = b ? 1 : 0;  // only this thread reads it back in later passes
                    bad += b;
                    if (b) continue;
                } else if (outlier[i]) {
                    continue;
                }
                accumulate_point(cam, Rt, Rt + 9, X0, X1, X2, ou, ov, robust, delta, acc);
            }
            const double v = warp_reduce27(acc, lane);
            sh.part[pass & 1][warp][lane] = v;
            if (pending) {
                for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
                if (lane == 0) sh.ipart[pass & 1][warp] = bad;
            }
            cluster.sync();  // the only barrier of a pass; partials are double-buffered by pass parity (pose_solve2)
            if (lane < 27) {
                double tot = 0;
                for (int r = 0; r < CL; r++)
                    for (int w = 0; w < NW; w++) tot += rpart[r][((pass & 1) * NW + w) * 32 + lane];
                sh.tot[warp][lane] = tot;
            }
            if (pending) {
                n_bad = 0;
                for (int r = 0; r < CL; r++)
                    for (int w = 0; w < NW; w++) n_bad += ripart[r][(pass & 1) * NW + w];
                stats[1]++;
                pending = false;
                if (n - n_bad < 3) {
                    stop = true;
                    pass++;
                    break;
                }
            }
            pass++;
            __syncwarp();
            stats[0]++;
            double dx[6];
            int flag;
            if (!solve6_d(sh.tot[warp], &sh.tot[warp][21], dx)) {
                flag = 2;
                stats[3]++;
            } else {
                double dR[9], dt[3], Rn[9], tn[3];
                se3_exp_d(dx, dR, dt);
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < 3; j++) Rn[i * 3 + j] = dR[i * 3] * Rt[j] + dR[i * 3 + 1] * Rt[3 + j] + dR[i * 3 + 2] * Rt[6 + j];
#pragma unroll
                for (int r = 0; r < 3; r++) tn[r] = dR[r * 3] * Rt[9] + dR[r * 3 + 1] * Rt[10] + dR[r * 3 + 2] * Rt[11] + dt[r];
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 9; i++) Rt[i] = Rn[i];
#pragma unroll
                    for (int i = 0; i < 3; i++) Rt[9 + i] = tn[i];
                }
                double m = 0;
#pragma unroll
                for (int a = 0; a < 6; a++) m = fmax(m, fabs(dx[a]));
                flag = m < 1e-10 ? 1 : 0;
            }
            __syncwarp();
            if (flag) break;
        }
        if (!stop) pending = true;
    }
    if (pending) {  // the classification that closes the last round
        int bad = 0;
        for (int i = first; i < n; i += stride) {
            float X0, X1, X2, ou, ov;
            src.get(i, X0, X1, X2, ou, ov);
            const bool b = classify_point(cam, Rt, Rt + 9, X0, X1, X2, ou, ov, chi2thr);
            outlier[i] = b ? 1 : 0;
            bad += b;
        }
        for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        if (lane == 0) sh.ipart[pass & 1][warp] = bad;
        cluster.sync();
        n_bad = 0;
        for (int r = 0; r < CL; r++)
            for (int w = 0; w < NW; w++) n_bad += ripart[r][(pass & 1) * NW + w];
        stats[1]++;
    }
    stats[2] = stats[0] + stats[1];
    if (rank == 0) {
        if (threadIdx.x < 9) pose->R[threadIdx.x] = Rt[threadIdx.x];
        else if (threadIdx.x < 12) pose->t[threadIdx.x - 9] = Rt[threadIdx.x];
    }
    cluster.sync();  // no CTA leaves while another may still read its shared memory
    return n - n_bad;
}

__global__ void __launch_bounds__(TP_THREADS, MOVFE_TP_MINB)
pose_cluster_kernel(const float *__restrict__ pts, const float *__restrict__ obs, const int32_t *__restrict__ off, movfe_camera cam,
                    movfe_pose_params pp, movfe_pose *__restrict__ poses, uint8_t *__restrict__ outlier, int32_t *__restrict__ n_inl,
                    int32_t *__restrict__ stats) {
    __shared__ Solver2Shared sh;
    namespace cg = cooperative_groups;
    const int CL = (int)cg::this_cluster().num_blocks();
    const int pidx = blockIdx.x / CL;
    const int b = off[pidx], n = off[pidx + 1] - b;
    DirectSrc src{pts + 3 * (size_t)b, obs + 2 * (size_t)b};
    int st[4];
    const int r = pose_solve_cluster(src, n, cam, pp, poses + pidx, outlier + b, st, sh);
    if (cg::this_cluster().block_rank() == 0 && threadIdx.x == 0) {
        n_inl[pidx] = r;
        if (stats) {
#pragma unroll
            for (int k = 0; k < 4; k++) stats[4 * pidx + k] = st[k];
        }
    }
}

int pow2_at_least(int v) {
    int c = 2;
    while (c < v) c <<= 1;
    return c;
}

}  // namespace

static size_t pose_tag_bytes(const movfe_ctx *ctx) {
    return ((size_t)ctx->cfg.n_streams * std::max(ctx->cfg.max_map_points, 1) * sizeof(int32_t) + 255) & ~(size_t)255;  // skip tags (first form)
}
size_t movfe_pose_scratch_bytes(const movfe_ctx *ctx) {
    // + first track / group leader of every map point, per frame of a window (tp_prep_kernel)
    return pose_tag_bytes(ctx) + (size_t)ctx->cfg.n_streams * ctx->cfg.window_frames * std::max(ctx->cfg.max_map_points, 1) * sizeof(uint32_t);
}

int movfe_ensure_op_scratch(movfe_ctx *ctx, size_t bytes);

// Dynamic shared memory of every kernel of this file is opted in to the device limit once, at create: the attribute is per
// function and process-wide, so setting it per launch would race between contexts on different host threads.
int movfe_pose_init(movfe_ctx *ctx) {
    MOVFE_CUDA(ctx, optin_dynamic_smem(track_poses_kernel, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(track_poses2_kernel, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(tp_prep_kernel, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(tp_join_kernel, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(tp_solve_kernel, ctx->smem_optin));
    MOVFE_CUDA(ctx, optin_dynamic_smem(join_kernel, ctx->smem_optin));
    return MOVFE_OK;
}

int movfe_track_poses_launch(movfe_ctx *ctx, int64_t first_frame, int n_frames) {
    const movfe_config &c = ctx->cfg;
    TrackPoseParams p;
    p.S = c.n_streams;
    p.W = c.width;
    p.H = c.height;
    p.maxT = c.max_tracks;
    p.maxMap = std::max(c.max_map_points, 1);
    p.TSLOTS = ctx->TSLOTS;
    p.F = c.window_frames;
    p.hash_cap = pow2_at_least(2 * p.maxMap);
    p.n_frames = n_frames;
    p.tslot0 = (int)(first_frame % p.TSLOTS);  // == tslot_of(first_frame) in extract.cu
    p.out0 = 0;
    p.view_cos = ctx->view_cos;
    p.cam = ctx->cam;
    p.pp = ctx->pp;
    const size_t smem = ((size_t)2 * p.hash_cap + p.maxMap) * sizeof(int) + (size_t)p.maxMap * 25;
    int optin = 0;
    cudaFuncAttributes fa;
    MOVFE_CUDA(ctx, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device));
    MOVFE_CUDA(ctx, cudaFuncGetAttributes(&fa, track_poses_kernel));
    if (smem + fa.sharedSizeBytes > (size_t)optin)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "track_poses: max_tracks=%d / max_map_points=%d need %zu bytes of shared memory per stream (limit %zu)",
                   c.max_tracks, c.max_map_points, smem, (size_t)optin - fa.sharedSizeBytes);
    // one chain per frame, each released by the event recorded after that frame's finalize: the pose chain of frame f
    // runs beside the propagation of frame f+1 instead of after the window
    const bool split = ctx->pose_split && ctx->d_pairs != nullptr && ctx->h_nmap_max <= 4096;  // larger maps: wide fused CTAs
    const size_t smem_join = ((size_t)2 * p.hash_cap + p.maxMap) * sizeof(int) + (size_t)p.maxMap;
    const size_t smem_solve = (size_t)p.maxMap * 21;
    // solver CTAs are sized by the correspondence count, which only the device knows: one launch per size class, the CTAs
    // of the other class leave at once (the classes a context can need follow from the largest local map installed)
    const int n_cls = ctx->h_nmap_max <= 64 ? 1 : 2;
    // second form (default): one wide preparation launch for the frames of the call, then one persistent CTA per stream
    const size_t smem_prep = (size_t)3 * p.hash_cap * sizeof(int);
    const size_t smem2 = ((size_t)10 * p.maxMap + 2 * (size_t)((p.maxT + 31) / 32)) * sizeof(int) + (size_t)2 * p.maxMap;
    cudaFuncAttributes fa2;
    MOVFE_CUDA(ctx, cudaFuncGetAttributes(&fa2, track_poses2_kernel));
    const bool second = !ctx->pose_v1 && !split && p.maxMap <= 0xffff && p.maxT < 0xffff && smem_prep <= (size_t)optin &&
                        smem2 + fa2.sharedSizeBytes <= (size_t)optin;
    if (second) {
        // a few frames per launch pair: the chain of the first frames of a window starts while propagation is still on the
        // later ones, and what is left after the window's last table is one group, not a whole window
        uint32_t *ft = (uint32_t *)((uint8_t *)ctx->d_pose_scratch + pose_tag_bytes(ctx));
        for (int k0 = 0; k0 < n_frames; k0 += ctx->pose_group) {
            const int ng = std::min(ctx->pose_group, n_frames - k0);
            for (int k = k0; k < k0 + ng; k++)
                for (int g = 0; g < ctx->n_groups; g++)  // every group of streams has finished the tables of these frames
                    MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->pose_stream, ctx->ev_frame[(size_t)g * c.window_frames + ctx->ev_of_frame[(first_frame + k) % c.window_frames]], 0));
            ProfScope prof(ctx, MOVFE_STAGE_POSE, ctx->pose_stream);
            prof.launches(2);
            p.n_frames = ng;
            p.tslot0 = (int)((first_frame + k0) % p.TSLOTS);
            p.out0 = k0;
            tp_prep_kernel<<<dim3(ng, c.n_streams), TP_THREADS, smem_prep, ctx->pose_stream>>>(p, ctx->d_tracks, ctx->d_ntracks, ctx->d_map, ctx->d_nmap, ft,
                                                                                                ctx->d_match, ctx->d_outlier);
            track_poses2_kernel<<<c.n_streams, TP_THREADS, smem2, ctx->pose_stream>>>(p, ctx->d_tracks, ctx->d_ntracks, ctx->d_map, ctx->d_nmap, ctx->d_nkf, ft,
                                                                                      ctx->d_pose_cur, ctx->d_poses, ctx->d_ninl, ctx->d_match, ctx->d_outlier,
                                                                                      ctx->d_stats);
        }
    }
    for (int k = 0; k < n_frames && !second; k++) {
        for (int g = 0; g < ctx->n_groups; g++)  // every group of streams has finished this frame's table
            MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->pose_stream, ctx->ev_frame[(size_t)g * c.window_frames + ctx->ev_of_frame[(first_frame + k) % c.window_frames]], 0));
        p.n_frames = 1;
        p.tslot0 = (int)((first_frame + k) % p.TSLOTS);
        p.out0 = k;
        ProfScope prof(ctx, MOVFE_STAGE_POSE, ctx->pose_stream);
        if (!split) {
            prof.launches(1);
            track_poses_kernel<<<c.n_streams, TP_THREADS, smem, ctx->pose_stream>>>(p, ctx->d_tracks, ctx->d_ntracks, ctx->d_map, ctx->d_nmap,
                                                                                    ctx->d_nkf, ctx->d_pose_cur, ctx->d_poses, ctx->d_ninl,
                                                                                    ctx->d_match, ctx->d_outlier, (int32_t *)ctx->d_pose_scratch, ctx->d_stats);
        } else {
            prof.launches(2 + 2 * n_cls);
            int32_t *tags = (int32_t *)ctx->d_pose_scratch;
            for (int phase = 0; phase < 2; phase++) {
                tp_join_kernel<<<c.n_streams, TP_THREADS, smem_join, ctx->pose_stream>>>(p, phase, ctx->d_tracks, ctx->d_ntracks, ctx->d_map,
                                                                                         ctx->d_nmap, ctx->d_nkf, ctx->d_pose_cur, ctx->d_match,
                                                                                         ctx->d_outlier, tags, ctx->d_pairs, ctx->d_npairs);
                for (int cls = 0; cls < n_cls; cls++)
                    tp_solve_kernel<<<c.n_streams, cls == 0 ? 64 : TP_SOLVE_THREADS, smem_solve, ctx->pose_stream>>>(
                        p, phase, cls == 0 ? 0 : 65, cls == 0 ? 64 : 0x7fffffff, ctx->d_ntracks, ctx->d_nmap, ctx->d_pose_cur, ctx->d_poses,
                        ctx->d_ninl, ctx->d_outlier, tags, ctx->d_pairs, ctx->d_npairs);
            }
        }
    }
    movfe_ctx::PoseLaunch &pl = ctx->pose_launches[ctx->pose_launch_head];
    ctx->pose_launch_head = (ctx->pose_launch_head + 1) % 4;
    pl.first = first_frame;
    pl.n = n_frames;
    MOVFE_CUDA(ctx, cudaEventRecord(pl.done, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaGetLastError());
    return MOVFE_OK;
}

// ----------------------------------------------------------------------------------------------- C-ABI --------
extern "C" int movfe_set_camera(movfe_ctx *ctx, const movfe_camera *cam, const movfe_pose_params *pp, float viewing_cos_limit) {
    if (!ctx || !cam || !pp) return MOVFE_E_INVALID;
    if (cam->model != MOVFE_CAM_PINHOLE && cam->model != MOVFE_CAM_FISHEYE) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "unknown camera model %d", cam->model);
    ctx->cam = *cam;
    ctx->pp = *pp;
    ctx->view_cos = viewing_cos_limit;
    return MOVFE_OK;
}

extern "C" int movfe_set_map_points(movfe_ctx *ctx, int stream, const movfe_map_point *pts, int n, int n_keyframe_points) {
    if (!ctx) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    if (stream < 0 || stream >= c.n_streams || n < 0 || (n > 0 && !pts) || n_keyframe_points < 0)
        MOVFE_FAIL(ctx, MOVFE_E_INVALID, "set_map_points: bad argument");
    if (n > c.max_map_points) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "set_map_points: %d points, capacity %d", n, c.max_map_points);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    if (n > 0)
        MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_map + (size_t)stream * std::max(c.max_map_points, 1), pts, (size_t)n * sizeof(movfe_map_point),
                                        cudaMemcpyHostToDevice, ctx->pose_stream));
    const int32_t nn = n, nk = std::min(n_keyframe_points, n);
    ctx->h_nmap_max = std::max(ctx->h_nmap_max, n);
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_nmap + stream, &nn, 4, cudaMemcpyHostToDevice, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_nkf + stream, &nk, 4, cudaMemcpyHostToDevice, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    return MOVFE_OK;
}

namespace {
// packed maps of all streams -> the per-stream tables (one CTA per stream, 32-bit words: a map point is 10 of them)
__global__ void map_install_kernel(const uint32_t *__restrict__ src, const int64_t *__restrict__ off, const int32_t *__restrict__ n_kf,
                                   int max_map, uint32_t *__restrict__ d_map, int32_t *__restrict__ d_nmap, int32_t *__restrict__ d_nkf) {
    const int s = blockIdx.x;
    const int64_t o = off[s];
    const int n = (int)min((int64_t)max_map, off[s + 1] - o);
    const uint32_t *from = src + o * 10;
    uint32_t *to = d_map + (size_t)s * max_map * 10;
    for (int i = threadIdx.x; i < n * 10; i += blockDim.x) to[i] = from[i];
    if (threadIdx.x == 0) {
        d_nmap[s] = n;
        d_nkf[s] = min(n_kf[s], n);
    }
}
// ------------------------------------------------------------------------ local map building (UpdateLocalPoints) -----
// One CTA per stream. The list is walked twice: first every non-NULL, non-bad entry bids for its point with its list position
// (integer atomicMin: the first occurrence wins whatever the thread order), then the entries that own their point are compacted
// in list order into the stream's local map; a third walk puts the stamps back.
constexpr int LP_THREADS = 1024;
__global__ void __launch_bounds__(LP_THREADS)
local_points_kernel(const movfe_map_point *__restrict__ store, int32_t *__restrict__ stamp, int store_cap, const int32_t *__restrict__ idx,
                    const int64_t *__restrict__ off, const int32_t *__restrict__ n_kf_entries, int max_map, movfe_map_point *__restrict__ d_map,
                    int32_t *__restrict__ d_nmap, int32_t *__restrict__ d_nkf) {
    __shared__ int wsum[LP_THREADS / 32];
    __shared__ int s_total, s_first;
    const int s = blockIdx.x;
    const movfe_map_point *st = store + (size_t)s * store_cap;
    int32_t *sp = stamp + (size_t)s * store_cap;
    const int32_t *L = idx + off[s];
    const int m = (int)(off[s + 1] - off[s]), n_first = n_kf_entries[s];
    movfe_map_point *out = d_map + (size_t)s * max_map;
    auto valid = [&](int j, int &i) {
        i = L[j];
        return i >= 0 && i < store_cap && !(st[i].flags & MOVFE_MP_BAD);  // :1187-1191
    };
    for (int j = threadIdx.x; j < m; j += LP_THREADS) {
        int i;
        if (valid(j, i)) atomicMin(&sp[i], j);
    }
    if (threadIdx.x == 0) {
        s_total = 0;
        s_first = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int base = 0; base < m; base += LP_THREADS) {
        const int j = base + threadIdx.x;
        int i = -1;
        const bool keep = j < m && valid(j, i) && sp[i] == j;  // first occurrence (mnTrackReferenceForFrame, :1189-1195)
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int before = s_total, tot = 0;
        for (int w = 0; w < LP_THREADS / 32; w++) {
            const int c = wsum[w];
            before += w < warp ? c : 0;
            tot += c;
        }
        const int pos = before + __popc(b & ((1u << lane) - 1u));
        if (keep && pos < max_map) {
            out[pos] = st[i];  // mvpLocalMapPoints.push_back(pMP) (:1194)
            if (j < n_first) atomicMax(&s_first, pos + 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_total += tot;
        __syncthreads();
    }
    for (int j = threadIdx.x; j < m; j += LP_THREADS) {
        int i;
        if (valid(j, i)) sp[i] = 0x7fffffff;
    }
    if (threadIdx.x == 0) {
        d_nmap[s] = min(s_total, max_map);
        d_nkf[s] = s_first;
    }
}
}  // namespace

extern "C" int movfe_reserve_map_store(movfe_ctx *ctx, int max_points_per_stream) {
    if (!ctx || max_points_per_stream < 1) return MOVFE_E_INVALID;
    if (max_points_per_stream <= ctx->store_cap) return MOVFE_OK;
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    const size_t S = ctx->cfg.n_streams, n = S * (size_t)max_points_per_stream;
    movfe_map_point *ns = nullptr;
    int32_t *nst = nullptr;
    MOVFE_CUDA(ctx, cudaMalloc(&ns, n * sizeof(movfe_map_point)));
    MOVFE_CUDA(ctx, cudaMalloc(&nst, n * sizeof(int32_t)));
    MOVFE_CUDA(ctx, cudaMemsetAsync(ns, 0xff, n * sizeof(movfe_map_point), ctx->pose_stream));  // flags all set: every slot starts as a bad point
    MOVFE_CUDA(ctx, cudaMemsetAsync(nst, 0x7f, n * sizeof(int32_t), ctx->pose_stream));         // 0x7f7f7f7f: above every list position
    if (ctx->d_store)  // keep what was installed
        MOVFE_CUDA(ctx, cudaMemcpy2DAsync(ns, (size_t)max_points_per_stream * sizeof(movfe_map_point), ctx->d_store,
                                          (size_t)ctx->store_cap * sizeof(movfe_map_point), (size_t)ctx->store_cap * sizeof(movfe_map_point), S,
                                          cudaMemcpyDeviceToDevice, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    if (ctx->d_store) cudaFree(ctx->d_store);
    if (ctx->d_store_stamp) cudaFree(ctx->d_store_stamp);
    ctx->d_store = ns;
    ctx->d_store_stamp = nst;
    ctx->store_cap = max_points_per_stream;
    return MOVFE_OK;
}

extern "C" int movfe_set_map_store(movfe_ctx *ctx, int stream, int first_index, const movfe_map_point *pts, int n) {
    if (!ctx) return MOVFE_E_INVALID;
    if (stream < 0 || stream >= ctx->cfg.n_streams || first_index < 0 || n < 0 || (n > 0 && !pts)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "set_map_store: bad argument");
    if (first_index + n > ctx->store_cap)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "set_map_store: points [%d, %d) exceed the reserved store of %d points per stream (movfe_reserve_map_store)",
                   first_index, first_index + n, ctx->store_cap);
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (n > 0) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_store + (size_t)stream * ctx->store_cap + first_index, pts, (size_t)n * sizeof(movfe_map_point),
                                        cudaMemcpyHostToDevice, ctx->pose_stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));  // the array is the caller's
    }
    return MOVFE_OK;
}

extern "C" int movfe_update_local_points(movfe_ctx *ctx, const int32_t *idx, const int64_t *off, const int32_t *n_keyframe_entries) {
    if (!ctx || !off || !n_keyframe_entries) return MOVFE_E_INVALID;
    if (!ctx->d_store) MOVFE_FAIL(ctx, MOVFE_E_STATE, "update_local_points: no map store (movfe_reserve_map_store / movfe_set_map_store)");
    const movfe_config &c = ctx->cfg;
    const int S = c.n_streams;
    const int64_t total = off[S];
    if (total < 0 || (total > 0 && !idx)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "update_local_points: bad offsets");
    for (int s = 0; s < S; s++)
        if (off[s + 1] < off[s] || n_keyframe_entries[s] < 0) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "update_local_points: bad list of stream %d", s);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    cudaStream_t st = ctx->pose_stream;  // ordered with the pose chains that read the local maps
    const size_t meta = (size_t)(S + 1) * 8 + (size_t)S * 4;
    if (ctx->lp_idx_cap < (size_t)total || !ctx->d_lp_off) {
        MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->d_lp_idx) cudaFree(ctx->d_lp_idx);
        ctx->d_lp_idx = nullptr;
        ctx->lp_idx_cap = 0;
        const size_t want = (size_t)total + (size_t)total / 2 + 1024;
        MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_lp_idx, want * sizeof(int32_t)));
        ctx->lp_idx_cap = want;
        if (!ctx->d_lp_off) MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_lp_off, meta));
        if (!ctx->h_lp_meta) MOVFE_CUDA(ctx, cudaMallocHost(&ctx->h_lp_meta, meta));
    } else {
        MOVFE_CUDA(ctx, cudaStreamSynchronize(st));  // the previous call's copies out of the pinned buffer have completed
    }
    memcpy(ctx->h_lp_meta, off, (size_t)(S + 1) * 8);
    memcpy((uint8_t *)ctx->h_lp_meta + (size_t)(S + 1) * 8, n_keyframe_entries, (size_t)S * 4);
    if (total > 0) MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_lp_idx, idx, (size_t)total * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_lp_off, ctx->h_lp_meta, meta, cudaMemcpyHostToDevice, st));
    local_points_kernel<<<S, LP_THREADS, 0, st>>>(ctx->d_store, ctx->d_store_stamp, ctx->store_cap, ctx->d_lp_idx, ctx->d_lp_off,
                                                  reinterpret_cast<const int32_t *>(ctx->d_lp_off + S + 1), std::max(c.max_map_points, 1), ctx->d_map,
                                                  ctx->d_nmap, ctx->d_nkf);
    MOVFE_CUDA(ctx, cudaGetLastError());
    ctx->h_nmap_max = std::max(ctx->h_nmap_max, c.max_map_points);
    if (total > 0) MOVFE_CUDA(ctx, cudaStreamSynchronize(st));  // idx is the caller's array
    return MOVFE_OK;
}

extern "C" int movfe_download_map_points(movfe_ctx *ctx, int stream, movfe_map_point *out, int capacity, int32_t *n_keyframe_points) {
    if (!ctx) return MOVFE_E_INVALID;
    if (stream < 0 || stream >= ctx->cfg.n_streams || capacity < 0 || (capacity > 0 && !out)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "download_map_points: bad argument");
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int32_t n = 0, nk = 0;
    MOVFE_CUDA(ctx, cudaMemcpyAsync(&n, ctx->d_nmap + stream, 4, cudaMemcpyDeviceToHost, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(&nk, ctx->d_nkf + stream, 4, cudaMemcpyDeviceToHost, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    if (n > capacity) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "download_map_points: %d points, capacity %d", n, capacity);
    if (n > 0) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_map + (size_t)stream * std::max(ctx->cfg.max_map_points, 1), (size_t)n * sizeof(movfe_map_point),
                                        cudaMemcpyDeviceToHost, ctx->pose_stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    }
    if (n_keyframe_points) *n_keyframe_points = nk;
    return n;
}

namespace {
}  // namespace

extern "C" int movfe_set_map_points_batch(movfe_ctx *ctx, const movfe_map_point *pts, const int64_t *off, const int32_t *n_keyframe_points,
                                          int max_points_per_stream, int on_device) {
    static_assert(sizeof(movfe_map_point) == 40, "map point = 10 words");
    if (!ctx || !off || !n_keyframe_points) return MOVFE_E_INVALID;
    const movfe_config &c = ctx->cfg;
    const int S = c.n_streams;
    if (max_points_per_stream < 0 || max_points_per_stream > c.max_map_points)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "set_map_points_batch: %d points per stream, capacity %d", max_points_per_stream, c.max_map_points);
    MOVFE_CUDA(ctx, cudaSetDevice(c.device));
    cudaStream_t st = ctx->pose_stream;  // ordered with the pose chains that read the maps
    const movfe_map_point *d_pts = pts;
    const int64_t *d_off = off;
    const int32_t *d_nkf = n_keyframe_points;
    int staged = -1;
    if (!on_device) {
        int64_t total = off[S];
        for (int s = 0; s < S; s++) {
            const int64_t n = off[s + 1] - off[s];
            if (n < 0 || n > max_points_per_stream)
                MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "set_map_points_batch: stream %d has %lld points, the call allows %d", s, (long long)n, max_points_per_stream);
        }
        if (total > 0 && !pts) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "set_map_points_batch: null points");
        const size_t b_pts = ((size_t)total * sizeof(movfe_map_point) + 255) & ~(size_t)255, b_off = ((size_t)(S + 1) * 8 + 255) & ~(size_t)255;
        const size_t need = b_pts + b_off + (size_t)S * 4;
        const int b = ctx->map_parity;
        ctx->map_parity ^= 1;
        if (!ctx->ev_map_staged[b]) MOVFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_map_staged[b], cudaEventDisableTiming));
        // buffer b was last used two hand-overs ago: its install kernel has normally long finished
        if (ctx->map_stage_bytes[b]) MOVFE_CUDA(ctx, cudaEventSynchronize(ctx->ev_map_staged[b]));
        if (ctx->map_stage_bytes[b] < need) {
            if (ctx->d_map_stage[b]) cudaFree(ctx->d_map_stage[b]);
            if (ctx->h_map_meta[b]) cudaFreeHost(ctx->h_map_meta[b]);
            ctx->d_map_stage[b] = nullptr;
            ctx->h_map_meta[b] = nullptr;
            ctx->map_stage_bytes[b] = 0;
            MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_map_stage[b], need + need / 2));
            MOVFE_CUDA(ctx, cudaMallocHost(&ctx->h_map_meta[b], b_off + (size_t)S * 4));
            ctx->map_stage_bytes[b] = need + need / 2;
        }
        uint8_t *base = (uint8_t *)ctx->d_map_stage[b];
        memcpy(ctx->h_map_meta[b], off, (size_t)(S + 1) * 8);
        memcpy((uint8_t *)ctx->h_map_meta[b] + b_off, n_keyframe_points, (size_t)S * 4);
        if (total > 0) MOVFE_CUDA(ctx, cudaMemcpyAsync(base, pts, (size_t)total * sizeof(movfe_map_point), cudaMemcpyHostToDevice, st));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(base + b_pts, ctx->h_map_meta[b], b_off + (size_t)S * 4, cudaMemcpyHostToDevice, st));
        staged = b;
        d_pts = (const movfe_map_point *)base;
        d_off = (const int64_t *)(base + b_pts);
        d_nkf = (const int32_t *)(base + b_pts + b_off);
    }
    ctx->h_nmap_max = std::max(ctx->h_nmap_max, max_points_per_stream);
    map_install_kernel<<<S, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(d_pts), d_off, d_nkf, std::max(c.max_map_points, 1),
                                          reinterpret_cast<uint32_t *>(ctx->d_map), ctx->d_nmap, ctx->d_nkf);
    MOVFE_CUDA(ctx, cudaGetLastError());
    if (staged >= 0) MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_map_staged[staged], st));
    return MOVFE_OK;
}

extern "C" int movfe_set_pose(movfe_ctx *ctx, int stream, const movfe_pose *pose) {
    if (!ctx || !pose) return MOVFE_E_INVALID;
    if (stream < 0 || stream >= ctx->cfg.n_streams) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "stream %d out of range", stream);
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(ctx->d_pose_cur + stream, pose, sizeof(movfe_pose), cudaMemcpyHostToDevice, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    return MOVFE_OK;
}

extern "C" int movfe_track_poses(movfe_ctx *ctx, int64_t first_frame, int n_frames) {
    if (!ctx) return MOVFE_E_INVALID;
    const int64_t ext_end = ctx->ext_first < 0 ? 0 : ctx->ext_first + ctx->ext_n;
    if (n_frames < 1 || n_frames > ctx->cfg.window_frames || ctx->ext_first < 0 || first_frame < ctx->ext_first || first_frame + n_frames > ext_end)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "track_poses: frames [%lld,%lld) are not inside the last extract call", (long long)first_frame,
                   (long long)(first_frame + n_frames));
    const int64_t next = ctx->pose_first < 0 ? 0 : ctx->pose_first + ctx->pose_n;
    if (first_frame != next)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "track_poses: frames must be consumed in order (expected %lld, got %lld)", (long long)next, (long long)first_frame);
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    int rc = movfe_track_poses_launch(ctx, first_frame, n_frames);
    if (rc) return rc;
    ctx->pose_first = first_frame;
    ctx->pose_n = n_frames;
    return MOVFE_OK;
}

extern "C" int movfe_download_poses(movfe_ctx *ctx, int64_t first_frame, int n_frames, movfe_pose *poses, int32_t *n_inliers) {
    if (!ctx || !poses) return MOVFE_E_INVALID;
    if (ctx->pose_first < 0 || first_frame < ctx->pose_first || first_frame + n_frames > ctx->pose_first + ctx->pose_n || n_frames < 1)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "download_poses: frames are not inside the last track_poses call");
    const int S = ctx->cfg.n_streams, F = ctx->cfg.window_frames;
    const int o = (int)(first_frame - ctx->pose_first);
    MOVFE_CUDA(ctx, cudaMemcpy2DAsync(poses, (size_t)n_frames * sizeof(movfe_pose), ctx->d_poses + o, (size_t)F * sizeof(movfe_pose),
                                      (size_t)n_frames * sizeof(movfe_pose), S, cudaMemcpyDeviceToHost, ctx->pose_stream));
    if (n_inliers)
        MOVFE_CUDA(ctx, cudaMemcpy2DAsync(n_inliers, (size_t)n_frames * 4, ctx->d_ninl + o, (size_t)F * 4, (size_t)n_frames * 4, S,
                                          cudaMemcpyDeviceToHost, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    return MOVFE_OK;
}

extern "C" int movfe_download_matches(movfe_ctx *ctx, int stream, int64_t frame, int32_t *match, uint8_t *outlier, int capacity) {
    if (!ctx) return MOVFE_E_INVALID;
    if (stream < 0 || stream >= ctx->cfg.n_streams) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "stream %d out of range", stream);
    if (ctx->pose_first < 0 || frame < ctx->pose_first || frame >= ctx->pose_first + ctx->pose_n)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "download_matches: frame is not inside the last track_poses call");
    int n = 0;
    int rc = movfe_track_count(ctx, stream, frame, &n, nullptr);
    if (rc) return rc;
    if (n > capacity) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "download_matches: %d tracks, capacity %d", n, capacity);
    const size_t o = ((size_t)stream * ctx->cfg.window_frames + (frame - ctx->pose_first)) * ctx->cfg.max_tracks;
    if (n > 0) {
        if (match) MOVFE_CUDA(ctx, cudaMemcpyAsync(match, ctx->d_match + o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->pose_stream));
        if (outlier) MOVFE_CUDA(ctx, cudaMemcpyAsync(outlier, ctx->d_outlier + o, (size_t)n, cudaMemcpyDeviceToHost, ctx->pose_stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    }
    return n;
}

// single-shot operators: inputs are staged into the context's operator scratch, results copied back
namespace {
struct Carver {
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

extern "C" int movfe_frustum(movfe_ctx *ctx, int n_problems, const movfe_pose *poses, const movfe_map_point *pts,
                             const int32_t *off, movfe_projection *out) {
    if (!ctx || n_problems < 1 || !poses || !off || !out) return MOVFE_E_INVALID;
    const int n = off[n_problems];
    if (n < 0 || (n > 0 && !pts)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "frustum: bad offsets");
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t need = (size_t)n_problems * sizeof(movfe_pose) + (size_t)n * (sizeof(movfe_map_point) + sizeof(movfe_projection)) +
                        (size_t)(n_problems + 1) * 4 + 4 * 256;
    int rc = movfe_ensure_op_scratch(ctx, need);
    if (rc) return rc;
    Carver cv{(uint8_t *)ctx->d_op};
    movfe_pose *d_pose = cv.take<movfe_pose>(n_problems);
    movfe_map_point *d_pts = cv.take<movfe_map_point>(n);
    int32_t *d_off = cv.take<int32_t>(n_problems + 1);
    movfe_projection *d_out = cv.take<movfe_projection>(n);
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pose, poses, (size_t)n_problems * sizeof(movfe_pose), cudaMemcpyHostToDevice, ctx->stream));
    if (n) MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pts, pts, (size_t)n * sizeof(movfe_map_point), cudaMemcpyHostToDevice, ctx->stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    dim3 g(std::max(1, std::min(64, (n / n_problems + 255) / 256)), n_problems);
    {
        ProfScope prof(ctx, MOVFE_STAGE_POSE);
        prof.launches(1);
        frustum_kernel<<<g, 256, 0, ctx->stream>>>(d_pose, d_pts, d_off, ctx->cam, ctx->cfg.width, ctx->cfg.height, ctx->view_cos, d_out);
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
    if (n) MOVFE_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(movfe_projection), cudaMemcpyDeviceToHost, ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MOVFE_OK;
}

extern "C" int movfe_join(movfe_ctx *ctx, int n_problems, const int32_t *track_ids, const int32_t *track_off,
                          const int32_t *probe_ids, const uint8_t *probe_valid, const int32_t *probe_off, int32_t *match,
                          int32_t *n_matches) {
    if (!ctx || n_problems < 1 || !track_off || !probe_off || !match || !n_matches) return MOVFE_E_INVALID;
    const int nt = track_off[n_problems], np = probe_off[n_problems];
    int max_n = 0;
    for (int i = 0; i < n_problems; i++) max_n = std::max(max_n, track_off[i + 1] - track_off[i]);
    {   // the hash table of one problem lives in shared memory: (2 * pow2(2n) + n) ints
        const size_t need_smem = ((size_t)2 * pow2_at_least(2 * std::max(max_n, 1)) + std::max(max_n, 1)) * sizeof(int);
        if (need_smem > (size_t)ctx->smem_optin)
            MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "join: %d tracks in one problem need %zu bytes of shared memory (device limit %d: about 8192 tracks)",
                       max_n, need_smem, ctx->smem_optin);
    }
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t need = (size_t)nt * 8 + (size_t)np * 5 + (size_t)(n_problems + 1) * 12 + 8 * 256;
    int rc = movfe_ensure_op_scratch(ctx, need);
    if (rc) return rc;
    Carver cv{(uint8_t *)ctx->d_op};
    int32_t *d_tid = cv.take<int32_t>(nt), *d_toff = cv.take<int32_t>(n_problems + 1);
    int32_t *d_pid = cv.take<int32_t>(np), *d_poff = cv.take<int32_t>(n_problems + 1);
    uint8_t *d_pv = cv.take<uint8_t>(np);
    int32_t *d_match = cv.take<int32_t>(nt), *d_nm = cv.take<int32_t>(n_problems);
    if (nt) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_tid, track_ids, (size_t)nt * 4, cudaMemcpyHostToDevice, ctx->stream));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_match, match, (size_t)nt * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (np) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pid, probe_ids, (size_t)np * 4, cudaMemcpyHostToDevice, ctx->stream));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pv, probe_valid, (size_t)np, cudaMemcpyHostToDevice, ctx->stream));
    }
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_toff, track_off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_poff, probe_off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int cap = pow2_at_least(2 * std::max(max_n, 1));
    const size_t smem = ((size_t)2 * cap + std::max(max_n, 1)) * sizeof(int);
    {
        ProfScope prof(ctx, MOVFE_STAGE_POSE);
        prof.launches(1);
        join_kernel<<<n_problems, TP_THREADS, smem, ctx->stream>>>(d_tid, d_toff, d_pid, d_pv, d_poff, d_match, d_nm, cap);
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
    if (nt) MOVFE_CUDA(ctx, cudaMemcpyAsync(match, d_match, (size_t)nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(n_matches, d_nm, (size_t)n_problems * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MOVFE_OK;
}

extern "C" int movfe_pose_optimize(movfe_ctx *ctx, int n_problems, const movfe_camera *cam, const movfe_pose_params *pp,
                                   const float *pts, const float *obs, const int32_t *off, movfe_pose *poses, uint8_t *outlier,
                                   int32_t *n_inliers, int32_t *stats) {
    if (!ctx || n_problems < 1 || !cam || !pp || !off || !poses || !outlier || !n_inliers) return MOVFE_E_INVALID;
    const int n = off[n_problems];
    if (n < 0 || (n > 0 && (!pts || !obs))) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "pose_optimize: bad offsets");
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t need = (size_t)n * 21 + (size_t)n_problems * (sizeof(movfe_pose) + 24) + 8 * 256;
    int rc = movfe_ensure_op_scratch(ctx, need);
    if (rc) return rc;
    Carver cv{(uint8_t *)ctx->d_op};
    float *d_pts = cv.take<float>(3 * (size_t)n), *d_obs = cv.take<float>(2 * (size_t)n);
    int32_t *d_off = cv.take<int32_t>(n_problems + 1);
    movfe_pose *d_pose = cv.take<movfe_pose>(n_problems);
    uint8_t *d_out = cv.take<uint8_t>(n);
    int32_t *d_ninl = cv.take<int32_t>(n_problems), *d_stats = cv.take<int32_t>(4 * (size_t)n_problems);
    if (n) {
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pts, pts, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_obs, obs, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_off, off, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(d_pose, poses, (size_t)n_problems * sizeof(movfe_pose), cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfScope prof(ctx, MOVFE_STAGE_POSE);
        prof.launches(1);
        // large problems: a cluster of CTAs per problem (pose_solve_cluster); MOVFE_POSE_CLUSTER=n forces the size (1 = off)
        int max_n = 0;
        for (int i = 0; i < n_problems; i++) max_n = std::max(max_n, off[i + 1] - off[i]);
        int cl = max_n >= 4096 ? ((int64_t)n_problems * 2 <= 2 * (int64_t)ctx->sm_count ? 2 : 1) : 1;
        if (const char *e = getenv("MOVFE_POSE_CLUSTER")) cl = std::max(1, std::min(8, atoi(e)));
        if (cl > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(n_problems * cl));
            cfg.blockDim = dim3(TP_THREADS);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)cl;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            MOVFE_CUDA(ctx, cudaLaunchKernelEx(&cfg, pose_cluster_kernel, (const float *)d_pts, (const float *)d_obs, (const int32_t *)d_off, *cam, *pp,
                                               d_pose, d_out, d_ninl, d_stats));
        } else {
            pose_kernel<<<n_problems, TP_THREADS, 0, ctx->stream>>>(d_pts, d_obs, d_off, *cam, *pp, d_pose, d_out, d_ninl, d_stats);
        }
    }
    MOVFE_CUDA(ctx, cudaGetLastError());
    MOVFE_CUDA(ctx, cudaMemcpyAsync(poses, d_pose, (size_t)n_problems * sizeof(movfe_pose), cudaMemcpyDeviceToHost, ctx->stream));
    if (n) MOVFE_CUDA(ctx, cudaMemcpyAsync(outlier, d_out, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    MOVFE_CUDA(ctx, cudaMemcpyAsync(n_inliers, d_ninl, (size_t)n_problems * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (stats) MOVFE_CUDA(ctx, cudaMemcpyAsync(stats, d_stats, (size_t)n_problems * 16, cudaMemcpyDeviceToHost, ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MOVFE_OK;
}
