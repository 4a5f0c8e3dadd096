// Dev probe: one bulk tensor copy of an unaligned 32x16-byte box out of a pitched byte plane (the propagation kernel's
// patch fetch), checked against the source. nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#ifdef USE_CUTE
#include <cute/arch/copy_sm90_tma.hpp>
#endif
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <dlfcn.h>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bisect kernels: mode 1 = mbarrier only, mode 2 = 1-D bulk copy, mode 3 = tensor copy without .tile
__global__ void bisect(int mode, int txb, const __grid_constant__ CUtensorMap pmap, const uint8_t *src, int x, int y, uint8_t *out, int *status, int dstoff) {
    __shared__ __align__(128) uint8_t boxbuf[1024];
    uint8_t *box = boxbuf + dstoff;   // DSTOFF: destination alignment experiments (the PTX manual asks for 128 bytes)
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode == 1) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        } else if (mode == 2) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(box)), "l"(src), "r"(512),
                         "r"(smem_u32(&bar))
                         : "memory");
        } else if (mode == 5) {
#ifdef USE_CUTE
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(txb) : "memory");
            cute::SM90_TMA_LOAD_2D::copy(&pmap, &bar, 0ull, box, x, y);
#endif
        } else if (mode == 6) {
            // nothing here: mode 6 issues from a converged warp below (elect.sync), as CUTLASS does
        } else if (mode == 4) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(txb) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(box)),
                         "l"(reinterpret_cast<uint64_t>(&pmap)), "r"(x), "r"(y), "r"(smem_u32(&bar))
                         : "memory");
        } else {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(txb) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(box)),
                         "l"(reinterpret_cast<uint64_t>(&pmap)), "r"(x), "r"(y), "r"(smem_u32(&bar))
                         : "memory");
        }
    }
    if (mode == 6 && threadIdx.x < 32) {
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        if (elected) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(txb) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(box)),
                         "l"(reinterpret_cast<uint64_t>(&pmap)), "r"(x), "r"(y), "r"(smem_u32(&bar))
                         : "memory");
        }
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 20) && !ok; spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) *status = ok;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = box[i];
}

template <bool PARAM>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int x, int y, uint8_t *out, int *status) {
    const CUtensorMap *map = PARAM ? &pmap : gmap;
    __shared__ __align__(128) uint8_t box[512];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(box)),
                     "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(&bar))
                     : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 20) && !ok; spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) *status = ok;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = box[i];
}

int main() {
    const int P = 1024, ROWS = 64;
    std::vector<uint8_t> h(P * ROWS);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 7 + (i / P) * 13) & 0xff);
    uint8_t *d, *dout; int *dst; CUtensorMap *dmap;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 512); cudaMalloc(&dst, 4); cudaMalloc(&dmap, sizeof(CUtensorMap));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    if (getenv("DLSYM")) {
        void *h = dlopen("libcuda.so.1", RTLD_NOW);
        void *f2 = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr;
        typedef CUresult (*CtxFn)(CUcontext *);
        CtxFn getctx = h ? (CtxFn)dlsym(h, "cuCtxGetCurrent") : nullptr;
        CUcontext cx = nullptr;
        if (getctx) getctx(&cx);
        printf("dlsym fn=%p ctx=%p\n", f2, (void *)cx);
        if (f2) fn = f2;
    }
    { int major = 0, minor = 0, drv = 0, rt = 0; cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, 0); cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, 0);
      cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("cc %d.%d driver %d runtime %d\n", major, minor, drv, rt); }
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)ROWS}, strides[1] = {(cuuint64_t)P};
    const int BW = getenv("BOXW") ? atoi(getenv("BOXW")) : 32, BH = getenv("BOXH") ? atoi(getenv("BOXH")) : 16;
    const cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
    const int l2p = getenv("L2P") ? atoi(getenv("L2P")) : 0;
    CUresult r;
    if (getenv("BF16")) {
        const cuuint64_t d2[2] = {(cuuint64_t)P / 2, (cuuint64_t)ROWS};
        const cuuint32_t b2[2] = {64, 4};
        r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, d2, strides, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("bf16 swizzle128 box 64x4: %d\n", (int)r);
    } else
    r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d box %dx%d l2p %d\n", (int)r, BW, BH, l2p);
    { const uint64_t *w = (const uint64_t *)&map; for (int i = 0; i < 16; i++) printf("%016llx%c", (unsigned long long)w[i], i % 4 == 3 ? '\n' : ' '); }
    cudaMemcpy(dmap, &map, sizeof map, cudaMemcpyHostToDevice);
    const int x = getenv("X") ? atoi(getenv("X")) : 37, y = 5;
    if (getenv("BISECT") && getenv("CLUSTER")) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(32);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, bisect, atoi(getenv("BISECT")), BW * BH, map, (const uint8_t *)(d + y * P), x, y, dout, dst, getenv("DSTOFF") ? atoi(getenv("DSTOFF")) : 0);
        printf("launchEx: %s\n", cudaGetErrorString(le));
    } else
    if (getenv("BISECT")) bisect<<<1, 32>>>(atoi(getenv("BISECT")), BW * BH, map, d + y * P, x, y, dout, dst, getenv("DSTOFF") ? atoi(getenv("DSTOFF")) : 0);
    else if (getenv("GLOBAL_MAP")) probe<false><<<1, 32>>>(map, dmap, x, y, dout, dst);
    else probe<true><<<1, 32>>>(map, dmap, x, y, dout, dst);
    e = cudaDeviceSynchronize();
    printf("kernel (%s map): %s\n", getenv("GLOBAL_MAP") ? "global" : "param", cudaGetErrorString(e));
    std::vector<uint8_t> o(512); int st = -1;
    cudaMemcpy(o.data(), dout, 512, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < 16; r2++) for (int c = 0; c < 32; c++) bad += o[r2 * 32 + c] != h[(y + r2) * P + x + c];
    printf("status=%d mismatches=%d first bytes %d %d %d (want %d %d %d)\n", st, bad, o[0], o[1], o[32], h[y * P + x], h[y * P + x + 1], h[(y + 1) * P + x]);
    return 0;
}
