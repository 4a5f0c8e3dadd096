// common.cuh — context layout and helpers shared by the front-end's translation units.
// Device code targets sm_100a only (B200): 148 SMs, 126 MB L2, HBM3e. Nothing here is a dense contraction, so
// there is no tensor-core path; the kernels are HBM / LSU / integer-ALU work.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/movfe.h"

#define MOVFE_WARP 32

// Compact device-side record: the 7 fields of the 40-byte AVMotionVector that the hot path reads
// (VideoDecoder.cc:211-228), repacked to one aligned 128-bit word by the ingest kernel.
struct __align__(16) Rec16 {
    int16_t sx, sy, dx, dy;
    uint8_t w, h;
    int8_t  src_sign;  // sign of AVMotionVector::source (-1, 0, +1)
    uint8_t pad;
    int32_t ref;
};
static_assert(sizeof(Rec16) == 16, "Rec16 must be one 128-bit word");
static_assert(sizeof(movfe_packed_record) == sizeof(Rec16), "movfe_packed_record is the public name of Rec16");

// Inclusive pixel rectangle a hop covers in its target frame's slot grid (VideoDecoder.cc:295-306,330-333).
// Empty rectangles are {0, 32767, -1, -32768} so that no overlap test ever accepts them.
struct __align__(8) HopRect {
    int16_t x0, y0, x1, y1;
};

// ---- fused (grid-free) mode: per-tile cell tables instead of the per-pixel slot grid --------------------------------------
// Propagation looks the slot grid up at a few thousand pixels per frame (one per track, plus the 16-px lattice of a back-fill
// pass) out of W*H. Hops are rectangles, so inside a 32x32 tile the slot vector only changes at block edges: the tile's
// columns fall into a few runs with the same covering set, its rows too, and the slots are constant on every (row run,
// column run) CELL - the raster kernel computes exactly these cells before it expands them to pixels (grid.cu). In fused mode
// (MOVFE_CFG_NO_GRID) it stops there and stores, per tile, the run index of every column and row (64 bytes) and the cells'
// slot vectors; a lookup is two dependent loads. The 16 bytes per pixel of the grid are never written: SURVEY.md 8d charges
// this mode 40 M + 12 Hops for the raster term.
#define MOVFE_TILE_CELLS 1024   // cell capacity of a tile (32 x 32: every pixel its own cell, the worst case)
struct TileCells {
    const int32_t *dim;    // [tiles] column runs | row runs << 8; 0: no hop meets the tile
    const uint8_t *runs;   // [tiles][64] run index of column 0..31, then of row 0..31
    const int4 *cells;     // [tiles][MOVFE_TILE_CELLS] slots of cell (row run * column runs + column run)
    int NT;                // tiles per tile row
};

// Slots of pixel (x, y): what VideoImage::mvi.at<Vec4i>(y, x) holds in the reference.
__device__ __forceinline__ int4 resolve_slots(const TileCells &q, int x, int y);

// Per-frame class counts produced by the count pass (index into cls_cnt[frame][...]):
//   [0..K]        valid P-branch records with ref >= k          -> hop segment k of frame (f-k)
//   [K+1]         valid records that push their block into this frame's kps (not "chained")
//   [K+2..2K+1]   valid chained records with ref == r (r=1..K)  -> kps segment r of frame (f-1-r)
#define MOVFE_MAX_K 10
#define MOVFE_NCLS(K) (2 * (K) + 2)

// Results of one movfe_raster call for frames [first, first+nout) of every stream (window-local index fi = frame - first).
struct RasterBuf {
    int64_t first = -1;
    int     nout = 0, nin = 0;
    int32_t *d_seg_cnt = nullptr;   // [S][NIN][n_rseg][NCLS+2] per-segment class counts, area, rejects (raster.cu)
    int32_t *d_cls_cnt = nullptr;   // [S][NIN][NCLS]
    int64_t *d_area = nullptr;      // [S][NIN]
    int32_t *d_hop_base = nullptr;  // [S][NIN][K+2]
    int32_t *d_kps_base = nullptr;  // [S][NIN][K+2]
    int32_t *d_nhops = nullptr;     // [S][NIN]
    int32_t *d_nkps = nullptr;      // [S][NIN]
    double  *d_cov = nullptr;       // [S][NIN]
    movfe_hop  *d_hops = nullptr;   // [S][F][max_hops]
    HopRect    *d_hop_rect = nullptr;
    movfe_rect *d_kps = nullptr;    // [S][F][max_kps]
    int2    *d_chunk_bbox = nullptr;  // [S][F][max_chunks]  extent of every 32-hop chunk: (ymin | ymax<<16, xmin | xmax<<16)
    int4    *d_grid = nullptr;      // [S][F][H*W]   (grid-output mode)
    int32_t *d_tc_dim = nullptr;    // [S][F][tiles]                    (fused mode: TileCells)
    uint8_t *d_tc_runs = nullptr;   // [S][F][tiles][64]
    int4    *d_tc_cells = nullptr;  // [S][F][tiles][MOVFE_TILE_CELLS]
    cudaEvent_t done = nullptr;      // recorded on raster_stream: the buffer is complete
    cudaEvent_t consumed = nullptr;  // recorded on stream after the last propagation launch that read it
    bool    consumed_valid = false;
};

struct movfe_ctx {
    movfe_config cfg;
    int K = 0, LA = 0, RING = 0, NIN = 0;  // max_ref, look-ahead frames, ring depth, max input frames per window
    int NB = 0, NT = 0;                    // 8-row bands per frame, 32-px tiles per band
    int NTR = 0;                           // 32-row tile rows per frame
    bool fused = false;                    // MOVFE_CFG_NO_GRID: per-tile cell tables instead of the slot grid
    int max_hops = 0, max_kps = 0, max_chunks = 0;
    int rseg = 0, n_rseg = 1;              // records per count/emit segment (multiple of 512), segments per frame
    int sm_count = 0;
    int smem_optin = 0;                    // cudaDevAttrMaxSharedMemoryPerBlockOptin: every dynamic-smem kernel is opted in to it ONCE at create
    cudaStream_t stream = nullptr;
    // join / frustum / pose of a window run on their own stream, concurrently with raster + propagation of the next
    // window (the chains are independent once a window's track tables exist)
    cudaStream_t pose_stream = nullptr;
    cudaEvent_t ev_tables = nullptr;   // scratch event of movfe_fence
    // Propagation is a dependent chain of three short launches per frame, but the video streams are independent: they are
    // split into n_groups groups, each walking its own chain on its own CUDA stream (group 0 on `stream`), so that one
    // group's cand/birth CTAs fill the tail and the sparse finalize launch of another group.
    static constexpr int MAX_GROUPS = 8;
    int n_groups = 1;
    cudaStream_t ext_stream[MAX_GROUPS] = {};   // [0] aliases `stream`
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_GROUPS] = {};
    std::vector<cudaEvent_t> ev_frame; // [n_groups][window_frames] recorded on the group's stream after the finalize of frame f (index f % F)
    // The propagation launches of consecutive frames are chained with programmatic dependent launch, which an event record
    // between two kernels would break, so a frame's table is announced in batches: ev_of_frame[f % F] is the frame whose
    // event (recorded after a later finalize of the same extract call) covers frame f.
    std::vector<int> ev_of_frame;      // [window_frames]
    int ev_batch = 4;                  // frames per event (MOVFE_EVENT_BATCH)
    bool cand_pipe = true;             // MOVFE_CAND_PIPE=0: candidate patches through registers instead of the cp.async window pipeline
    int  cand_bps = 0;                 // MOVFE_CAND_BPS (development): CTAs per stream of the candidate / birth kernels (0: enough to fill the chip)
    int  cand_pad_bytes = 0;           // MOVFE_CAND_PAD_KB (development): extra dynamic shared memory per cand_lane_kernel CTA, i.e. fewer of them per SM
    int  birth_chunks = 2;             // MOVFE_BIRTH_CHUNKS: 32-entry kps chunks a warp of birth_lane_kernel scans (more: fuller steps; fewer: more warps in flight)
    bool cand_lane = true;             // MOVFE_CAND_LANE=0: warp-level descriptor evaluation (cand_kernel / birth_kernel) instead of the thread-level kernels
    int pdl_mode = 0;                  // MOVFE_PDL: 0 off (default, see profiles/README.md), 1 every edge of the chain, 2 all but finalize -> next frame's cand
    struct PoseLaunch { int64_t first; int n; cudaEvent_t done; };
    PoseLaunch pose_launches[4] = {};  // ring of the last pose launches (events created at movfe_create)
    int     pose_launch_head = 0;
    int     TSLOTS = 0;                // track-table slots per stream: 2*window_frames + 1
    std::string err;

    int64_t pushed = 0;            // frames pushed per stream so far
    int64_t ext_first = -1;        // last extract window
    int     ext_n = 0;
    int64_t pose_first = -1;
    int     pose_n = 0;

    // staging for host pushes (grown on demand), double-buffered: the host->device copies of push k+1 run on
    // copy_stream while the kernels of window k run on `stream`; events order buffer reuse
    cudaStream_t copy_stream = nullptr;
    void   *d_stage[2] = {nullptr, nullptr};
    size_t  stage_bytes[2] = {0, 0};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};    // recorded on copy_stream after the copies into d_stage[b]
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};  // recorded on stream after the ingest kernels read d_stage[b]
    bool    stage_used[2] = {false, false};
    void   *h_meta[2] = {nullptr, nullptr};   // pinned copies of the offsets + flags of a push (callers pass stack arrays)
    size_t  h_meta_bytes[2] = {0, 0};
    int     push_parity = 0;
    bool    grey_direct = false;   // MOVFE_GREY_DIRECT=1: tightly packed host luma planes also go into the ring by strided copies (planes with a row
                                   // stride always do). Measured 5 % slower end to end than one flat copy + grey_ingest_kernel: 134 k vs 142 k frames/s
    int     grey_pitch = 0;        // row pitch of the grey ring: power of two >= width (compile-time strides in extract.cu)

    // record / image ring, slot = absolute frame % RING
    Rec16   *d_rec = nullptr;      // [S][RING][max_records]
    int32_t *d_rec_cnt = nullptr;  // [S][RING]
    uint8_t *d_fflags = nullptr;   // [S][RING]
    uint8_t *d_grey = nullptr;     // [S][RING][H*grey_pitch]   (has_grey)
    unsigned long long *d_rejected = nullptr;

    // raster results: two buffers, so that the raster of window k+1 (on raster_stream) runs beside the propagation of
    // window k (on stream); rb_cur is the buffer of the last movfe_raster call
    RasterBuf rb[2];
    int rb_cur = 0;
    cudaStream_t raster_stream = nullptr;  // ingest kernels + raster kernels (low priority: they fill what propagation leaves)
    // MOVFE_INGEST_STREAM=1: the ingest kernels of a push on a stream of their own (low priority), so that the ingest of window
    // k+2 runs beside the hop lists / slot resolution of window k+1 instead of queueing behind them; otherwise an alias of
    // raster_stream. ev_ingested orders the raster after the pushes it reads.
    cudaStream_t ingest_stream = nullptr;
    bool ingest_split = false;
    cudaEvent_t ev_ingested = nullptr;
    // (default; MOVFE_HOPS_PRIO=0 turns it off) the five short hop-list kernels of a raster (count .. bbox, ~0.12 ms alone) on a HIGH-priority stream between
    // two events, the long slot-resolution kernel stays on the low-priority raster stream (a chain of short launches at low priority
    // waits for a free SM five times over)
    cudaStream_t hops_stream = nullptr;
    cudaEvent_t ev_hops = nullptr;
    bool serial_raster = false;            // MOVFE_CFG_SERIAL_RASTER: raster waits for all earlier propagation (timing a kernel alone)
    cudaEvent_t ev_serial = nullptr;
    struct ExtLaunch { int64_t first; int n; cudaEvent_t done; };
    static constexpr int N_EXT_LAUNCHES = 8;
    ExtLaunch ext_launches[N_EXT_LAUNCHES] = {};  // ring of the last extract launches (ring-slot reuse by later pushes waits on them)
    int ext_launch_head = 0;               // next entry to overwrite == oldest entry once the ring has wrapped
    int64_t ext_launch_count = 0;

    // track tables: [S][TSLOTS][max_tracks], slot = frame % TSLOTS; two windows are resident so that the pose stream can
    // still read window k while propagation writes window k+1
    movfe_track *d_tracks = nullptr;
    int32_t *d_ntracks = nullptr;   // [S][TSLOTS]
    int32_t *d_cur_id = nullptr;    // [S][TSLOTS]  mCurrentId after each frame
    void    *d_ext_scratch = nullptr;
    size_t   ext_scratch_bytes = 0;
    bool     lk_pending = false;    // movfe_set_lk_results installed host LK results for the next frame to be propagated

    // map / pose
    movfe_camera cam;
    movfe_pose_params pp;
    float view_cos = 0.5f;
    movfe_map_point *d_map = nullptr;  // [S][max_map_points]
    int32_t *d_nmap = nullptr, *d_nkf = nullptr;  // [S]
    movfe_pose *d_pose_cur = nullptr;  // [S]
    movfe_pose *d_poses = nullptr;     // [S][F]
    int32_t *d_ninl = nullptr;         // [S][F]
    int32_t *d_match = nullptr;        // [S][F][max_tracks]
    uint8_t *d_outlier = nullptr;      // [S][F][max_tracks]
    // staging of movfe_set_map_points_batch (host hand-over of all streams' local maps), double-buffered so that a
    // hand-over never waits for the pose chain that consumed the previous one
    void    *d_map_stage[2] = {nullptr, nullptr};
    void    *h_map_meta[2] = {nullptr, nullptr};
    size_t   map_stage_bytes[2] = {0, 0};
    cudaEvent_t ev_map_staged[2] = {nullptr, nullptr};   // recorded on the pose stream after the install kernel read buffer b
    int      map_parity = 0;
    // device-resident map-point store + scratch of movfe_update_local_points (allocated by movfe_reserve_map_store)
    movfe_map_point *d_store = nullptr;    // [S][store_cap]
    int32_t *d_store_stamp = nullptr;      // [S][store_cap] first list position of a point (0x7fffffff between calls)
    int      store_cap = 0;
    int32_t *d_lp_idx = nullptr;           // staged index lists
    int64_t *d_lp_off = nullptr;           // [S + 1] offsets, then [S] keyframe entry counts as int32
    size_t   lp_idx_cap = 0;
    void    *h_lp_meta = nullptr;          // pinned copy of offsets + counts
    void    *d_lk_scratch = nullptr;   // pyramids + derivatives of two frames of every stream (movfe_lk_carry), allocated on first use
    size_t   lk_scratch_bytes = 0;
    void    *d_pose_scratch = nullptr;
    size_t   pose_scratch_bytes = 0;
    // split pose chain (join kernels + small solver kernels, pose.cu): correspondences of one frame per stream
    float   *d_pairs = nullptr;        // [S][6][max_map_points]: x y z u v idx
    int32_t *d_npairs = nullptr;       // [S]
    int      h_nmap_max = 0;           // largest local map installed so far (sizes the solver CTAs)
    bool     pose_v1 = false;          // MOVFE_POSE_V1=1: the first form of the fused pose chain (one launch per frame, joins over the whole track table)
    int      pose_group = 4;           // MOVFE_POSE_GROUP: frames per launch pair of the second form
    bool     pose_split = false;       // MOVFE_POSE_SPLIT=1: join kernels + small solver kernels instead of the fused
                                       // one-kernel-per-frame chain (measured slower under load)

    // workload counters (movfe_workload_stats): diagnostic integer atomics, never on the data path
    //  [0] tracks looked up   [1] candidate hops of those tracks   [2] pose solves   [3] correspondences of those solves
    //  [4] passes over the correspondences (GN iterations + re-classifications)   [5] hops of rastered frames   [6] frames rastered
    unsigned long long *d_stats = nullptr;

    // instrumentation
    bool prof_on = false;
    struct ProfSpan { int stage; cudaEvent_t a, b; };
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_free;
    double  prof_ms[MOVFE_N_STAGES] = {0};
    int64_t prof_launches[MOVFE_N_STAGES] = {0};

    // scratch for the single-shot operators (grown on demand)
    void   *d_op = nullptr;
    size_t  op_bytes = 0;
};

#define MOVFE_FAIL(ctx, code, ...)                         \
    do {                                                   \
        char _b[512];                                      \
        snprintf(_b, sizeof _b, __VA_ARGS__);              \
        (ctx)->err = _b;                                   \
        return (code);                                     \
    } while (0)

#define MOVFE_CUDA(ctx, expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) MOVFE_FAIL(ctx, MOVFE_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

__device__ __forceinline__ int4 resolve_slots(const TileCells &q, int x, int y) {
    const int tile = (y >> 5) * q.NT + (x >> 5);
    const int dim = __ldg(&q.dim[tile]);
    if (dim == 0) return make_int4(-1, -1, -1, -1);
    const uint8_t *r = q.runs + (size_t)tile * 64;
    const int cr = __ldg(&r[x & 31]), rr = __ldg(&r[32 + (y & 31)]);
    return __ldg(&q.cells[(size_t)tile * MOVFE_TILE_CELLS + rr * (dim & 0xff) + cr]);
}

// 128-bit streaming store: the slot grid is written once and not re-read by the writer.
__device__ __forceinline__ void st_cs_v4(int4 *p, int4 v) {
    asm volatile("st.global.cs.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled while the previous kernel
// of its stream is still running; it must execute pdl_wait() before it touches anything that kernel wrote. pdl_trigger()
// lets the NEXT kernel of the stream be scheduled early in the same way. Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Opts a kernel in to the largest dynamic shared memory the device allows beside the kernel's static allocation. Called
// once per kernel at movfe_create (the attribute is per function and process-wide; the value is the same from every context).
template <typename K>
static inline cudaError_t optin_dynamic_smem(K kernel, int smem_optin, int *dynamic_limit = nullptr) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    const int lim = smem_optin - (int)fa.sharedSizeBytes;
    if (dynamic_limit) *dynamic_limit = lim;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
}

// RAII span: records an event pair around a stage when profiling is on, and always counts kernel launches.
struct ProfScope {
    movfe_ctx *ctx;
    int stage;
    cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(movfe_ctx *c) {
        cudaEvent_t e;
        if (!c->prof_free.empty()) { e = c->prof_free.back(); c->prof_free.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
    cudaStream_t on;
    ProfScope(movfe_ctx *c, int st, cudaStream_t s = nullptr) : ctx(c), stage(st), on(s ? s : c->stream) {
        if (ctx->prof_on) { a = get(ctx); b = get(ctx); cudaEventRecord(a, on); }
    }
    void launches(int n) { ctx->prof_launches[stage] += n; }
    ~ProfScope() {
        if (a) { cudaEventRecord(b, on); ctx->prof_spans.push_back({stage, a, b}); }
    }
};

// Parameters of one raster window (frames [first, first+n_in) of every stream; the first n_out are outputs).
struct WinParams {
    int S, n_in, n_out, K, RING, maxM, W, H;
    int64_t first;
    int max_hops, max_kps, max_chunks;
};

// grid.cu
int movfe_grid_launch(movfe_ctx *ctx, const WinParams &p, RasterBuf &w);
// raster.cu
int movfe_raster_launch(movfe_ctx *ctx, RasterBuf &w, int64_t first_frame, int n_out, int n_in);
int movfe_ingest_launch(movfe_ctx *ctx, int n_frames, const void *d_recs, bool packed, const int64_t *d_rec_off,
                        int64_t n_records, const uint8_t *d_flags, const uint8_t *d_grey);
// extract.cu
struct movfe_lk_handover {
    int32_t *n;         // [S] results installed for the next propagated frame (-1: none)
    uint8_t *status;    // [S][max_tracks]
    float   *pts;       // [S][max_tracks][2]
    uint16_t *order;    // [S][max_tracks] sorted rank -> index in the newest table
};
void movfe_lk_buffers(const movfe_ctx *ctx, movfe_lk_handover *out);
int movfe_extract_launch(movfe_ctx *ctx, int64_t first_frame, int n_frames);
size_t movfe_extract_scratch_bytes(const movfe_ctx *ctx);
int movfe_extract_init(movfe_ctx *ctx);
int movfe_pose_init(movfe_ctx *ctx);    // pose.cu
int movfe_bucket_init(movfe_ctx *ctx);  // bucket.cu
// match.cu / pose.cu
int movfe_track_poses_launch(movfe_ctx *ctx, int64_t first_frame, int n_frames);
size_t movfe_pose_scratch_bytes(const movfe_ctx *ctx);
