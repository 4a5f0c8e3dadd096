// api.cu — the C-ABI of include/movfe.h: context lifetime, ingest, raster and downloads.
// There is no CPU fallback anywhere in this library: without a CUDA device movfe_create fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

static thread_local std::string g_create_error;  // per host thread: contexts are created from their own threads

extern "C" const char *movfe_version(void) { return "movfe 0.1 (sm_100a)"; }

extern "C" const char *movfe_last_error(const movfe_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

template <typename T>
static cudaError_t dalloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    return cudaMalloc((void **)p, n * sizeof(T));
}

extern "C" void movfe_destroy(movfe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->pose_stream) cudaStreamSynchronize(ctx->pose_stream);
    if (ctx->ingest_split && ctx->ingest_stream) cudaStreamSynchronize(ctx->ingest_stream);
    if (ctx->raster_stream) cudaStreamSynchronize(ctx->raster_stream);
    for (int g = 1; g < movfe_ctx::MAX_GROUPS; g++)
        if (ctx->ext_stream[g]) cudaStreamSynchronize(ctx->ext_stream[g]);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void *bufs[] = {ctx->d_stage[0], ctx->d_stage[1], ctx->d_rec, ctx->d_rec_cnt, ctx->d_fflags, ctx->d_grey, ctx->d_rejected, ctx->d_stats,
                    ctx->d_tracks, ctx->d_ntracks,
                    ctx->d_cur_id, ctx->d_ext_scratch, ctx->d_map, ctx->d_nmap, ctx->d_nkf, ctx->d_pose_cur,
                    ctx->d_poses, ctx->d_ninl, ctx->d_match, ctx->d_outlier, ctx->d_pose_scratch, ctx->d_op, ctx->d_pairs, ctx->d_npairs};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (ctx->d_lk_scratch) cudaFree(ctx->d_lk_scratch);
    if (ctx->d_store) cudaFree(ctx->d_store);
    if (ctx->d_store_stamp) cudaFree(ctx->d_store_stamp);
    if (ctx->d_lp_idx) cudaFree(ctx->d_lp_idx);
    if (ctx->d_lp_off) cudaFree(ctx->d_lp_off);
    if (ctx->h_lp_meta) cudaFreeHost(ctx->h_lp_meta);
    for (int b = 0; b < 2; b++) {
        if (ctx->d_map_stage[b]) cudaFree(ctx->d_map_stage[b]);
        if (ctx->h_map_meta[b]) cudaFreeHost(ctx->h_map_meta[b]);
        if (ctx->ev_map_staged[b]) cudaEventDestroy(ctx->ev_map_staged[b]);
    }
    for (RasterBuf &w : ctx->rb) {
        void *wb[] = {w.d_seg_cnt, w.d_cls_cnt, w.d_area, w.d_hop_base, w.d_kps_base, w.d_nhops, w.d_nkps, w.d_cov, w.d_hops, w.d_hop_rect, w.d_kps,
                      w.d_chunk_bbox, w.d_grid, w.d_tc_dim, w.d_tc_runs, w.d_tc_cells};
        for (void *b : wb)
            if (b) cudaFree(b);
        if (w.done) cudaEventDestroy(w.done);
        if (w.consumed) cudaEventDestroy(w.consumed);
    }
    for (auto &el : ctx->ext_launches)
        if (el.done) cudaEventDestroy(el.done);
    if (ctx->ev_serial) cudaEventDestroy(ctx->ev_serial);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int g = 1; g < movfe_ctx::MAX_GROUPS; g++) {
        if (ctx->ev_join[g]) cudaEventDestroy(ctx->ev_join[g]);
        if (ctx->ext_stream[g]) cudaStreamDestroy(ctx->ext_stream[g]);
    }
    for (auto &sp : ctx->prof_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    for (int b = 0; b < 2; b++) {
        if (ctx->h_meta[b]) cudaFreeHost(ctx->h_meta[b]);
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
        if (ctx->ev_consumed[b]) cudaEventDestroy(ctx->ev_consumed[b]);
    }
    if (ctx->ev_tables) cudaEventDestroy(ctx->ev_tables);
    for (auto e : ctx->ev_frame)
        if (e) cudaEventDestroy(e);
    for (auto &pl : ctx->pose_launches)
        if (pl.done) cudaEventDestroy(pl.done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->pose_stream) cudaStreamDestroy(ctx->pose_stream);
    if (ctx->ingest_split && ctx->ingest_stream) cudaStreamDestroy(ctx->ingest_stream);
    if (ctx->ev_ingested) cudaEventDestroy(ctx->ev_ingested);
    if (ctx->hops_stream) cudaStreamDestroy(ctx->hops_stream);
    if (ctx->ev_hops) cudaEventDestroy(ctx->ev_hops);
    if (ctx->raster_stream) cudaStreamDestroy(ctx->raster_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int movfe_create(const movfe_config *cfg, movfe_ctx **out) {
    if (!cfg || !out) {
        g_create_error = "movfe_create: null argument";
        return MOVFE_E_INVALID;
    }
    *out = nullptr;
    movfe_ctx *ctx = new (std::nothrow) movfe_ctx;
    if (!ctx) {
        g_create_error = "movfe_create: out of host memory";
        return MOVFE_E_INVALID;
    }
    ctx->cfg = *cfg;
    auto fail = [&](int code) {
        g_create_error = ctx->err;
        movfe_destroy(ctx);
        return code;
    };
    const movfe_config &c = ctx->cfg;
    if (c.n_streams < 1 || c.width < 16 || c.height < 16 || c.width > 16384 || c.height > 16384 ||
        c.max_records_per_frame < 1 || c.max_ref < 0 || c.max_ref > MOVFE_MAX_K || c.window_frames < 1 ||
        c.max_tracks < 1 || c.max_tracks > 65536 || c.max_map_points < 0) {
        ctx->err = "movfe_create: configuration out of range";
        return fail(MOVFE_E_INVALID);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        ctx->err = std::string("movfe_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path";
        return fail(MOVFE_E_CUDA);
    }
#define CK(x)                                                                     \
    do {                                                                          \
        cudaError_t _e = (x);                                                     \
        if (_e != cudaSuccess) {                                                  \
            ctx->err = std::string(#x) + ": " + cudaGetErrorString(_e);           \
            return fail(MOVFE_E_CUDA);                                            \
        }                                                                         \
    } while (0)
    CK(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c.device));
    ctx->sm_count = prop.multiProcessorCount;
    CK(cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device));
    // propagation is a serial chain of short launches (three per frame): it gets the high priority, so that its CTAs are
    // placed as soon as they are ready and the long, throughput-bound raster kernels of the NEXT window fill what is left
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    ctx->serial_raster = (c.flags & MOVFE_CFG_SERIAL_RASTER) != 0;
    ctx->fused = (c.flags & MOVFE_CFG_NO_GRID) != 0;
    // pose chain and propagation share the high priority (a pose stream ABOVE propagation measured slower: 4.86-5.05 ms per
    // step against 4.75 ms)
    const int prio_prop = prio_hi;
    CK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_prop));
    CK(cudaStreamCreateWithPriority(&ctx->pose_stream, cudaStreamNonBlocking, prio_hi));
    // MOVFE_RASTER_PRIO=1 (development): the raster stream at the propagation priority
    CK(cudaStreamCreateWithPriority(&ctx->raster_stream, cudaStreamNonBlocking, (getenv("MOVFE_RASTER_PRIO") && atoi(getenv("MOVFE_RASTER_PRIO"))) ? prio_hi : prio_lo));
    // MOVFE_COPY_PRIO=1 (development): the host->device copies on a high-priority stream
    CK(cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, (getenv("MOVFE_COPY_PRIO") && atoi(getenv("MOVFE_COPY_PRIO"))) ? prio_hi : prio_lo));
    ctx->ingest_split = getenv("MOVFE_INGEST_STREAM") && atoi(getenv("MOVFE_INGEST_STREAM"));
    if (ctx->ingest_split) CK(cudaStreamCreateWithPriority(&ctx->ingest_stream, cudaStreamNonBlocking, atoi(getenv("MOVFE_INGEST_STREAM")) >= 2 ? prio_hi : prio_lo));
    else ctx->ingest_stream = ctx->raster_stream;
    CK(cudaEventCreateWithFlags(&ctx->ev_ingested, cudaEventDisableTiming));
    if (!getenv("MOVFE_HOPS_PRIO") || atoi(getenv("MOVFE_HOPS_PRIO"))) {  // default on: 3.26 against 3.30 ms per C2 step
        CK(cudaStreamCreateWithPriority(&ctx->hops_stream, cudaStreamNonBlocking, prio_hi));
        CK(cudaEventCreateWithFlags(&ctx->ev_hops, cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_tables, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_serial, cudaEventDisableTiming));
    for (RasterBuf &w : ctx->rb) {
        CK(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&w.consumed, cudaEventDisableTiming));
    }
    for (auto &el : ctx->ext_launches) {
        el.first = -1;
        el.n = 0;
        CK(cudaEventCreateWithFlags(&el.done, cudaEventDisableTiming));
    }
    {
        // groups of independent propagation chains (MOVFE_EXTRACT_GROUPS). With the thread-level kernels a chain is bound by
        // latency, not by issue slots, and its per-stream finalize launch leaves most SMs idle: three chains measured 3-4 %
        // faster than one at 64 streams (1 / 2 / 3 / 4 / 8 groups: 3.40 / 3.36 / 3.27 / 3.33 / 3.62 ms per step - the host's
        // launch rate takes over beyond four). Contexts with few streams keep one chain.
        int g = c.n_streams >= 24 ? 3 : 1;
        if (const char *e = getenv("MOVFE_EXTRACT_GROUPS")) g = atoi(e);
        ctx->n_groups = std::max(1, std::min(std::min(g, c.n_streams), (int)movfe_ctx::MAX_GROUPS));
    }
    ctx->ext_stream[0] = ctx->stream;
    for (int g = 1; g < ctx->n_groups; g++) {
        CK(cudaStreamCreateWithPriority(&ctx->ext_stream[g], cudaStreamNonBlocking, prio_prop));
        CK(cudaEventCreateWithFlags(&ctx->ev_join[g], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    ctx->ev_of_frame.assign(c.window_frames, 0);
    if (const char *e = getenv("MOVFE_EVENT_BATCH")) ctx->ev_batch = std::max(1, atoi(e));
    if (const char *e = getenv("MOVFE_PDL")) ctx->pdl_mode = atoi(e);
    if (const char *e = getenv("MOVFE_CAND_PIPE")) ctx->cand_pipe = atoi(e) != 0;
    if (const char *e = getenv("MOVFE_CAND_LANE")) ctx->cand_lane = atoi(e) != 0;
    if (const char *e = getenv("MOVFE_GREY_DIRECT")) ctx->grey_direct = atoi(e) != 0;
    if (const char *e = getenv("MOVFE_BIRTH_CHUNKS")) ctx->birth_chunks = std::max(1, atoi(e));
    if (const char *e = getenv("MOVFE_CAND_PAD_KB")) ctx->cand_pad_bytes = std::max(0, atoi(e)) * 1024;
    if (const char *e = getenv("MOVFE_CAND_BPS")) ctx->cand_bps = std::max(0, atoi(e));
    ctx->ev_frame.resize((size_t)ctx->n_groups * c.window_frames, nullptr);
    for (auto &e : ctx->ev_frame) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &pl : ctx->pose_launches) {
        pl.first = -1;
        pl.n = 0;
        CK(cudaEventCreateWithFlags(&pl.done, cudaEventDisableTiming));
    }
    for (int b = 0; b < 2; b++) {
        CK(cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_consumed[b], cudaEventDisableTiming));
    }
    ctx->grey_pitch = 1024;
    while (ctx->grey_pitch < c.width) ctx->grey_pitch <<= 1;

    ctx->K = c.max_ref;
    ctx->LA = ctx->K + 1;
    ctx->NIN = c.window_frames + ctx->LA;
    // two windows deep: the push + raster of window k+1 overwrite the slots of window k-1 while propagation still reads
    // the grey planes of window k
    ctx->RING = 2 * c.window_frames + ctx->LA;
    if (getenv("MOVFE_RING_EXTRA")) ctx->RING += std::max(0, atoi(getenv("MOVFE_RING_EXTRA"))) * c.window_frames;  // development: deeper ring
    ctx->NB = (c.height + 7) / 8;
    ctx->NT = (c.width + 31) / 32;
    ctx->NTR = (c.height + 31) / 32;
    ctx->max_hops = c.max_records_per_frame * (ctx->K + 1);
    ctx->max_kps = c.max_records_per_frame * (ctx->K + 1);
    ctx->max_chunks = (ctx->max_hops + 31) / 32;
    // count / emit: one CTA per frame up to 8192 records, segments of 4096 records beyond (dense 4x4 fields)
    ctx->rseg = c.max_records_per_frame <= 8192 ? ((c.max_records_per_frame + 511) / 512) * 512 : 4096;
    ctx->n_rseg = (c.max_records_per_frame + ctx->rseg - 1) / ctx->rseg;
    if (ctx->max_hops >= (1 << 22)) {
        ctx->err = "movfe_create: max_records_per_frame*(max_ref+1) must stay below 2^22";
        return fail(MOVFE_E_INVALID);
    }
    const size_t S = c.n_streams, F = c.window_frames, NIN = ctx->NIN, RING = ctx->RING;
    const size_t plane = (size_t)c.width * c.height;
    CK(dalloc(&ctx->d_rec, S * RING * c.max_records_per_frame));
    CK(dalloc(&ctx->d_rec_cnt, S * RING));
    CK(dalloc(&ctx->d_fflags, S * RING));
    // + 16 rows: the window pipeline of cand_kernel fetches 16 rows for every block shape (8-row blocks near the bottom of the
    // last plane read past it; the extra rows are never used)
    if (c.has_grey) CK(dalloc(&ctx->d_grey, S * RING * (size_t)c.height * ctx->grey_pitch + 16 * (size_t)ctx->grey_pitch));
    CK(dalloc(&ctx->d_stats, 8));
    CK(cudaMemset(ctx->d_stats, 0, 8 * sizeof(unsigned long long)));
    CK(dalloc(&ctx->d_rejected, 1));
    CK(cudaMemset(ctx->d_rejected, 0, sizeof(unsigned long long)));
    CK(cudaMemset(ctx->d_rec_cnt, 0, S * RING * sizeof(int32_t)));
    CK(cudaMemset(ctx->d_fflags, 0, S * RING));
    for (RasterBuf &w : ctx->rb) {
        CK(dalloc(&w.d_seg_cnt, S * NIN * ctx->n_rseg * (MOVFE_NCLS(MOVFE_MAX_K) + 2)));
        CK(dalloc(&w.d_cls_cnt, S * NIN * MOVFE_NCLS(MOVFE_MAX_K)));
        CK(dalloc(&w.d_area, S * NIN));
        CK(dalloc(&w.d_hop_base, S * NIN * (ctx->K + 2)));
        CK(dalloc(&w.d_kps_base, S * NIN * (ctx->K + 2)));
        CK(dalloc(&w.d_nhops, S * NIN));
        CK(dalloc(&w.d_nkps, S * NIN));
        CK(dalloc(&w.d_cov, S * NIN));
        CK(dalloc(&w.d_hops, S * F * ctx->max_hops));
        CK(dalloc(&w.d_hop_rect, S * F * ctx->max_hops));
        CK(dalloc(&w.d_kps, S * F * ctx->max_kps));
        CK(dalloc(&w.d_chunk_bbox, S * F * ctx->max_chunks));
        if (ctx->fused) {
            const size_t tiles = (size_t)ctx->NT * ctx->NTR;
            CK(dalloc(&w.d_tc_dim, S * F * tiles));
            CK(dalloc(&w.d_tc_runs, S * F * tiles * 64));
            CK(dalloc(&w.d_tc_cells, S * F * tiles * MOVFE_TILE_CELLS));
        } else {
            CK(dalloc(&w.d_grid, S * F * plane));
        }
    }
    // track tables
    ctx->TSLOTS = 2 * c.window_frames + 1;
    const size_t TS = ctx->TSLOTS;
    CK(dalloc(&ctx->d_tracks, S * TS * c.max_tracks));
    CK(dalloc(&ctx->d_ntracks, S * TS));
    CK(dalloc(&ctx->d_cur_id, S * TS));
    CK(cudaMemset(ctx->d_ntracks, 0, S * TS * sizeof(int32_t)));
    CK(cudaMemset(ctx->d_cur_id, 0, S * TS * sizeof(int32_t)));
    ctx->ext_scratch_bytes = movfe_extract_scratch_bytes(ctx);
    CK(cudaMalloc(&ctx->d_ext_scratch, std::max<size_t>(ctx->ext_scratch_bytes, 16)));
    if (int rc = movfe_extract_init(ctx)) return fail(rc);
    if (int rc = movfe_pose_init(ctx)) return fail(rc);
    if (int rc = movfe_bucket_init(ctx)) return fail(rc);
    // map / pose
    CK(dalloc(&ctx->d_map, S * (size_t)std::max(c.max_map_points, 1)));
    CK(dalloc(&ctx->d_nmap, S));
    CK(dalloc(&ctx->d_nkf, S));
    CK(cudaMemset(ctx->d_nmap, 0, S * sizeof(int32_t)));
    CK(cudaMemset(ctx->d_nkf, 0, S * sizeof(int32_t)));
    CK(dalloc(&ctx->d_pose_cur, S));
    CK(dalloc(&ctx->d_poses, S * F));
    CK(dalloc(&ctx->d_ninl, S * F));
    CK(dalloc(&ctx->d_match, S * F * c.max_tracks));
    CK(dalloc(&ctx->d_outlier, S * F * c.max_tracks));
    {
        std::vector<movfe_pose> id(S);
        for (auto &p : id) {
            memset(&p, 0, sizeof p);
            p.R[0] = p.R[4] = p.R[8] = 1.0;
        }
        CK(cudaMemcpy(ctx->d_pose_cur, id.data(), S * sizeof(movfe_pose), cudaMemcpyHostToDevice));
    }
    CK(dalloc(&ctx->d_pairs, S * 6 * (size_t)std::max(c.max_map_points, 1)));
    CK(dalloc(&ctx->d_npairs, S));
    CK(cudaMemset(ctx->d_npairs, 0, S * sizeof(int32_t)));
    if (const char *e = getenv("MOVFE_POSE_SPLIT")) ctx->pose_split = atoi(e) != 0;
    if (const char *e = getenv("MOVFE_POSE_V1")) ctx->pose_v1 = atoi(e) != 0;
    if (const char *e = getenv("MOVFE_POSE_GROUP")) ctx->pose_group = std::max(1, atoi(e));
    ctx->pose_scratch_bytes = movfe_pose_scratch_bytes(ctx);
    CK(cudaMalloc(&ctx->d_pose_scratch, std::max<size_t>(ctx->pose_scratch_bytes, 16)));
    CK(cudaMemset(ctx->d_pose_scratch, 0, std::max<size_t>(ctx->pose_scratch_bytes, 16)));
    memset(&ctx->cam, 0, sizeof ctx->cam);
    ctx->cam.fx = ctx->cam.fy = 1.f;
    memset(&ctx->pp, 0, sizeof ctx->pp);
    ctx->pp.iteration_count = 50;
    ctx->pp.reprojection_error = 5.0;
    ctx->pp.reprojection_error_lost = 8.0;
    ctx->pp.confidence = 0.95;
    ctx->pp.algorithm = 38;
    CK(cudaStreamSynchronize(ctx->stream));
#undef CK
    *out = ctx;
    return MOVFE_OK;
}

extern "C" int movfe_synchronize(movfe_ctx *ctx) {
    if (!ctx) return MOVFE_E_INVALID;
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->ingest_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->raster_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    return MOVFE_OK;
}

extern "C" int movfe_fence(movfe_ctx *ctx) {
    if (!ctx) return MOVFE_E_INVALID;
    // device-side join: everything enqueued so far on the pose stream precedes whatever the caller enqueues next on the
    // primary stream (an event record for timing, a dependent kernel)
    MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_tables, ctx->pose_stream));
    MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_tables, 0));
    MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_tables, ctx->raster_stream));
    MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_tables, 0));
    if (ctx->ingest_split) {
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_tables, ctx->ingest_stream));
        MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_tables, 0));
    }
    return MOVFE_OK;
}

extern "C" void *movfe_cuda_stream(movfe_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t movfe_frames_pushed(const movfe_ctx *ctx) { return ctx ? ctx->pushed : -1; }

static int ensure_stage(movfe_ctx *ctx, int b, size_t bytes) {
    if (bytes <= ctx->stage_bytes[b]) return MOVFE_OK;
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->ingest_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->raster_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_stage[b]) cudaFree(ctx->d_stage[b]);
    ctx->d_stage[b] = nullptr;
    ctx->stage_bytes[b] = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_stage[b], want));
    ctx->stage_bytes[b] = want;
    return MOVFE_OK;
}

static int ensure_op(movfe_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->op_bytes) return MOVFE_OK;
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_op) cudaFree(ctx->d_op);
    ctx->d_op = nullptr;
    ctx->op_bytes = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    MOVFE_CUDA(ctx, cudaMalloc(&ctx->d_op, want));
    ctx->op_bytes = want;
    return MOVFE_OK;
}
int movfe_ensure_op_scratch(movfe_ctx *ctx, size_t bytes) { return ensure_op(ctx, bytes); }

// The ingest kernels of a push overwrite the ring slots of frames [pushed-RING, pushed+n-RING): propagation launches (on
// the primary stream) that read the grey planes / frame flags of those frames must have finished. Launches are ordered on
// one stream, so waiting for the newest launch that starts at or before the last overwritten frame covers all of them;
// when that launch has already left the bookkeeping ring, its oldest entry (a later launch) stands in for it.
static int wait_ring_readers(movfe_ctx *ctx, int n_frames, cudaStream_t waiter = nullptr) {
    if (!waiter) waiter = ctx->ingest_stream;
    const int64_t last_overwritten = ctx->pushed + n_frames - 1 - ctx->RING;
    if (last_overwritten < 0) return MOVFE_OK;
    if (waiter != ctx->raster_stream) {
        // the raster launches read the record slots as well: on their own stream nothing orders them implicitly any more. A raster
        // whose input frames meet the overwritten ones must have finished (an older raster precedes these on the raster stream).
        const int64_t first_overwritten = ctx->pushed - ctx->RING;
        for (const RasterBuf &w : ctx->rb)
            if (w.first >= 0 && w.first <= last_overwritten && w.first + w.nin > first_overwritten)
                MOVFE_CUDA(ctx, cudaStreamWaitEvent(waiter, w.done, 0));
    }
    if (ctx->ext_launch_count == 0) return MOVFE_OK;
    const int N = movfe_ctx::N_EXT_LAUNCHES;
    const int live = (int)std::min<int64_t>(ctx->ext_launch_count, N);
    const movfe_ctx::ExtLaunch *pick = nullptr;
    for (int i = 0; i < live; i++) {  // newest first
        const movfe_ctx::ExtLaunch &el = ctx->ext_launches[((ctx->ext_launch_head - 1 - i) % N + N) % N];
        if (el.first <= last_overwritten) {
            pick = &el;
            break;
        }
    }
    if (!pick && ctx->ext_launch_count > N) pick = &ctx->ext_launches[ctx->ext_launch_head];  // oldest entry
    if (pick) MOVFE_CUDA(ctx, cudaStreamWaitEvent(waiter, pick->done, 0));
    return MOVFE_OK;
}

// whatever reads the ring is ordered behind the raster stream (the raster itself, then propagation through RasterBuf::done; the
// LK carry-over through its own wait): with the ingest kernels on a stream of their own the raster stream waits for them here
static int after_ingest(movfe_ctx *ctx) {
    if (!ctx->ingest_split) return MOVFE_OK;
    MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_ingested, ctx->ingest_stream));
    MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->raster_stream, ctx->ev_ingested, 0));
    return MOVFE_OK;
}

static int check_push(movfe_ctx *ctx, int n_frames) {
    if (!ctx) return MOVFE_E_INVALID;
    if (n_frames < 1 || n_frames > ctx->RING)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "push of %d frames exceeds the ring depth %d (2*window_frames + max_ref + 1)", n_frames, ctx->RING);
    return MOVFE_OK;
}

extern "C" int movfe_push_frames_device(movfe_ctx *ctx, int n_frames, const movfe_mv_record *d_recs,
                                        const int64_t *d_rec_off, int64_t n_records, const uint8_t *d_frame_flags,
                                        const uint8_t *d_grey) {
    int rc = check_push(ctx, n_frames);
    if (rc) return rc;
    if (!d_rec_off || !d_frame_flags || (n_records > 0 && !d_recs)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "push: null pointer");
    if (((uintptr_t)d_recs & 15) != 0) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "push: device record array must be 16-byte aligned");
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    rc = wait_ring_readers(ctx, n_frames);
    if (rc) return rc;
    rc = movfe_ingest_launch(ctx, n_frames, d_recs, false, d_rec_off, n_records, d_frame_flags, d_grey);
    if (rc) return rc;
    rc = after_ingest(ctx);
    if (rc) return rc;
    ctx->pushed += n_frames;
    return MOVFE_OK;
}

static int push_host(movfe_ctx *ctx, int n_frames, const void *recs, size_t rec_size, const int64_t *rec_off,
                     const uint8_t *frame_flags, const uint8_t *grey, int grey_stride) {
    int rc = check_push(ctx, n_frames);
    if (rc) return rc;
    if (!rec_off || !frame_flags) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "push: null pointer");
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const size_t n_seg = (size_t)ctx->cfg.n_streams * n_frames;
    const int64_t n_records = rec_off[n_seg];
    if (n_records < 0 || (n_records > 0 && !recs)) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "push: bad record offsets");
    // staging layout: [records, padded to 16 B][offsets][flags, padded to 16 B][grey planes]
    const size_t rec_bytes = ((size_t)n_records * rec_size + 15) & ~(size_t)15;
    const size_t off_bytes = ((n_seg + 1) * sizeof(int64_t) + 15) & ~(size_t)15;
    const size_t flag_bytes = (n_seg + 15) & ~(size_t)15;
    const bool with_grey = ctx->cfg.has_grey && grey;
    if (grey_stride == 0) grey_stride = ctx->cfg.width;
    if (with_grey && grey_stride < ctx->cfg.width) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "push: grey_stride %d below the frame width %d", grey_stride, ctx->cfg.width);
    // luma planes: straight into the pitched ring with one strided copy per stream (AVFrame::linesize / cv::Mat::step rows are
    // taken as they are), or - the planes tightly packed, unless MOVFE_GREY_DIRECT=1 - through the staging buffer + grey_ingest_kernel
    // (one flat copy moves faster over PCIe than 64 strided ones: common.cuh)
    const bool direct = with_grey && (ctx->grey_direct || grey_stride != ctx->cfg.width);
    const size_t grey_bytes = (with_grey && !direct) ? n_seg * (size_t)ctx->cfg.width * ctx->cfg.height : 0;
    const size_t total = rec_bytes + off_bytes + flag_bytes + grey_bytes + 16;
    const int b = ctx->push_parity;
    rc = ensure_stage(ctx, b, total);
    if (rc) return rc;
    // the copies run on copy_stream so that they overlap the kernels of the previous window; the buffer must have been
    // consumed by the ingest kernels of the push two calls ago
    if (ctx->stage_used[b]) MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[b], 0));
    uint8_t *base = (uint8_t *)ctx->d_stage[b];
    cudaStream_t cs = ctx->copy_stream;
    if (n_records > 0)
        MOVFE_CUDA(ctx, cudaMemcpyAsync(base, recs, (size_t)n_records * rec_size, cudaMemcpyHostToDevice, cs));
    // offsets and flags are small and usually live on the caller's stack: they go through a pinned copy owned by the
    // context (reused only after ev_consumed[b], like the device buffer). Records and grey planes are the caller's to keep.
    if (ctx->h_meta_bytes[b] < off_bytes + flag_bytes) {
        if (ctx->h_meta[b]) {
            MOVFE_CUDA(ctx, cudaStreamSynchronize(cs));
            cudaFreeHost(ctx->h_meta[b]);
            ctx->h_meta[b] = nullptr;
            ctx->h_meta_bytes[b] = 0;
        }
        MOVFE_CUDA(ctx, cudaMallocHost(&ctx->h_meta[b], 2 * (off_bytes + flag_bytes)));
        ctx->h_meta_bytes[b] = 2 * (off_bytes + flag_bytes);
    } else if (ctx->stage_used[b]) {
        MOVFE_CUDA(ctx, cudaEventSynchronize(ctx->ev_consumed[b]));  // the previous copy out of h_meta[b] has completed
    }
    memcpy(ctx->h_meta[b], rec_off, (n_seg + 1) * sizeof(int64_t));
    memcpy((uint8_t *)ctx->h_meta[b] + off_bytes, frame_flags, n_seg);
    MOVFE_CUDA(ctx, cudaMemcpyAsync(base + rec_bytes, ctx->h_meta[b], off_bytes + flag_bytes, cudaMemcpyHostToDevice, cs));
    uint8_t *d_grey = nullptr;
    if (with_grey && !direct) {
        d_grey = base + rec_bytes + off_bytes + flag_bytes;
        MOVFE_CUDA(ctx, cudaMemcpyAsync(d_grey, grey, grey_bytes, cudaMemcpyHostToDevice, cs));
    } else if (with_grey) {
        // the ring slots of these frames may still be read by the propagation of the window they held, or be written by the
        // ingest kernels of the previous push (staged planes): the copies wait for both, nothing else
        rc = wait_ring_readers(ctx, n_frames, cs);
        if (rc) return rc;
        if (ctx->stage_used[b ^ 1]) MOVFE_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_consumed[b ^ 1], 0));
        const int W = ctx->cfg.width, H = ctx->cfg.height, R = ctx->RING;
        const size_t P = (size_t)ctx->grey_pitch;
        for (int st = 0; st < ctx->cfg.n_streams; st++) {
            int f = 0;
            while (f < n_frames) {  // consecutive frames are consecutive slots until the ring wraps: at most two copies per stream
                const int slot = (int)((ctx->pushed + f) % R);
                const int cnt = std::min(n_frames - f, R - slot);
                MOVFE_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_grey + ((size_t)st * R + slot) * H * P, P,
                                                  grey + ((size_t)st * n_frames + f) * (size_t)grey_stride * H, (size_t)grey_stride, (size_t)W,
                                                  (size_t)cnt * H, cudaMemcpyHostToDevice, cs));
                f += cnt;
            }
        }
    }
    MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b], cs));
    MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->ingest_stream, ctx->ev_copied[b], 0));
    rc = wait_ring_readers(ctx, n_frames);
    if (rc) return rc;
    rc = movfe_ingest_launch(ctx, n_frames, base, rec_size == sizeof(movfe_packed_record), (const int64_t *)(base + rec_bytes), n_records,
                             base + rec_bytes + off_bytes, d_grey);
    if (rc) return rc;
    MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_consumed[b], ctx->ingest_stream));
    rc = after_ingest(ctx);
    if (rc) return rc;
    ctx->stage_used[b] = true;
    ctx->push_parity ^= 1;
    ctx->pushed += n_frames;
    return MOVFE_OK;
}

extern "C" int movfe_push_frames(movfe_ctx *ctx, int n_frames, const movfe_mv_record *recs, const int64_t *rec_off,
                                 const uint8_t *frame_flags, const uint8_t *grey) {
    return push_host(ctx, n_frames, recs, sizeof(movfe_mv_record), rec_off, frame_flags, grey, 0);
}

extern "C" int movfe_push_frames_packed(movfe_ctx *ctx, int n_frames, const movfe_packed_record *recs, const int64_t *rec_off,
                                        const uint8_t *frame_flags, const uint8_t *grey, int grey_stride) {
    return push_host(ctx, n_frames, recs, sizeof(movfe_packed_record), rec_off, frame_flags, grey, grey_stride);
}

// Host code: the 40-byte side-data record -> the 16 bytes the path reads (the same repacking ingest_kernel does on the device).
extern "C" void movfe_pack_records(const movfe_mv_record *recs, int64_t n_records, movfe_packed_record *out) {
    for (int64_t i = 0; i < n_records; i++) {
        const movfe_mv_record &r = recs[i];
        movfe_packed_record o;
        o.src_x = r.src_x;
        o.src_y = r.src_y;
        o.dst_x = r.dst_x;
        o.dst_y = r.dst_y;
        o.w = r.w;
        o.h = r.h;
        o.source_sign = r.source < 0 ? -1 : (r.source > 0 ? 1 : 0);
        o.reserved = 0;
        o.ref = r.ref;
        out[i] = o;
    }
}

extern "C" int movfe_raster(movfe_ctx *ctx, int64_t first_frame, int n_out) {
    if (!ctx) return MOVFE_E_INVALID;
    if (n_out < 1 || n_out > ctx->cfg.window_frames)
        MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "raster: n_out=%d outside [1, window_frames=%d]", n_out, ctx->cfg.window_frames);
    if (first_frame < 0 || first_frame + n_out > ctx->pushed)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "raster: frames [%lld,%lld) were not pushed (pushed=%lld)", (long long)first_frame,
                   (long long)(first_frame + n_out), (long long)ctx->pushed);
    if (ctx->pushed - first_frame > ctx->RING)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "raster: frame %lld has left the ring (pushed=%lld, ring=%d)", (long long)first_frame,
                   (long long)ctx->pushed, ctx->RING);
    const int n_in = (int)std::min<int64_t>(ctx->pushed - first_frame, (int64_t)n_out + ctx->LA);
    MOVFE_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    // results go to the buffer that the propagation of the previous window is NOT reading
    RasterBuf &w = ctx->rb[ctx->rb_cur ^ 1];
    if (w.consumed_valid) MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->raster_stream, w.consumed, 0));
    if (ctx->serial_raster) {  // alone: after all propagation AND all pose work enqueued so far
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_serial, ctx->stream));
        MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->raster_stream, ctx->ev_serial, 0));
        MOVFE_CUDA(ctx, cudaEventRecord(ctx->ev_serial, ctx->pose_stream));
        MOVFE_CUDA(ctx, cudaStreamWaitEvent(ctx->raster_stream, ctx->ev_serial, 0));
    }
    int rc = movfe_raster_launch(ctx, w, first_frame, n_out, n_in);
    if (rc) return rc;
    MOVFE_CUDA(ctx, cudaEventRecord(w.done, ctx->raster_stream));
    w.first = first_frame;
    w.nout = n_out;
    w.nin = n_in;
    w.consumed_valid = false;
    ctx->rb_cur ^= 1;
    return MOVFE_OK;
}

static int win_index(movfe_ctx *ctx, int stream, int64_t frame, int *fi) {
    if (!ctx) return MOVFE_E_INVALID;
    if (stream < 0 || stream >= ctx->cfg.n_streams) MOVFE_FAIL(ctx, MOVFE_E_INVALID, "stream %d out of range", stream);
    const RasterBuf &w = ctx->rb[ctx->rb_cur];
    if (w.first < 0 || frame < w.first || frame >= w.first + w.nout)
        MOVFE_FAIL(ctx, MOVFE_E_STATE, "frame %lld is not in the last raster window", (long long)frame);
    *fi = (int)(frame - w.first);
    return MOVFE_OK;
}

extern "C" int movfe_raster_counts(movfe_ctx *ctx, int stream, int64_t frame, int32_t *n_hops, int32_t *n_kps,
                                   double *coverage_area) {
    int fi;
    int rc = win_index(ctx, stream, frame, &fi);
    if (rc) return rc;
    const RasterBuf &w = ctx->rb[ctx->rb_cur];
    const size_t sg = (size_t)stream * w.nin + fi;
    cudaStream_t st = ctx->raster_stream;  // the stream the results were produced on
    if (n_hops) MOVFE_CUDA(ctx, cudaMemcpyAsync(n_hops, w.d_nhops + sg, 4, cudaMemcpyDeviceToHost, st));
    if (n_kps) MOVFE_CUDA(ctx, cudaMemcpyAsync(n_kps, w.d_nkps + sg, 4, cudaMemcpyDeviceToHost, st));
    if (coverage_area) MOVFE_CUDA(ctx, cudaMemcpyAsync(coverage_area, w.d_cov + sg, 8, cudaMemcpyDeviceToHost, st));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(st));
    return MOVFE_OK;
}

extern "C" int movfe_download_grid(movfe_ctx *ctx, int stream, int64_t frame, int32_t *out) {
    int fi;
    int rc = win_index(ctx, stream, frame, &fi);
    if (rc) return rc;
    const size_t plane = (size_t)ctx->cfg.width * ctx->cfg.height;
    const RasterBuf &w = ctx->rb[ctx->rb_cur];
    if (ctx->fused) MOVFE_FAIL(ctx, MOVFE_E_STATE, "download_grid: the context was created with MOVFE_CFG_NO_GRID (the slot grid is not materialised)");
    MOVFE_CUDA(ctx, cudaMemcpyAsync(out, w.d_grid + ((size_t)stream * w.nout + fi) * plane, plane * sizeof(int4),
                                    cudaMemcpyDeviceToHost, ctx->raster_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->raster_stream));
    return MOVFE_OK;
}

extern "C" int movfe_download_hops(movfe_ctx *ctx, int stream, int64_t frame, movfe_hop *out, int capacity) {
    int fi, n = 0;
    int rc = win_index(ctx, stream, frame, &fi);
    if (rc) return rc;
    rc = movfe_raster_counts(ctx, stream, frame, &n, nullptr, nullptr);
    if (rc) return rc;
    if (n > capacity) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "download_hops: %d hops, capacity %d", n, capacity);
    if (n > 0) {
        const RasterBuf &w = ctx->rb[ctx->rb_cur];
        MOVFE_CUDA(ctx, cudaMemcpyAsync(out, w.d_hops + ((size_t)stream * w.nout + fi) * ctx->max_hops,
                                        (size_t)n * sizeof(movfe_hop), cudaMemcpyDeviceToHost, ctx->raster_stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->raster_stream));
    }
    return n;
}

extern "C" int movfe_download_kps(movfe_ctx *ctx, int stream, int64_t frame, movfe_rect *out, int capacity) {
    int fi, n = 0;
    int rc = win_index(ctx, stream, frame, &fi);
    if (rc) return rc;
    rc = movfe_raster_counts(ctx, stream, frame, nullptr, &n, nullptr);
    if (rc) return rc;
    if (n > capacity) MOVFE_FAIL(ctx, MOVFE_E_CAPACITY, "download_kps: %d kps, capacity %d", n, capacity);
    if (n > 0) {
        const RasterBuf &w = ctx->rb[ctx->rb_cur];
        MOVFE_CUDA(ctx, cudaMemcpyAsync(out, w.d_kps + ((size_t)stream * w.nout + fi) * ctx->max_kps,
                                        (size_t)n * sizeof(movfe_rect), cudaMemcpyDeviceToHost, ctx->raster_stream));
        MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->raster_stream));
    }
    return n;
}

extern "C" int64_t movfe_rejected_records(movfe_ctx *ctx) {
    if (!ctx) return -1;
    unsigned long long v = 0;
    if (cudaStreamSynchronize(ctx->ingest_stream) != cudaSuccess) return -1;  // the ingest kernels count rejects as well
    if (cudaMemcpyAsync(&v, ctx->d_rejected, sizeof v, cudaMemcpyDeviceToHost, ctx->raster_stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->raster_stream) != cudaSuccess) return -1;
    return (int64_t)v;
}

extern "C" int movfe_workload_stats(movfe_ctx *ctx, uint64_t *out, int reset) {
    if (!ctx || !out) return MOVFE_E_INVALID;
    int rc = movfe_synchronize(ctx);
    if (rc) return rc;
    MOVFE_CUDA(ctx, cudaMemcpy(out, ctx->d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) MOVFE_CUDA(ctx, cudaMemset(ctx->d_stats, 0, 8 * sizeof(unsigned long long)));
    return MOVFE_OK;
}

extern "C" int movfe_profile_enable(movfe_ctx *ctx, int on) {
    if (!ctx) return MOVFE_E_INVALID;
    ctx->prof_on = on != 0;
    return MOVFE_OK;
}

extern "C" int movfe_profile_read(movfe_ctx *ctx, double *ms, int64_t *launches, int reset) {
    if (!ctx) return MOVFE_E_INVALID;
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->ingest_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->raster_stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MOVFE_CUDA(ctx, cudaStreamSynchronize(ctx->pose_stream));
    const bool timeline = getenv("MOVFE_PROF_TIMELINE") != nullptr && !ctx->prof_spans.empty();  // development: every span against the first one's start
    for (auto &sp : ctx->prof_spans) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) ctx->prof_ms[sp.stage] += t;
        if (timeline) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, ctx->prof_spans.front().a, sp.a);
            fprintf(stderr, "span stage %d start %.3f end %.3f ms\n", sp.stage, t0, t0 + t);
        }
        ctx->prof_free.push_back(sp.a);
        ctx->prof_free.push_back(sp.b);
    }
    ctx->prof_spans.clear();
    for (int i = 0; i < MOVFE_N_STAGES; i++) {
        if (ms) ms[i] = ctx->prof_ms[i];
        if (launches) launches[i] = ctx->prof_launches[i];
        if (reset) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
    }
    return MOVFE_OK;
}
