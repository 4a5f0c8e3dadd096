// batched_frontend.cc — the batched front-end driven from C++ through the C ABI alone (include/movfe.h): S streams
// replaying one clip, window after window, enqueued back to back exactly as INTEGRATION.md §3 shows (push of window k+1
// right after the calls of window k, one device->host read per window). Inputs are the dump written by
// tests/test_shim.py::test_batched_cpp_driver; the track tables and poses of the last window are dumped for comparison
// with the oracle. usage: batched_frontend <dir>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "movfe.h"

template <typename T>
static std::vector<T> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path.c_str());
        exit(2);
    }
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> v(raw.size() / sizeof(T));
    memcpy(v.data(), raw.data(), v.size() * sizeof(T));
    return v;
}

template <typename T>
static void dump(const std::string &path, const T *p, size_t n) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(p), (std::streamsize)(n * sizeof(T)));
}

#define CK(call)                                                                      \
    do {                                                                              \
        const int _rc = (call);                                                       \
        if (_rc < 0) {                                                                \
            fprintf(stderr, "%s failed (%d): %s\n", #call, _rc, movfe_last_error(ctx)); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string d = argv[1];
    const auto meta = slurp<int32_t>(d + "/meta.bin");  // W H NF K thr n_map n_kf S F
    const int W = meta[0], H = meta[1], NF = meta[2], K = meta[3], thr = meta[4], n_map = meta[5], n_kf = meta[6], S = meta[7], F = meta[8];
    const auto recs = slurp<movfe_mv_record>(d + "/recs.bin");
    const auto off = slurp<int64_t>(d + "/off.bin");
    const auto flags = slurp<uint8_t>(d + "/flags.bin");
    const auto grey = slurp<uint8_t>(d + "/grey.bin");
    const auto map = slurp<movfe_map_point>(d + "/map.bin");
    const auto pose0 = slurp<double>(d + "/pose0.bin");
    const auto camv = slurp<float>(d + "/cam.bin");
    const int LA = K + 1, NW = (NF - LA) / F;
    const size_t plane = (size_t)W * H;

    movfe_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_streams = S;
    cfg.width = W;
    cfg.height = H;
    cfg.max_records_per_frame = 4800;
    cfg.max_ref = K;
    cfg.window_frames = F;
    cfg.max_tracks = 8192;
    cfg.max_map_points = 2048;
    cfg.express_threshold = thr;
    cfg.coverage_threshold = 0.20;
    cfg.has_grey = 1;
    movfe_ctx *ctx = nullptr;
    if (movfe_create(&cfg, &ctx) != MOVFE_OK) {
        fprintf(stderr, "movfe_create: %s\n", movfe_last_error(nullptr));
        return 1;
    }
    movfe_camera cam;
    memset(&cam, 0, sizeof cam);
    cam.model = MOVFE_CAM_PINHOLE;
    cam.fx = camv[0], cam.fy = camv[1], cam.cx = camv[2], cam.cy = camv[3];
    movfe_pose_params pp;
    memset(&pp, 0, sizeof pp);
    pp.iteration_count = 50, pp.reprojection_error = 5.0, pp.reprojection_error_lost = 8.0, pp.confidence = 0.95, pp.algorithm = 38;
    CK(movfe_set_camera(ctx, &cam, &pp, 0.5f));
    movfe_pose p0;
    memcpy(p0.R, pose0.data(), 9 * sizeof(double));
    memcpy(p0.t, pose0.data() + 9, 3 * sizeof(double));
    for (int s = 0; s < S; s++) {
        CK(movfe_set_map_points(ctx, s, map.data(), n_map, n_kf));
        CK(movfe_set_pose(ctx, s, &p0));
    }

    // stream-major packing of frames [f0, f1) of S streams that all replay the clip; three rotating buffers because a
    // host buffer must stay unchanged until the second following push
    struct Pack {
        std::vector<movfe_packed_record> r;  // 16-byte records: packed while the side data is copied (what a decoder pool does)
        std::vector<int64_t> o;
        std::vector<uint8_t> fl, g;
    } packs[3];
    int turn = 0;
    auto push = [&](int f0, int f1) -> int {
        Pack &pk = packs[turn++ % 3];
        const int n = f1 - f0;
        pk.r.clear(), pk.o.assign(1, 0), pk.fl.clear();
        pk.g.resize((size_t)S * n * plane);
        for (int s = 0; s < S; s++)
            for (int f = f0; f < f1; f++) {
                const size_t at = pk.r.size();
                pk.r.resize(at + (size_t)(off[f + 1] - off[f]));
                movfe_pack_records(recs.data() + off[f], off[f + 1] - off[f], pk.r.data() + at);
                pk.o.push_back((int64_t)pk.r.size());
                pk.fl.push_back(flags[f]);
                memcpy(&pk.g[((size_t)s * n + (f - f0)) * plane], &grey[(size_t)f * plane], plane);
            }
        return movfe_push_frames_packed(ctx, n, pk.r.data(), pk.o.data(), pk.fl.data(), pk.g.data(), 0);
    };
    CK(push(0, F + LA));
    std::vector<movfe_pose> poses((size_t)S * F);
    std::vector<int32_t> ninl((size_t)S * F);
    for (int k = 0; k < NW; k++) {
        const int first = F * k;
        CK(movfe_raster(ctx, first, F));
        CK(movfe_extract(ctx, first, F));
        CK(movfe_track_poses(ctx, first, F));
        if (k + 1 < NW) CK(push(F * (k + 1) + LA, F * (k + 2) + LA));
        CK(movfe_download_poses(ctx, first, F, poses.data(), ninl.data()));  // the only call of the loop that waits
    }
    const int last = F * (NW - 1);
    std::vector<movfe_track> tr(8192);
    for (int s : {0, S - 1})
        for (int f = last; f < last + F; f++) {
            const int n = movfe_download_tracks(ctx, s, f, tr.data(), (int)tr.size());
            if (n < 0) {
                fprintf(stderr, "download_tracks: %s\n", movfe_last_error(ctx));
                return 1;
            }
            dump(d + "/b_tracks_" + std::to_string(s) + "_" + std::to_string(f) + ".bin", tr.data(), (size_t)n);
        }
    dump(d + "/b_poses.bin", poses.data(), poses.size());
    dump(d + "/b_ninl.bin", ninl.data(), ninl.size());
    printf("batched_frontend: %d streams x %d windows of %d frames, rejected records %lld\n", S, NW, F, (long long)movfe_rejected_records(ctx));
    movfe_destroy(ctx);
    return 0;
}
