// test_shim.cc — drives the drop-in shims the way Tracking.cc / the example mains drive the reference classes
// (mono_video_tartan.cc:74 NextImage -> Frame.cc:390-393 MOVExtractor -> Tracking.cc:804-811 SearchByVideoFeature +
// PoseOptimization) on inputs written by tests/test_shim.py, and dumps the results for comparison with the oracle.
// The frames come out of the VideoDecoder shim (VideoDecoder_movfe.cc) running on the fake libav back-end
// (standin/fake_libav.cc): demux / decode are faked, the side data takes the GPU raster path. cv::calcOpticalFlowPyrLK is
// replaced by a deterministic hook that tests/test_shim.py mirrors, so the LK hand-over (I-frame carry-over, coverage
// tracks) is exercised end to end. Built against the stand-in headers; links libmovfe.so. usage: test_shim <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "MOVExtractor_movfe.h"
#include "MOVMatcher_movfe.h"
#include "VideoDecoder_movfe.h"
#include "movfe_shim.h"

// deterministic stand-in for cv::calcOpticalFlowPyrLK (mirrored by tests/test_shim.py::pseudo_lk)
static void pseudo_lk(const cv::Mat &, const cv::Mat &, const std::vector<cv::Point2f> &pts, std::vector<cv::Point2f> &out,
                      std::vector<unsigned char> &status) {
    out.resize(pts.size());
    status.resize(pts.size());
    for (size_t i = 0; i < pts.size(); i++) {
        out[i] = cv::Point2f(pts[i].x + 0.5f, pts[i].y - 0.25f);
        status[i] = (((int)pts[i].x * 7 + (int)pts[i].y * 3) % 5) != 0;
    }
}

namespace MOV_SLAM {
class Optimizer {
public:
    static int PoseOptimization(Frame *pFrame, const bool isLost, const int iterationCount = 50, const double reprojectionError = 5.0,
                                const double reprojectErrorLost = 8.0, const double confidence = 0.95, const int algorithm = 38);
};
}  // namespace MOV_SLAM

using namespace MOV_SLAM;

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
static void dump(const std::string &path, const std::vector<T> &v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    const std::vector<int32_t> meta = slurp<int32_t>(dir + "/meta.bin");  // W H NF max_ref threshold n_map n_kf [cov_thr permille]
    const int W = meta[0], H = meta[1], NF = meta[2], K = meta[3], thr = meta[4], n_map = meta[5], n_kf = meta[6];
    const double cov_thr = meta.size() > 7 ? meta[7] / 1000.0 : 0.20;
    // SHIM_LK=gpu: no hook - the extractor shim's own provider, movfe_lk (the GPU tracker), runs at the reference's call sites
    movfe_shim::lk_override = (getenv("SHIM_LK") && std::string(getenv("SHIM_LK")) == "gpu") ? nullptr : pseudo_lk;
    const std::vector<movfe_mv_record> recs = slurp<movfe_mv_record>(dir + "/recs.bin");
    const std::vector<int64_t> off = slurp<int64_t>(dir + "/off.bin");
    const std::vector<uint8_t> flags = slurp<uint8_t>(dir + "/flags.bin");
    const std::vector<uint8_t> grey = slurp<uint8_t>(dir + "/grey.bin");
    const std::vector<movfe_map_point> mps = slurp<movfe_map_point>(dir + "/map.bin");
    const std::vector<double> pose0 = slurp<double>(dir + "/pose0.bin");  // R(9) t(3)
    const std::vector<float> camp = slurp<float>(dir + "/cam.bin");       // fx fy cx cy

    // --- decoder side: the VideoDecoder shim on the fake libav back-end (qlen = K + 2: the reference's 12 for K = 10) -------
    std::vector<uint8_t> is_p(NF);
    std::vector<const uint8_t *> luma(NF), side(NF, nullptr);
    std::vector<int> side_bytes(NF, 0);
    for (int f = 0; f < NF; f++) {
        is_p[f] = (flags[f] & MOVFE_FRAME_P) ? 1 : 0;
        luma[f] = &grey[(size_t)f * W * H];
        if ((flags[f] & MOVFE_FRAME_MV) && off[f + 1] > off[f]) {
            side[f] = reinterpret_cast<const uint8_t *>(&recs[off[f]]);
            side_bytes[f] = (int)((off[f + 1] - off[f]) * (int64_t)sizeof(movfe_mv_record));
        }
    }
    fake_av_clip fc = {W, H, NF, is_p.data(), luma.data(), side.data(), side_bytes.data()};
    fake_av_install(&fc);
    VideoDecoder decoder("fake://clip", K + 2);
    if (!decoder.Init() || decoder.GetWidth() != W || decoder.GetHeight() != H) return 3;
    MOVExtractor extractor(thr, cov_thr, 0.25);
    GeometricCamera cam(std::vector<float>(camp.begin(), camp.end()), GeometricCamera::CAM_PINHOLE);
    std::vector<MapPoint> points(n_map);
    std::vector<MapPoint *> local;
    KeyFrame kf;
    for (int i = 0; i < n_map; i++) {
        for (int a = 0; a < 3; a++) points[i].mWorldPos(a) = mps[i].pos[a];
        points[i].mTrackId = mps[i].track_id;
        points[i].mbBad = mps[i].flags & MOVFE_MP_BAD;
        local.push_back(&points[i]);
        if (i < n_kf) kf.mvpMapPoints.push_back((mps[i].flags & MOVFE_MP_NULL) ? nullptr : &points[i]);
    }
    Eigen::Matrix3f R0;
    Eigen::Vector3f t0;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) R0(r, c) = (float)pose0[r * 3 + c];
        t0(r) = (float)pose0[9 + r];
    }
    Sophus::SE3f last_pose(R0, t0);

    std::vector<std::shared_ptr<Frame>> frames;
    Frame *prev = nullptr;
    int produced = 0;
    auto consume = [&](const std::shared_ptr<MotionVectorImage> &img) {
        // Frame ctor -> ExtractMOV (Frame.cc:390-393)
        auto F = std::make_shared<Frame>();
        F->mpCamera = &cam;
        F->imgLeft = img->imGray;  // Frame.cc:121
        F->imageCols = W;
        F->imageRows = H;
        {   // what the decoder shim delivered: hop list, kps, slot grid, coverage
            std::vector<movfe_hop> hv;
            for (const auto &m : img->mvs) hv.push_back({m.pt.x, m.pt.y, m.dIndx, 0});
            std::vector<movfe_rect> kv;
            for (const auto &r : img->kps) kv.push_back({(int16_t)r.x, (int16_t)r.y, (int16_t)r.width, (int16_t)r.height});
            dump(dir + "/out_hops_" + std::to_string(produced) + ".bin", hv);
            dump(dir + "/out_kps_" + std::to_string(produced) + ".bin", kv);
            std::vector<int32_t> gv(reinterpret_cast<const int32_t *>(img->mvi.data), reinterpret_cast<const int32_t *>(img->mvi.data) + (size_t)W * H * 4);
            dump(dir + "/out_grid_" + std::to_string(produced) + ".bin", gv);
            dump(dir + "/out_cov_" + std::to_string(produced) + ".bin", std::vector<double>{img->coverageArea, (double)img->frame,
                                                                                            (double)(img->ft == FrameType::P_FRAME)});
            dump(dir + "/out_ndesc_" + std::to_string(produced) + ".bin", std::vector<int32_t>{0});
        }
        const int n = extractor(img, F->mvKeys, F->mvVF, F->mvVFMap, F->mDescriptors, prev);
        F->N = n < 0 ? 0 : n;
        dump(dir + "/out_ndesc_" + std::to_string(produced) + ".bin", std::vector<int32_t>{(int32_t)F->mDescriptors.size()});
        F->mvKeysUn = F->mvKeys;
        F->mvpMapPoints.assign(F->N, nullptr);
        F->mvbOutlier.assign(F->N, false);
        std::vector<movfe_track> tab;
        for (const auto &vf : F->mvVF) tab.push_back(movfe_shim::pack(vf));
        dump(dir + "/out_tracks_" + std::to_string(produced) + ".bin", tab);
        // TrackReferenceKeyFrame (Tracking.cc:796-811)
        std::vector<MapPoint *> matches;
        const int nm = MOVMatcher::SearchByVideoFeature(&kf, *F, matches);
        F->mvpMapPoints = matches;
        F->SetPose(last_pose);
        const int ninl = Optimizer::PoseOptimization(F.get(), false);
        std::vector<int32_t> m(F->N, -1);
        for (int i = 0; i < F->N; i++)
            if (F->mvpMapPoints[i]) m[i] = (int32_t)(F->mvpMapPoints[i] - points.data());
        dump(dir + "/out_match_" + std::to_string(produced) + ".bin", m);
        std::vector<double> pose(14);
        const Sophus::SE3f T = F->GetPose();
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) pose[r * 3 + c] = T.rotationMatrix()(r, c);
            pose[9 + r] = T.translation()(r);
        }
        pose[12] = nm;
        pose[13] = ninl;
        dump(dir + "/out_pose_" + std::to_string(produced) + ".bin", pose);
        std::vector<uint8_t> ol(F->N);
        for (int i = 0; i < F->N; i++) ol[i] = F->mvbOutlier[i];
        dump(dir + "/out_outlier_" + std::to_string(produced) + ".bin", ol);
        if (ninl > 0) last_pose = T;
        frames.push_back(F);
        prev = F.get();
        produced++;
    };
    while (auto img = decoder.NextImage(true)) consume(img);
    if (!frames.empty() && frames.back()->N > 0) {
        // MOVMatcher::SearchByProjection (the shim's addition): every keypoint of the last frame as a map point that projects onto
        // it and carries its descriptor. Keypoints with a twin (same descriptor within reach) may trade places; the others must
        // come back as the identity.
        Frame &F = *frames.back();
        F.imageCols = W;
        F.imageRows = H;
        std::vector<MapPoint> mps((size_t)F.N);
        std::vector<MapPoint *> list;
        for (int i = 0; i < F.N; i++) {
            mps[i].mbTrackInView = true;
            mps[i].mTrackProjX = F.mvKeysUn[i].pt.x;
            mps[i].mTrackProjY = F.mvKeysUn[i].pt.y;
            mps[i].mTrackViewCos = 0.9f;
            mps[i].mTrackDepth = 1.f;
            mps[i].mDescriptor = F.mvVF[i].desc;
            list.push_back(&mps[i]);
        }
        F.mvpMapPoints.assign(F.N, nullptr);
        const int nsp = MOVMatcher::SearchByProjection(F, list, 1.0f, false, 0.f, 100, 1.0f);
        int ident = 0;
        for (int i = 0; i < F.N; i++) ident += F.mvpMapPoints[i] == &mps[i];
        printf("test_shim: search_by_projection %d matches, %d identities of %d keypoints\n", nsp, ident, F.N);
    }
    printf("test_shim: %d frames, %lld carried tracks dropped for want of LK results\n", produced,
           (long long)movfe_dropped_lk_tracks(movfe_shim::extractor_context(W, H, thr, cov_thr, true)));
    return produced == NF ? 0 : 4;
}
