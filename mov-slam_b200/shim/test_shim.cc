// test_shim.cc — drives the drop-in shims the way Tracking.cc / the example mains drive the reference classes
// (mono_video_tartan.cc:74 NextImage -> Frame.cc:390-393 MOVExtractor -> Tracking.cc:804-811 SearchByVideoFeature +
// PoseOptimization) on inputs written by tests/test_shim.py, and dumps the results for comparison with the oracle.
// Built against the stand-in headers; links libmovfe.so. usage: test_shim <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "MOVExtractor_movfe.h"
#include "MOVMatcher_movfe.h"
#include "movfe_shim.h"

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
    const std::vector<int32_t> meta = slurp<int32_t>(dir + "/meta.bin");  // W H NF max_ref threshold n_map n_kf
    const int W = meta[0], H = meta[1], NF = meta[2], K = meta[3], thr = meta[4], n_map = meta[5], n_kf = meta[6];
    const std::vector<movfe_mv_record> recs = slurp<movfe_mv_record>(dir + "/recs.bin");
    const std::vector<int64_t> off = slurp<int64_t>(dir + "/off.bin");
    const std::vector<uint8_t> flags = slurp<uint8_t>(dir + "/flags.bin");
    const std::vector<uint8_t> grey = slurp<uint8_t>(dir + "/grey.bin");
    const std::vector<movfe_map_point> mps = slurp<movfe_map_point>(dir + "/map.bin");
    const std::vector<double> pose0 = slurp<double>(dir + "/pose0.bin");  // R(9) t(3)
    const std::vector<float> camp = slurp<float>(dir + "/cam.bin");       // fx fy cx cy

    // --- decoder side: RasterQueue stands where VideoDecoder::NextImage's MV loop was ---------------------------------
    movfe_shim::RasterQueue rq(W, H, K);
    MOVExtractor extractor(thr, 0.20, 0.25);
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
        const int n = extractor(img, F->mvKeys, F->mvVF, F->mvVFMap, F->mDescriptors, prev);
        F->N = n < 0 ? 0 : n;
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
    for (int f = 0; f < NF; f++) {
        auto img = std::make_shared<MotionVectorImage>(W, H);
        img->imGray = cv::Mat(H, W, 1);
        memcpy(img->imGray.data, &grey[(size_t)f * W * H], (size_t)W * H);
        img->ft = (flags[f] & MOVFE_FRAME_P) ? P_FRAME : I_FRAME;
        const int n = (int)(off[f + 1] - off[f]);
        if (!rq.push(img, n ? &recs[off[f]] : nullptr, n, (flags[f] & MOVFE_FRAME_MV) != 0)) return 3;
        while (auto done = rq.pop(false)) consume(done);
    }
    while (auto done = rq.pop(true)) consume(done);
    printf("test_shim: %d frames\n", produced);
    return produced == NF ? 0 : 4;
}
