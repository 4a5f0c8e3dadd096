// MOVExtractor_movfe.cc — drop-in for src/MOVExtractor.cc: same class, same signature. The P-frame propagation, births,
// coverage back-fill, I-frame seeding and the MERGE of LK-carried features run on the GPU through movfe_extract_frame
// (include/movfe.h). cv::calcOpticalFlowPyrLK itself is OpenCV arithmetic and is called here, on the host, with the
// reference's arguments at the reference's three call sites (src/MOVExtractor.cc:91-92, 196-197, 347-348); its results are
// handed to the device as lk_status / lk_pts / relocalisation seeds.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#ifdef MOVFE_IN_TREE
#include <opencv2/video/tracking.hpp>
#include "MOVExtractor.h"
#else
#include "MOVExtractor_movfe.h"
#endif
#include "movfe_shim.h"

namespace movfe_shim {

lk_fn lk_override = nullptr;
#ifdef MOVFE_IN_TREE
bool use_gpu_lk = false;   // in the tree OpenCV is there: its own calcOpticalFlowPyrLK is the default (bit-for-bit the reference's results)
#else
bool use_gpu_lk = true;
#endif

// cv::calcOpticalFlowPyrLK(prev, next, pts, out, status, err, Size(31,31), 3, TermCriteria(COUNT+EPS, 20, 0.01),
//                          OPTFLOW_LK_GET_MIN_EIGENVALS, 1e-4)  — MOVExtractor.cc:69,91-92
// Three providers: a test hook; movfe_lk (the same tracker on the GPU: OpenCV's arithmetic restated, status flags equal and
// positions within 5e-3 px of OpenCV's in the parity tests); OpenCV itself when built in the MoV-SLAM tree.
static void run_lk(movfe_ctx *ctx, const cv::Mat &prev_img, const cv::Mat &next_img, const std::vector<cv::Point2f> &pts, std::vector<cv::Point2f> &out,
                   std::vector<unsigned char> &status) {
    if (lk_override) {
        lk_override(prev_img, next_img, pts, out, status);
        return;
    }
    if (use_gpu_lk && ctx && !pts.empty() && prev_img.data && next_img.data && prev_img.step == next_img.step) {
        std::vector<float> in(2 * pts.size()), res(2 * pts.size()), err(pts.size());
        for (size_t i = 0; i < pts.size(); i++) {
            in[2 * i] = pts[i].x;
            in[2 * i + 1] = pts[i].y;
        }
        status.assign(pts.size(), 0);
        const int32_t off[2] = {0, (int32_t)pts.size()};
        if (movfe_lk(ctx, 1, prev_img.data, next_img.data, (int)prev_img.step, in.data(), off, 31, 3, 20, 0.01, 1e-4, res.data(), status.data(), err.data()) == MOVFE_OK) {
            out.resize(pts.size());
            for (size_t i = 0; i < pts.size(); i++) out[i] = cv::Point2f(res[2 * i], res[2 * i + 1]);
            return;
        }
    }
#ifdef MOVFE_IN_TREE
    std::vector<float> err;
    cv::TermCriteria criteria = cv::TermCriteria((cv::TermCriteria::COUNT) + (cv::TermCriteria::EPS), 20, 0.01);
    cv::calcOpticalFlowPyrLK(prev_img, next_img, pts, out, status, err, cv::Size(31, 31), 3, criteria, cv::OPTFLOW_LK_GET_MIN_EIGENVALS, 1e-4);
#else
    out.assign(pts.size(), cv::Point2f());  // no OpenCV outside the tree, no hook, no usable images: every point is lost
    status.assign(pts.size(), 0);
#endif
}

}  // namespace movfe_shim

namespace MOV_SLAM {

MOVExtractor::MOVExtractor(int threshold, double coverageThreshold, double relocalizationDistance)
    : mCurrentId(0), mThreshold(threshold), mCoverageThreshold(coverageThreshold), mRelocalizationDistance(relocalizationDistance) {}

int MOVExtractor::operator()(const shared_ptr<MotionVectorImage> &_smv, std::vector<cv::KeyPoint> &_keypoints,
                             std::vector<VideoFeature> &_vf, std::map<int, int> &_vfmap, std::vector<std::bitset<256>> &descriptors,
                             Frame *_prev_frame) {
    _keypoints.clear();                        // src/MOVExtractor.cc:66
    if (_smv->imGray.empty()) return -1;       // :71-72
    const cv::Mat &imGrey = _smv->imGray;
    const int W = imGrey.cols, H = imGrey.rows;
    movfe_ctx *ctx = movfe_shim::extractor_context(W, H, mThreshold, mCoverageThreshold, true);
    if (!ctx) return -1;
    const bool is_p = _smv->ft == FrameType::P_FRAME;

    std::vector<cv::Point2f> pts, pts_out;
    std::vector<unsigned char> status;
    std::vector<movfe_reloc_seed> reloc;
    std::vector<float> lk_pts;
    std::vector<uint8_t> lk_status;
    int n_lk = -1;
    auto hand_over = [&]() {  // LK output -> the arrays of movfe_set_lk_results (the bounds test :98,:354 is the device's)
        n_lk = (int)pts.size();
        lk_status.assign(status.begin(), status.end());
        lk_status.resize(pts.size(), 0);
        lk_pts.resize(2 * pts.size());
        for (size_t i = 0; i < pts.size(); i++) {
            lk_pts[2 * i] = pts_out[i].x;
            lk_pts[2 * i + 1] = pts_out[i].y;
        }
    };

    std::vector<movfe_track> prev;
    if (_prev_frame && !is_p) {
        // I frame (:81-120): every previous track, in TABLE order (no sort on this branch), goes to LK
        if (!_prev_frame->mvVF.empty()) {
            for (const VideoFeature &pvf : _prev_frame->mvVF) pts.push_back(pvf.pt);
            movfe_shim::run_lk(ctx, _prev_frame->imgLeft, imGrey, pts, pts_out, status);
            hand_over();
        }
    } else if (_prev_frame) {
        if (_prev_frame->mLost) {
            // lost relocalisation (:161-243): the reference keyframe's in-view map points carried into this image
            KeyFrame *lpLastKeyFrame = _prev_frame->mpReferenceKF;
            std::vector<int> trackIds;
            std::vector<cv::Point2f> kpts, kout;
            std::vector<unsigned char> kstatus;
            const std::vector<MapPoint *> vpMapPointsKF = lpLastKeyFrame->GetMapPointMatches();
            for (MapPoint *pMP : vpMapPointsKF) {  // :171-192
                if (!pMP || !pMP->mbTrackInView || pMP->isBad()) continue;
                kpts.push_back(cv::Point2f(pMP->mTrackProjX, pMP->mTrackProjY));
                trackIds.push_back(pMP->mTrackId);
            }
            if (!kpts.empty()) {
                movfe_shim::run_lk(ctx, lpLastKeyFrame->mImage, imGrey, kpts, kout, kstatus);
                const double thresholdDist = mRelocalizationDistance * sqrt(double(H * H + W * W));  // :201
                for (size_t i = 0; i < kout.size(); i++) {
                    const cv::Point2f &ptL = kpts[i], &ptR = kout[i];
                    if (kstatus[i] == 0 || ptR.x < 0 || ptR.y < 0 || ptR.x >= W || ptR.y >= H) continue;  // :207
                    const float dx = ptR.x - ptL.x, dy = ptR.y - ptL.y;
                    const double dist = std::sqrt((double)dx * dx + (double)dy * dy);  // cv::norm(Point2f), :213
                    if (dist < thresholdDist) reloc.push_back({trackIds[i], (int32_t)i, ptR.x, ptR.y});
                }
            }
        }
        // the reference sorts prev->mvVF in place (age desc, popcount desc, :249-252) - a visible side effect that
        // VideoFeature::qIndx refers to - so the same order (ties kept stable) is applied here
        std::stable_sort(_prev_frame->mvVF.begin(), _prev_frame->mvVF.end(), [](const VideoFeature &a, const VideoFeature &b) {
            if (a.age != b.age) return a.age > b.age;
            return a.desc.count() > b.desc.count();
        });
        // coverage features (:258-262, :337-377): LK from the previous image, in sorted order
        for (const VideoFeature &pvf : _prev_frame->mvVF)
            if (pvf.coverage) pts.push_back(pvf.pt);
        if (!pts.empty()) {
            movfe_shim::run_lk(ctx, _prev_frame->imgLeft, imGrey, pts, pts_out, status);
            hand_over();
        }
    }
    if (_prev_frame) {
        prev.reserve(_prev_frame->mvVF.size());
        for (const VideoFeature &vf : _prev_frame->mvVF) prev.push_back(movfe_shim::pack(vf));
    }
    std::vector<movfe_hop> hops(_smv->mvs.size());
    for (size_t i = 0; i < hops.size(); i++) hops[i] = {_smv->mvs[i].pt.x, _smv->mvs[i].pt.y, _smv->mvs[i].dIndx, 0};
    std::vector<movfe_rect> kps(_smv->kps.size());
    for (size_t i = 0; i < kps.size(); i++)
        kps[i] = {(int16_t)_smv->kps[i].x, (int16_t)_smv->kps[i].y, (int16_t)_smv->kps[i].width, (int16_t)_smv->kps[i].height};

    const uint32_t flags = (is_p ? MOVFE_FRAME_P : 0u) | MOVFE_FRAME_MV;
    // the table holds at most: every previous track, one birth per kps entry, the 16-px lattice, the relocalisation seeds
    const size_t cap = prev.size() + kps.size() + (size_t)((W + 15) / 16) * ((H + 15) / 16) + reloc.size() + 1;
    std::vector<movfe_track> out(cap);
    int32_t cid = mCurrentId;
    // imGray rows may be padded (cv::Mat::step): the stride goes with the pointer, nothing is repacked on the host
    const int n = movfe_extract_frame(ctx, flags, imGrey.data, (int)(size_t)imGrey.step, reinterpret_cast<const int32_t *>(_smv->mvi.data),
                                      hops.data(), (int)hops.size(), kps.data(), (int)kps.size(), _smv->coverageArea, prev.data(),
                                      (int)prev.size(), lk_status.data(), lk_pts.data(), n_lk, reloc.data(), (int)reloc.size(), &cid,
                                      out.data(), (int)out.size());
    if (getenv("MOVFE_SHIM_DEBUG"))
        fprintf(stderr, "shim extract: flags=%u prev=%zu hops=%zu kps=%zu cov=%.6f lk=%d reloc=%zu -> n=%d cid=%d\n", flags, prev.size(),
                hops.size(), kps.size(), _smv->coverageArea, n_lk, reloc.size(), n, cid);
    if (n < 0) {
        movfe_shim::fail(ctx, "extract_frame");
        return -1;
    }
    mCurrentId = cid;
    for (int i = 0; i < n; i++) {  // :318-331
        const VideoFeature vf = movfe_shim::unpack(out[i], i);
        _keypoints.push_back(cv::KeyPoint(vf.pt, (float)vf.mb.width));
        _vf.push_back(vf);
        _vfmap.insert({vf.trackId, i});  // first wins
        // back-fill features push no descriptor in the reference: its inner `descriptors` shadows the argument (:421,:445)
        if (!(vf.coverage && vf.qIndx < 0)) descriptors.push_back(vf.desc);
    }
    return (int)_keypoints.size();
}

}  // namespace MOV_SLAM
