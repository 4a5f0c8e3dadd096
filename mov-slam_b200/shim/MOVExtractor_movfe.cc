// MOVExtractor_movfe.cc — drop-in for src/MOVExtractor.cc: same class, same signature, the P-frame propagation, births,
// coverage back-fill and I-frame seeding run on the GPU through movfe_extract_frame (include/movfe.h).
// What stays on the host and is NOT done here: the LK carry-over branches (cv::calcOpticalFlowPyrLK,
// src/MOVExtractor.cc:81-120,161-243,337-377); coverage features are emitted with coverage = true and are dropped at the
// next frame, exactly like the oracle with lk_status == NULL.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#ifdef MOVFE_IN_TREE
#include "MOVExtractor.h"
#else
#include "MOVExtractor_movfe.h"
#endif
#include "movfe_shim.h"

namespace MOV_SLAM {

MOVExtractor::MOVExtractor(int threshold, double coverageThreshold, double relocalizationDistance)
    : mCurrentId(0), mThreshold(threshold), mCoverageThreshold(coverageThreshold), mRelocalizationDistance(relocalizationDistance) {}

int MOVExtractor::operator()(const shared_ptr<MotionVectorImage> &_smv, std::vector<cv::KeyPoint> &_keypoints,
                             std::vector<VideoFeature> &_vf, std::map<int, int> &_vfmap, std::vector<std::bitset<256>> &descriptors,
                             Frame *_prev_frame) {
    _keypoints.clear();                        // src/MOVExtractor.cc:66
    if (_smv->imGray.empty()) return -1;       // :71-72
    const int W = _smv->imGray.cols, H = _smv->imGray.rows;
    movfe_ctx *ctx = movfe_shim::extractor_context(W, H, mThreshold, mCoverageThreshold, true);
    if (!ctx) return -1;

    // previous table: the reference sorts prev->mvVF in place (age desc, popcount desc, :249-252) - a visible side effect
    // that VideoFeature::qIndx refers to - so the same (stable) order is applied here before packing
    std::vector<movfe_track> prev;
    if (_prev_frame) {
        std::stable_sort(_prev_frame->mvVF.begin(), _prev_frame->mvVF.end(), [](const VideoFeature &a, const VideoFeature &b) {
            if (a.age != b.age) return a.age > b.age;
            return a.desc.count() > b.desc.count();
        });
        prev.reserve(_prev_frame->mvVF.size());
        for (const VideoFeature &vf : _prev_frame->mvVF) prev.push_back(movfe_shim::pack(vf));
    }
    std::vector<movfe_hop> hops(_smv->mvs.size());
    for (size_t i = 0; i < hops.size(); i++) hops[i] = {_smv->mvs[i].pt.x, _smv->mvs[i].pt.y, _smv->mvs[i].dIndx, 0};
    std::vector<movfe_rect> kps(_smv->kps.size());
    for (size_t i = 0; i < kps.size(); i++)
        kps[i] = {(int16_t)_smv->kps[i].x, (int16_t)_smv->kps[i].y, (int16_t)_smv->kps[i].width, (int16_t)_smv->kps[i].height};
    // grey plane: cv::Mat rows may be padded
    std::vector<uint8_t> grey((size_t)W * H);
    for (int y = 0; y < H; y++) memcpy(&grey[(size_t)y * W], _smv->imGray.data + (size_t)y * _smv->imGray.step, W);

    const uint32_t flags = (_smv->ft == P_FRAME ? MOVFE_FRAME_P : 0u) | MOVFE_FRAME_MV;
    std::vector<movfe_track> out(8192);
    int32_t cid = mCurrentId;
    const int n = movfe_extract_frame(ctx, flags, grey.data(), reinterpret_cast<const int32_t *>(_smv->mvi.data), hops.data(), (int)hops.size(),
                                      kps.data(), (int)kps.size(), _smv->coverageArea, prev.data(), (int)prev.size(), &cid, out.data(),
                                      (int)out.size());
    if (getenv("MOVFE_SHIM_DEBUG"))
        fprintf(stderr, "shim extract: flags=%u prev=%zu hops=%zu kps=%zu cov=%.6f -> n=%d cid=%d\n", flags, prev.size(), hops.size(), kps.size(),
                _smv->coverageArea, n, cid);
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
        descriptors.push_back(vf.desc);
    }
    return (int)_keypoints.size();
}

}  // namespace MOV_SLAM
