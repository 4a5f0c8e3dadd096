// movfe_shim.h — drop-in C++ layer over the C ABI (include/movfe.h) with the reference's class signatures
// (INTEGRATION.md). Built inside the MoV-SLAM tree with -DMOVFE_IN_TREE it includes the reference's own
// headers; built here (tests) it uses the stand-in declarations under standin/.
#pragma once
#ifdef MOVFE_IN_TREE
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#else
#include "standin/mov_slam_min.h"
#endif
#include <string>

#include "movfe.h"

namespace movfe_shim {

// Process-wide one-stream contexts, keyed by what fixes their buffers. The reference runs one Tracking thread
// (SURVEY.md §8b), so there is no locking; a context is created on first use and lives until process exit.
movfe_ctx *extractor_context(int width, int height, int threshold, double coverage_threshold, bool has_grey);
movfe_ctx *operator_context();  // joins / frustum / pose: geometry-only, no frame buffers
movfe_ctx *frame_operator_context(int width, int height);  // the same for operators that read the frame size (bucket grid)
void fail(movfe_ctx *ctx, const char *what);  // prints movfe_last_error to stderr like the reference's cerr paths

// Test hook: when set, replaces cv::calcOpticalFlowPyrLK at the three call sites of MOVExtractor::operator() (outside the
// MoV-SLAM tree there is no OpenCV; without a hook every carried point counts as lost).
typedef void (*lk_fn)(const cv::Mat &prev_img, const cv::Mat &next_img, const std::vector<cv::Point2f> &pts,
                      std::vector<cv::Point2f> &out, std::vector<unsigned char> &status);
extern lk_fn lk_override;
extern bool use_gpu_lk;   // movfe_lk instead of cv::calcOpticalFlowPyrLK at the reference's three call sites (default outside the MoV-SLAM tree)

movfe_track pack(const MOV_SLAM::VideoFeature &vf);
MOV_SLAM::VideoFeature unpack(const movfe_track &t, int index);
movfe_camera pack(MOV_SLAM::GeometricCamera *cam);

// Replaces the MV loop of VideoDecoder::NextImage (src/VideoDecoder.cc:211-350). The decoder keeps demuxing/decoding
// on the host and hands every frame's AVMotionVector side data (40-byte records, ffmpeg-ref-patch.patch:122-129) to
// push(); a frame's hop list / kps / slot grid are complete once max_ref+1 later frames were pushed (the reference's
// deque look-ahead), at which point pop() fills its MotionVectorImage exactly as the reference's loop would have.
class RasterQueue {
public:
    RasterQueue(int width, int height, int max_ref, int max_records_per_frame = 0);
    ~RasterQueue();
    // frame type / mv flag as in VideoDecoder.cc:193,200; side_data may be null when n_records == 0
    bool push(const std::shared_ptr<MOV_SLAM::MotionVectorImage> &img, const void *side_data, int n_records, bool mv);
    // next frame whose raster is final, or nullptr; flush = true at end of stream (no more look-ahead will come)
    std::shared_ptr<MOV_SLAM::MotionVectorImage> pop(bool flush = false);
    int pending() const;  // frames pushed and not yet popped
private:
    struct Impl;
    Impl *d;
};

}  // namespace movfe_shim
