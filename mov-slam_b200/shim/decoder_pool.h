// decoder_pool.h — batched host side of the front-end: one libavcodec context per stream, decoded on a pool
// of host threads, handed to the GPU as the stream-major windows movfe_push_frames_packed takes; and the trajectory writers of
// src/System.cc:363-423,778-838 for the batched poses.
//
// The reference runs ONE VideoDecoder on the tracking thread (mono_video_tartan.cc:74): demux + decode + colour conversion + the MV
// loop, one frame at a time. A batched front-end needs S of them in step. The pool keeps the per-stream libav state exactly as
// VideoDecoder::Init sets it up (src/VideoDecoder.cc:37-149: export_mvs, first video stream, GRAY8 conversion), decodes the next F
// pictures of every stream in parallel, packs every picture's AVMotionVector side data to the 16-byte records while it copies it out
// of the AVFrame (the side data belongs to the frame and is gone at the next receive), and lays the window out as
// [records, stream-major then frame order][offsets][flags][luma planes]. H.264 decode itself stays libavcodec's (NVDEC exports no
// motion vectors); built against the fake libav of standin/ here, against FFmpeg 4.4.3 + ffmpeg-ref-patch in the MoV-SLAM tree.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "movfe.h"

namespace movfe_shim {

// One window of every stream in the layout of movfe_push_frames_packed. Keep it unchanged until the second following push (the
// library reads the caller's buffers asynchronously); pin the vectors' storage with cudaHostRegister for full PCIe speed.
struct HostWindow {
    int n_frames = 0;                        // frames per stream in this window (fewer than asked for at the end of the streams)
    std::vector<movfe_packed_record> recs;   // all records, stream-major then frame order
    std::vector<int64_t> rec_off;            // n_streams * n_frames + 1
    std::vector<uint8_t> flags;              // n_streams * n_frames, MOVFE_FRAME_*
    std::vector<uint8_t> grey;               // n_streams * n_frames planes of width * height bytes
};

class DecoderPool {
public:
    // paths: one video per stream. n_threads <= 0: one per hardware thread, at most one per stream.
    DecoderPool(const std::vector<std::string> &paths, int n_threads);
    ~DecoderPool();
    bool ok() const { return ok_; }   // every stream opened, all of one size
    int width() const { return W_; }
    int height() const { return H_; }
    int n_streams() const { return (int)streams_.size(); }
    float fps() const { return fps_; }
    // Decodes the next n_frames pictures of every stream (all streams advance in step; a window ends at the shortest stream's end).
    // Returns the number of frames per stream placed in `w` (0: end of the streams).
    int next_window(int n_frames, HostWindow &w);
    // next_window + movfe_push_frames_packed. Returns frames pushed per stream, or a negative MOVFE_E_* code.
    int push_next_window(movfe_ctx *ctx, int n_frames, HostWindow &w);

private:
    struct Stream;
    void worker();
    std::vector<Stream *> streams_;
    std::vector<std::vector<uint8_t>> plane_store_;   // per stream: the luma planes of the window being decoded
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::atomic<int> next_stream_{0};
    int job_frames_ = 0, job_id_ = 0, running_ = 0;
    bool quit_ = false, ok_ = false;
    int W_ = 0, H_ = 0;
    float fps_ = 0.f;
};

// src/System.cc:363-423 (TUM: "timestamp tx ty tz qx qy qz qw", Twc) and :778-838 (this fork's KITTI form: frame id, then the 3x4 of Twc row-major), for the poses
// of one stream as the batched front-end returns them (Tcw per frame); frames with lost[i] != 0 are skipped like the reference's
// mlbLost entries. The arithmetic is the reference's binary32 (Sophus::SE3f inverse, Eigen's quaternion-from-matrix).
bool write_trajectory_tum(const std::string &path, const double *timestamps, const movfe_pose *Tcw, const uint8_t *lost, int n);
bool write_trajectory_kitti(const std::string &path, const int64_t *frame_ids, const movfe_pose *Tcw, const uint8_t *lost, int n);

}  // namespace movfe_shim
