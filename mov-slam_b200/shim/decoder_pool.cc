// decoder_pool.cc — see decoder_pool.h.
#include "decoder_pool.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>

extern "C" {
#ifdef MOVFE_IN_TREE
#include <libavcodec/avcodec.h>
#include <libavdevice/avdevice.h>
#include <libavformat/avformat.h>
#include <libavutil/motion_vector.h>
#include <libswscale/swscale.h>
#else
#include "standin/libav_standin.h"
#endif
}

namespace movfe_shim {

static_assert(sizeof(AVMotionVector) == sizeof(movfe_mv_record), "AVMotionVector (+ref) is the 40-byte record of the C ABI");

// The libav state of one stream, set up as VideoDecoder::Init does (src/VideoDecoder.cc:37-149), and its share of the current window.
struct DecoderPool::Stream {
    AVFormatContext *fmt = nullptr;
    AVCodecContext *cc = nullptr;
    AVFrame *frame = nullptr;
    AVPacket *pkt = nullptr;
    SwsContext *to_grey = nullptr;
    int video_index = -1, W = 0, H = 0;
    float fps = 0.f;
    bool eof = false;
    // this window
    std::vector<movfe_packed_record> recs;
    std::vector<int64_t> off;     // records before frame f of this stream, n + 1 entries
    std::vector<uint8_t> flags;
    int n_done = 0;

    bool open(const std::string &path) {
        avdevice_register_all();
        if (!(fmt = avformat_alloc_context())) return false;
        AVInputFormat *ifmt = av_find_input_format("libx264");
        AVDictionary *opts = nullptr;
        av_dict_set(&opts, "flags2", "+export_mvs", 0);  // side data = AVMotionVector records (:62)
        if (avformat_open_input(&fmt, path.c_str(), ifmt, &opts) != 0) return false;
        if (avformat_find_stream_info(fmt, nullptr) < 0) return false;
        AVCodec *codec = nullptr;
        AVCodecParameters *par = nullptr;
        for (unsigned i = 0; i < fmt->nb_streams && video_index < 0; i++) {
            AVCodecParameters *p = fmt->streams[i]->codecpar;
            AVCodec *c = avcodec_find_decoder(p->codec_id);
            if (c && p->codec_type == AVMEDIA_TYPE_VIDEO) {
                video_index = (int)i;
                codec = c;
                par = p;
            }
        }
        if (video_index < 0) return false;
        if (!(cc = avcodec_alloc_context3(codec))) return false;
        if (avcodec_parameters_to_context(cc, par) < 0) return false;
        if (avcodec_open2(cc, codec, &opts) < 0) return false;
        if (!(frame = av_frame_alloc()) || !(pkt = av_packet_alloc())) return false;
        W = cc->width;
        H = cc->height;
        to_grey = sws_getContext(W, H, cc->pix_fmt, W, H, AV_PIX_FMT_GRAY8, SWS_FAST_BILINEAR, nullptr, nullptr, nullptr);
        fps = (float)av_q2d(fmt->streams[video_index]->r_frame_rate);
        return to_grey != nullptr;
    }

    void close() {
        if (fmt) avformat_close_input(&fmt);
        if (pkt) av_packet_free(&pkt);
        if (frame) av_frame_free(&frame);
        if (cc) avcodec_free_context(&cc);
        if (to_grey) sws_freeContext(to_grey);
    }

    // one picture (the loop of src/VideoDecoder.cc:163-196): records packed, flag, luma plane to `luma` (W*H, tightly packed)
    bool decode_one(uint8_t *luma) {
        while (!eof) {
            if (av_read_frame(fmt, pkt) < 0) {
                eof = true;
                break;
            }
            bool got = false, bad = false;
            while (!got) {
                if (avcodec_send_packet(cc, pkt) < 0) {
                    bad = true;
                    break;
                }
                if (pkt->stream_index != video_index) break;
                const int rc = avcodec_receive_frame(cc, frame);
                if (rc == AVERROR(EAGAIN) || rc == AVERROR_EOF) continue;
                if (rc < 0) {
                    bad = true;
                    break;
                }
                got = true;
            }
            if (got) {
                uint8_t *dst[1] = {luma};
                int ls[1] = {W};
                sws_scale(to_grey, frame->data, frame->linesize, 0, frame->height, dst, ls);  // VideoBase.h:50-59
                AVFrameSideData *sd = av_frame_get_side_data(frame, AV_FRAME_DATA_MOTION_VECTORS);
                const int64_t n = sd ? (int64_t)(sd->size / sizeof(AVMotionVector)) : 0;
                const size_t at = recs.size();
                recs.resize(at + (size_t)n);
                if (n) movfe_pack_records(reinterpret_cast<const movfe_mv_record *>(sd->data), n, recs.data() + at);
                off.push_back((int64_t)recs.size());
                flags.push_back((uint8_t)((frame->pict_type != AV_PICTURE_TYPE_I ? MOVFE_FRAME_P : 0u) | (n > 0 ? MOVFE_FRAME_MV : 0u)));
            }
            av_packet_unref(pkt);
            if (bad) eof = true;
            if (got) return true;
        }
        return false;
    }
};

DecoderPool::DecoderPool(const std::vector<std::string> &paths, int n_threads) {
    ok_ = !paths.empty();
    for (const std::string &p : paths) {
        Stream *s = new Stream;
        streams_.push_back(s);
        if (!s->open(p)) ok_ = false;
    }
    if (ok_) {
        W_ = streams_[0]->W;
        H_ = streams_[0]->H;
        fps_ = streams_[0]->fps;
        for (Stream *s : streams_) ok_ = ok_ && s->W == W_ && s->H == H_;
    }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, (int)streams_.size()));
    for (int i = 0; i < nt; i++) threads_.emplace_back(&DecoderPool::worker, this);
}

DecoderPool::~DecoderPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        quit_ = true;
    }
    cv_work_.notify_all();
    for (std::thread &t : threads_) t.join();
    for (Stream *s : streams_) {
        s->close();
        delete s;
    }
}

// A worker takes whole streams: a stream's pictures must be decoded in order by one thread (the libav context is not shared).
void DecoderPool::worker() {
    int seen = 0;
    for (;;) {
        int n_frames;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return quit_ || job_id_ != seen; });
            if (quit_) return;
            seen = job_id_;
            n_frames = job_frames_;
        }
        for (;;) {
            const int si = next_stream_.fetch_add(1);
            if (si >= (int)streams_.size()) break;
            Stream &s = *streams_[si];
            s.recs.clear();
            s.off.assign(1, 0);
            s.flags.clear();
            s.n_done = 0;
            // the luma planes go to a per-stream scratch first: the window's plane area is laid out once every stream's frame
            // count is known (a short stream shortens the window for all)
            plane_store_[si].resize((size_t)n_frames * W_ * H_);
            for (int f = 0; f < n_frames; f++) {
                if (!s.decode_one(plane_store_[si].data() + (size_t)f * W_ * H_)) break;
                s.n_done++;
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (--running_ == 0) cv_done_.notify_all();
        }
    }
}

int DecoderPool::next_window(int n_frames, HostWindow &w) {
    w.n_frames = 0;
    w.recs.clear();
    w.rec_off.assign(1, 0);
    w.flags.clear();
    if (!ok_ || n_frames < 1) return 0;
    plane_store_.resize(streams_.size());
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_frames_ = n_frames;
        next_stream_ = 0;
        running_ = (int)threads_.size();
        job_id_++;
    }
    cv_work_.notify_all();
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return running_ == 0; });
    }
    int n = n_frames;
    for (Stream *s : streams_) n = std::min(n, s->n_done);
    if (n <= 0) return 0;
    const size_t plane = (size_t)W_ * H_;
    const int S = (int)streams_.size();
    w.n_frames = n;
    w.grey.resize((size_t)S * n * plane);
    for (int si = 0; si < S; si++) {
        Stream &s = *streams_[si];
        const int64_t base = (int64_t)w.recs.size(), cnt = s.off[n];
        w.recs.insert(w.recs.end(), s.recs.begin(), s.recs.begin() + cnt);
        for (int f = 0; f < n; f++) {
            w.rec_off.push_back(base + s.off[f + 1]);
            w.flags.push_back(s.flags[f]);
        }
        memcpy(w.grey.data() + (size_t)si * n * plane, plane_store_[si].data(), (size_t)n * plane);
    }
    return n;
}

int DecoderPool::push_next_window(movfe_ctx *ctx, int n_frames, HostWindow &w) {
    const int n = next_window(n_frames, w);
    if (n <= 0) return 0;
    const int rc = movfe_push_frames_packed(ctx, n, w.recs.data(), w.rec_off.data(), w.flags.data(), w.grey.data(), 0);
    return rc == MOVFE_OK ? n : -std::abs(rc);
}

// ------------------------------------------------------------------------------------------------ trajectories ----
namespace {
struct Twc32 {
    float R[9], t[3];
};
// Sophus::SE3f(Tcw).inverse(): R^T, -(R^T t), binary32 as the reference's SE3f
Twc32 inverse32(const movfe_pose &T) {
    Twc32 o;
    float R[9], t[3];
    for (int i = 0; i < 9; i++) R[i] = (float)T.R[i];
    for (int i = 0; i < 3; i++) t[i] = (float)T.t[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.R[i * 3 + j] = R[j * 3 + i];
    for (int i = 0; i < 3; i++) o.t[i] = -(o.R[i * 3] * t[0] + o.R[i * 3 + 1] * t[1] + o.R[i * 3 + 2] * t[2]);
    return o;
}
// Eigen::Quaternionf(Matrix3f) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>): x y z w
void quat32(const float *m, float q[4]) {
    auto M = [&](int r, int c) { return m[r * 3 + c]; };
    float t = M(0, 0) + M(1, 1) + M(2, 2);
    if (t > 0.f) {
        t = std::sqrt(t + 1.0f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (M(2, 1) - M(1, 2)) * t;
        q[1] = (M(0, 2) - M(2, 0)) * t;
        q[2] = (M(1, 0) - M(0, 1)) * t;
    } else {
        int i = 0;
        if (M(1, 1) > M(0, 0)) i = 1;
        if (M(2, 2) > M(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (M(k, j) - M(j, k)) * t;
        q[j] = (M(j, i) + M(i, j)) * t;
        q[k] = (M(k, i) + M(i, k)) * t;
    }
}
}  // namespace

bool write_trajectory_tum(const std::string &path, const double *timestamps, const movfe_pose *Tcw, const uint8_t *lost, int n) {
    std::ofstream f(path.c_str());
    if (!f) return false;
    f << std::fixed;
    for (int i = 0; i < n; i++) {
        if (lost && lost[i]) continue;  // frames not localized are not saved (System.cc:398-399)
        const Twc32 T = inverse32(Tcw[i]);
        float q[4];
        quat32(T.R, q);
        f << std::setprecision(6) << timestamps[i] << " " << std::setprecision(9) << T.t[0] << " " << T.t[1] << " " << T.t[2] << " " << q[0] << " " << q[1]
          << " " << q[2] << " " << q[3] << std::endl;  // System.cc:419
    }
    return true;
}

bool write_trajectory_kitti(const std::string &path, const int64_t *frame_ids, const movfe_pose *Tcw, const uint8_t *lost, int n) {
    std::ofstream f(path.c_str());
    if (!f) return false;
    f << std::fixed;
    for (int i = 0; i < n; i++) {
        if (lost && lost[i]) continue;
        const Twc32 T = inverse32(Tcw[i]);
        f << std::setprecision(9) << (frame_ids ? frame_ids[i] : (int64_t)i) << " " << T.R[0] << " " << T.R[1] << " " << T.R[2] << " " << T.t[0] << " " << T.R[3] << " " << T.R[4] << " " << T.R[5] << " "
          << T.t[1] << " " << T.R[6] << " " << T.R[7] << " " << T.R[8] << " " << T.t[2] << std::endl;  // System.cc:832-834
    }
    return true;
}

}  // namespace movfe_shim
