// Same public interface as the reference's include/VideoDecoder.h:28-60 + include/VideoBase.h:40-91 (used when the shim is
// built outside the MoV-SLAM tree, against the libav stand-in of standin/libav_standin.h). In the tree the reference's own
// headers declare the class and only src/VideoDecoder.cc is replaced by VideoDecoder_movfe.cc.
#pragma once
#include <string>

#include "movfe_shim.h"
extern "C" {
#include "standin/libav_standin.h"
}

namespace MOV_SLAM {
class VideoBase {
public:
    inline float fps(void) { return mFPS; }

protected:
    SwsContext *conversion_rgb = nullptr;
    SwsContext *conversion_grey = nullptr;
    SwsContext *conversion_yuv = nullptr;
    float mFPS = 0.f;
};

class VideoDecoder : public VideoBase {
public:
    VideoDecoder(const std::string &path, int qlen);
    virtual ~VideoDecoder(void);
    bool Init(void);
    shared_ptr<MotionVectorImage> NextImage(bool mv = true);
    int GetWidth();
    int GetHeight();

private:
    const std::string dataset_path;
    int video_stream_index;
    int qlen;
    int frames;
    AVFormatContext *pFormatContext = nullptr;
    AVInputFormat *inputFormat = nullptr;
    AVCodec *pCodec = nullptr;
    AVCodecParameters *pCodecParameters = nullptr;
    AVCodecContext *pCodecContext = nullptr;
    AVFrame *pFrame = nullptr;
    AVPacket *pPacket = nullptr;
};
}  // namespace MOV_SLAM
