// fake_libav.cc — TEST INFRASTRUCTURE. A fake libav back-end for the reference's unmodified src/VideoDecoder.cc
// (declarations in libav_standin.h): "decoding" hands out the frames of the clip installed with fake_av_install().
// The luma plane stands in for the decoded picture (sws_scale to GRAY8 copies it; BGR output is zero-filled, nothing
// on the front-end path reads it).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "libav_standin.h"

struct SwsContext {
    int w, h;
    AVPixelFormat dst;
};

namespace {
fake_av_clip g_clip = {};
int g_next = 0;
std::vector<uint8_t> g_flat;
AVFrameSideData g_sd;
AVCodecParameters g_par;
AVStream g_stream;
AVStream *g_streams[1];
AVCodec g_codec = {"fake-h264", AV_CODEC_ID_H264};
AVInputFormat g_ifmt = {"fake"};
}  // namespace

extern "C" {

void fake_av_install(const fake_av_clip *clip) {
    g_clip = *clip;
    g_next = 0;
    g_flat.assign((size_t)clip->width * clip->height, 128);
}

void avdevice_register_all(void) {}
AVFormatContext *avformat_alloc_context(void) { return (AVFormatContext *)calloc(1, sizeof(AVFormatContext)); }
AVInputFormat *av_find_input_format(const char *) { return &g_ifmt; }
int av_dict_set(AVDictionary **, const char *, const char *, int) { return 0; }
int avformat_open_input(AVFormatContext **ps, const char *, AVInputFormat *, AVDictionary **) {
    g_par.codec_type = AVMEDIA_TYPE_VIDEO;
    g_par.codec_id = AV_CODEC_ID_H264;
    g_par.width = g_clip.width;
    g_par.height = g_clip.height;
    g_stream.codecpar = &g_par;
    g_stream.r_frame_rate = {30, 1};
    g_streams[0] = &g_stream;
    (*ps)->nb_streams = 1;
    (*ps)->streams = g_streams;
    return 0;
}
int avformat_find_stream_info(AVFormatContext *, AVDictionary **) { return 0; }
AVCodec *avcodec_find_decoder(enum AVCodecID) { return &g_codec; }
AVCodecContext *avcodec_alloc_context3(const AVCodec *) { return (AVCodecContext *)calloc(1, sizeof(AVCodecContext)); }
int avcodec_parameters_to_context(AVCodecContext *c, const AVCodecParameters *par) {
    c->width = par->width;
    c->height = par->height;
    c->pix_fmt = AV_PIX_FMT_YUV420P;
    return 0;
}
int avcodec_open2(AVCodecContext *, const AVCodec *, AVDictionary **) { return 0; }
AVFrame *av_frame_alloc(void) { return (AVFrame *)calloc(1, sizeof(AVFrame)); }
AVPacket *av_packet_alloc(void) { return (AVPacket *)calloc(1, sizeof(AVPacket)); }
double av_q2d(AVRational a) { return a.num / (double)a.den; }

int av_read_frame(AVFormatContext *, AVPacket *pkt) {
    if (g_next >= g_clip.n_frames) return AVERROR_EOF;
    pkt->stream_index = 0;
    return 0;
}
int avcodec_send_packet(AVCodecContext *, const AVPacket *) { return 0; }
int avcodec_receive_frame(AVCodecContext *, AVFrame *f) {
    if (g_next >= g_clip.n_frames) return AVERROR_EOF;
    const int k = g_next++;
    memset(f, 0, sizeof *f);
    f->width = g_clip.width;
    f->height = g_clip.height;
    f->pict_type = g_clip.pict_is_p[k] ? AV_PICTURE_TYPE_P : AV_PICTURE_TYPE_I;
    f->data[0] = const_cast<uint8_t *>(g_clip.luma && g_clip.luma[k] ? g_clip.luma[k] : g_flat.data());
    f->linesize[0] = g_clip.width;
    f->side_data = nullptr;
    if (g_clip.side && g_clip.side[k]) {
        g_sd.type = AV_FRAME_DATA_MOTION_VECTORS;
        g_sd.data = const_cast<uint8_t *>(g_clip.side[k]);
        g_sd.size = g_clip.side_bytes[k];
        f->side_data = &g_sd;
    }
    return 0;
}
AVFrameSideData *av_frame_get_side_data(const AVFrame *f, enum AVFrameSideDataType type) {
    return f->side_data && f->side_data->type == type ? f->side_data : nullptr;
}
void av_packet_unref(AVPacket *) {}
void avformat_close_input(AVFormatContext **s) {
    free(*s);
    *s = nullptr;
}
void av_packet_free(AVPacket **p) {
    free(*p);
    *p = nullptr;
}
void av_frame_free(AVFrame **f) {
    free(*f);
    *f = nullptr;
}
void avcodec_free_context(AVCodecContext **c) {
    free(*c);
    *c = nullptr;
}
int av_image_alloc(uint8_t *pointers[4], int linesizes[4], int w, int h, enum AVPixelFormat, int) {
    pointers[0] = (uint8_t *)malloc((size_t)w * h * 3 / 2);
    linesizes[0] = w;
    return w * h * 3 / 2;
}

SwsContext *sws_getContext(int srcW, int srcH, enum AVPixelFormat, int, int, enum AVPixelFormat dstFormat, int, SwsFilter *, SwsFilter *,
                           const double *) {
    SwsContext *c = (SwsContext *)calloc(1, sizeof(SwsContext));
    c->w = srcW;
    c->h = srcH;
    c->dst = dstFormat;
    return c;
}
int sws_scale(SwsContext *c, const uint8_t *const src[], const int srcStride[], int, int srcSliceH, uint8_t *const dst[],
              const int dstStride[]) {
    if (c->dst == AV_PIX_FMT_GRAY8) {
        for (int y = 0; y < srcSliceH; y++) memcpy(dst[0] + (size_t)y * dstStride[0], src[0] + (size_t)y * srcStride[0], c->w);
    } else {
        for (int y = 0; y < srcSliceH; y++) memset(dst[0] + (size_t)y * dstStride[0], 0, dstStride[0]);
    }
    return srcSliceH;
}
void sws_freeContext(SwsContext *c) { free(c); }

}  // extern "C"
