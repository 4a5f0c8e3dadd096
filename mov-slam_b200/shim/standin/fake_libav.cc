// fake_libav.cc — TEST INFRASTRUCTURE. A fake libav back-end for the reference's unmodified src/VideoDecoder.cc
// (declarations in libav_standin.h): "decoding" hands out the frames of the clip installed with fake_av_install().
// The luma plane stands in for the decoded picture (sws_scale to GRAY8 copies it; BGR output is zero-filled, nothing
// on the front-end path reads it).
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "libav_standin.h"

struct SwsContext {
    int w, h;
    AVPixelFormat dst;
};

namespace {
// Installed clips. A decoder opens clip <id> through the url "fake://clip/<id>" (anything else: clip 0), so several decoder
// objects - one per stream, on their own threads (decoder_pool.cc) - can run side by side: every piece of decoder state lives in
// the context objects below, the clip table is read-only once the decoders are open.
std::vector<fake_av_clip> g_clips(1);
std::vector<std::vector<uint8_t>> g_flat(1);
AVCodec g_codec = {"fake-h264", AV_CODEC_ID_H264};
AVInputFormat g_ifmt = {"fake"};

struct FakeFormat {       // what avformat_alloc_context really hands out
    AVFormatContext pub;  // first member: the public struct the caller sees
    int clip, packets;
    AVCodecParameters par;
    AVStream stream;
    AVStream *streams[1];
};
struct FakeCodec {        // what avcodec_alloc_context3 really hands out
    AVCodecContext pub;
    int clip, next;
    AVFrameSideData sd;
};
}  // namespace

extern "C" {

void fake_av_install_clip(int id, const fake_av_clip *clip) {
    if (id < 0) return;
    if ((size_t)id >= g_clips.size()) {
        g_clips.resize((size_t)id + 1);
        g_flat.resize((size_t)id + 1);
    }
    g_clips[(size_t)id] = *clip;
    g_flat[(size_t)id].assign((size_t)clip->width * clip->height, 128);
}
void fake_av_install(const fake_av_clip *clip) { fake_av_install_clip(0, clip); }

void avdevice_register_all(void) {}
AVFormatContext *avformat_alloc_context(void) { return &((FakeFormat *)calloc(1, sizeof(FakeFormat)))->pub; }
AVInputFormat *av_find_input_format(const char *) { return &g_ifmt; }
int av_dict_set(AVDictionary **, const char *, const char *, int) { return 0; }
int avformat_open_input(AVFormatContext **ps, const char *url, AVInputFormat *, AVDictionary **) {
    FakeFormat *f = (FakeFormat *)*ps;
    const char *slash = url ? strrchr(url, '/') : nullptr;
    const int id = (slash && slash[1] >= '0' && slash[1] <= '9') ? atoi(slash + 1) : 0;
    if ((size_t)id >= g_clips.size() || g_clips[(size_t)id].width == 0) return -1;
    f->clip = id;
    f->packets = 0;
    f->par.codec_type = AVMEDIA_TYPE_VIDEO;
    f->par.codec_id = AV_CODEC_ID_H264;
    f->par.width = g_clips[(size_t)id].width;
    f->par.height = g_clips[(size_t)id].height;
    f->stream.codecpar = &f->par;
    f->stream.r_frame_rate = {30, 1};
    f->streams[0] = &f->stream;
    f->pub.nb_streams = 1;
    f->pub.streams = f->streams;
    return 0;
}
int avformat_find_stream_info(AVFormatContext *, AVDictionary **) { return 0; }
AVCodec *avcodec_find_decoder(enum AVCodecID) { return &g_codec; }
AVCodecContext *avcodec_alloc_context3(const AVCodec *) { return &((FakeCodec *)calloc(1, sizeof(FakeCodec)))->pub; }
int avcodec_parameters_to_context(AVCodecContext *c, const AVCodecParameters *par) {
    c->width = par->width;
    c->height = par->height;
    c->pix_fmt = AV_PIX_FMT_YUV420P;
    // the parameters live inside the format context they were read from: that is how the decoder learns its clip
    const FakeFormat *f = (const FakeFormat *)((const char *)par - offsetof(FakeFormat, par));
    ((FakeCodec *)c)->clip = f->clip;
    ((FakeCodec *)c)->next = 0;
    return 0;
}
int avcodec_open2(AVCodecContext *, const AVCodec *, AVDictionary **) { return 0; }
AVFrame *av_frame_alloc(void) { return (AVFrame *)calloc(1, sizeof(AVFrame)); }
AVPacket *av_packet_alloc(void) { return (AVPacket *)calloc(1, sizeof(AVPacket)); }
double av_q2d(AVRational a) { return a.num / (double)a.den; }

int av_read_frame(AVFormatContext *s, AVPacket *pkt) {
    FakeFormat *f = (FakeFormat *)s;
    if (f->packets >= g_clips[(size_t)f->clip].n_frames) return AVERROR_EOF;
    f->packets++;
    pkt->stream_index = 0;
    return 0;
}
int avcodec_send_packet(AVCodecContext *, const AVPacket *) { return 0; }
int avcodec_receive_frame(AVCodecContext *ctx, AVFrame *f) {
    FakeCodec *c = (FakeCodec *)ctx;
    const fake_av_clip &clip = g_clips[(size_t)c->clip];
    if (c->next >= clip.n_frames) return AVERROR_EOF;
    const int k = c->next++;
    memset(f, 0, sizeof *f);
    f->width = clip.width;
    f->height = clip.height;
    f->pict_type = clip.pict_is_p[k] ? AV_PICTURE_TYPE_P : AV_PICTURE_TYPE_I;
    f->data[0] = const_cast<uint8_t *>(clip.luma && clip.luma[k] ? clip.luma[k] : g_flat[(size_t)c->clip].data());
    f->linesize[0] = clip.width;
    f->side_data = nullptr;
    if (clip.side && clip.side[k]) {
        c->sd.type = AV_FRAME_DATA_MOTION_VECTORS;
        c->sd.data = const_cast<uint8_t *>(clip.side[k]);
        c->sd.size = clip.side_bytes[k];
        f->side_data = &c->sd;
    }
    return 0;
}
AVFrameSideData *av_frame_get_side_data(const AVFrame *f, enum AVFrameSideDataType type) {
    return f->side_data && f->side_data->type == type ? f->side_data : nullptr;
}
void av_packet_unref(AVPacket *) {}
void avformat_close_input(AVFormatContext **s) {
    free(*s);  // == the FakeFormat (first member)
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
