/* libav_standin.h — TEST INFRASTRUCTURE. Declarations of the libavcodec / libavformat / libswscale entry points and
 * structures that the reference's src/VideoDecoder.cc and include/VideoBase.h use, restated from FFmpeg 4.4's public
 * API so that those sources compile UNMODIFIED here (FFmpeg is absent from this image). The implementations are a FAKE
 * decoder (fake_libav.cc): avcodec_receive_frame hands out the frames of a clip the test driver installed (size,
 * picture type, luma plane, motion-vector side data); no bitstream is involved. AVMotionVector is FFmpeg 4.4.3's
 * libavutil/motion_vector.h plus the `ref` member added by the reference's ffmpeg-ref-patch.patch:122-129. */
#pragma once
#include <errno.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AVMotionVector {
    int32_t source;
    uint8_t w, h;
    int16_t src_x, src_y;
    int16_t dst_x, dst_y;
    uint64_t flags;
    int32_t motion_x, motion_y;
    uint16_t motion_scale;
    int32_t ref; /* ffmpeg-ref-patch.patch:122-129 */
} AVMotionVector;

#define AVERROR(e) (-(e))
#define AVERROR_EOF (-0x20464f45) /* -MKTAG('E','O','F',' ') */

enum AVMediaType { AVMEDIA_TYPE_UNKNOWN = -1, AVMEDIA_TYPE_VIDEO = 0, AVMEDIA_TYPE_AUDIO = 1 };
enum AVPictureType { AV_PICTURE_TYPE_NONE = 0, AV_PICTURE_TYPE_I, AV_PICTURE_TYPE_P, AV_PICTURE_TYPE_B };
enum AVPixelFormat { AV_PIX_FMT_NONE = -1, AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_BGR24 = 3, AV_PIX_FMT_GRAY8 = 8 };
enum AVFrameSideDataType { AV_FRAME_DATA_MOTION_VECTORS = 8 };
enum AVCodecID { AV_CODEC_ID_NONE = 0, AV_CODEC_ID_H264 = 27 };
#define SWS_FAST_BILINEAR 1

typedef struct AVRational { int num, den; } AVRational;
typedef struct AVDictionary AVDictionary;
typedef struct AVInputFormat { const char *name; } AVInputFormat;
typedef struct AVCodec { const char *name; enum AVCodecID id; } AVCodec;
typedef struct AVCodecParameters { enum AVMediaType codec_type; enum AVCodecID codec_id; int width, height; } AVCodecParameters;
typedef struct AVStream { AVCodecParameters *codecpar; AVRational r_frame_rate; } AVStream;
typedef struct AVFormatContext { unsigned int nb_streams; AVStream **streams; } AVFormatContext;
typedef struct AVCodecContext { int width, height; enum AVPixelFormat pix_fmt; } AVCodecContext;
typedef struct AVPacket { int stream_index; } AVPacket;
typedef struct AVFrameSideData { enum AVFrameSideDataType type; uint8_t *data; int size; } AVFrameSideData;
typedef struct AVFrame {
    uint8_t *data[8];
    int linesize[8];
    int width, height;
    enum AVPictureType pict_type;
    AVFrameSideData *side_data; /* fake: at most one entry */
} AVFrame;
typedef struct SwsContext SwsContext;
typedef struct SwsFilter SwsFilter;

void avdevice_register_all(void);
AVFormatContext *avformat_alloc_context(void);
AVInputFormat *av_find_input_format(const char *short_name);
int av_dict_set(AVDictionary **pm, const char *key, const char *value, int flags);
int avformat_open_input(AVFormatContext **ps, const char *url, AVInputFormat *fmt, AVDictionary **options);
int avformat_find_stream_info(AVFormatContext *ic, AVDictionary **options);
AVCodec *avcodec_find_decoder(enum AVCodecID id);
AVCodecContext *avcodec_alloc_context3(const AVCodec *codec);
int avcodec_parameters_to_context(AVCodecContext *codec, const AVCodecParameters *par);
int avcodec_open2(AVCodecContext *avctx, const AVCodec *codec, AVDictionary **options);
AVFrame *av_frame_alloc(void);
AVPacket *av_packet_alloc(void);
double av_q2d(AVRational a);
int av_read_frame(AVFormatContext *s, AVPacket *pkt);
int avcodec_send_packet(AVCodecContext *avctx, const AVPacket *avpkt);
int avcodec_receive_frame(AVCodecContext *avctx, AVFrame *frame);
AVFrameSideData *av_frame_get_side_data(const AVFrame *frame, enum AVFrameSideDataType type);
void av_packet_unref(AVPacket *pkt);
void avformat_close_input(AVFormatContext **s);
void av_packet_free(AVPacket **pkt);
void av_frame_free(AVFrame **frame);
void avcodec_free_context(AVCodecContext **avctx);
int av_image_alloc(uint8_t *pointers[4], int linesizes[4], int w, int h, enum AVPixelFormat pix_fmt, int align);

SwsContext *sws_getContext(int srcW, int srcH, enum AVPixelFormat srcFormat, int dstW, int dstH, enum AVPixelFormat dstFormat,
                           int flags, SwsFilter *srcFilter, SwsFilter *dstFilter, const double *param);
int sws_scale(SwsContext *c, const uint8_t *const srcSlice[], const int srcStride[], int srcSliceY, int srcSliceH,
              uint8_t *const dst[], const int dstStride[]);
void sws_freeContext(SwsContext *c);

/* ---- the fake decoder's input (set by the test driver before VideoDecoder::Init) ---- */
typedef struct fake_av_clip {
    int width, height, n_frames;
    const uint8_t *pict_is_p;        /* n_frames: 1 = P picture, 0 = I picture */
    const uint8_t *const *luma;      /* n_frames pointers to width*height planes (NULL entry: flat 128) */
    const uint8_t *const *side;      /* n_frames pointers to AVMotionVector arrays (NULL: no side data) */
    const int *side_bytes;           /* n_frames */
} fake_av_clip;
void fake_av_install(const fake_av_clip *clip);                    /* clip 0: url "fake://clip" */
void fake_av_install_clip(int id, const fake_av_clip *clip);      /* clip id: url "fake://clip/<id>" (one decoder per stream) */

#ifdef __cplusplus
}
#endif
