// VideoDecoder_movfe.cc — drop-in for src/VideoDecoder.cc. Demux, decode and colour conversion stay on host libavcodec
// (H.264 with the repo's ffmpeg-ref-patch exporting AVMotionVector::ref); what is replaced is the MV loop of NextImage
// (src/VideoDecoder.cc:198-351): every decoded frame's side data goes to the GPU raster (movfe_shim::RasterQueue ->
// movfe_push_frames / movfe_raster) and the VideoImage a caller receives carries mvs / kps / mvi / coverageArea bit-identical
// to the reference's. The look-ahead is the reference's: a frame is handed out once `qlen` frames are buffered
// (VideoDecoder.cc:163,363-368), by which time every later frame that can back-fill hops into it (ref <= qlen-2) was pushed.
#include <cstdio>
#include <iostream>
#include <map>

#ifdef MOVFE_IN_TREE
#include "VideoDecoder.h"
#else
#include "VideoDecoder_movfe.h"
#endif
#include "movfe_shim.h"

namespace {
// one RasterQueue per decoder object; the reference's class has no member to spare, so the queues live beside it
std::map<const void *, movfe_shim::RasterQueue *> &queues() {
    static std::map<const void *, movfe_shim::RasterQueue *> q;
    return q;
}
}  // namespace

namespace MOV_SLAM {

VideoDecoder::VideoDecoder(const std::string &path, int qlen) : dataset_path(path), video_stream_index(-1), qlen(qlen), frames(0) {}

VideoDecoder::~VideoDecoder(void) {
    auto it = queues().find(this);
    if (it != queues().end()) {
        delete it->second;
        queues().erase(it);
    }
    if (pFormatContext) avformat_close_input(&pFormatContext);
    if (pPacket) av_packet_free(&pPacket);
    if (pFrame) av_frame_free(&pFrame);
    if (pCodecContext) avcodec_free_context(&pCodecContext);
    if (conversion_rgb) sws_freeContext(conversion_rgb);
    if (conversion_grey) sws_freeContext(conversion_grey);
}

// Opens the stream with motion-vector export switched on and picks the first video stream, as src/VideoDecoder.cc:37-149.
bool VideoDecoder::Init() {
    auto fail = [](const char *what) {
        std::cerr << "ERROR " << what << std::endl;
        return false;
    };
    avdevice_register_all();
    if (!(pFormatContext = avformat_alloc_context())) return fail("could not allocate memory for Format Context");
    inputFormat = av_find_input_format("libx264");
    AVDictionary *options = NULL;
    av_dict_set(&options, "flags2", "+export_mvs", 0);  // side data = AVMotionVector records
    if (avformat_open_input(&pFormatContext, dataset_path.c_str(), inputFormat, &options) != 0) return fail("could not open the file");
    if (avformat_find_stream_info(pFormatContext, NULL) < 0) return fail("could not get the stream info");
    for (unsigned int i = 0; i < pFormatContext->nb_streams && video_stream_index < 0; i++) {
        AVCodecParameters *par = pFormatContext->streams[i]->codecpar;
        AVCodec *codec = avcodec_find_decoder(par->codec_id);
        if (codec && par->codec_type == AVMEDIA_TYPE_VIDEO) {
            video_stream_index = (int)i;
            pCodec = codec;
            pCodecParameters = par;
        }
    }
    if (video_stream_index < 0) return fail("file does not contain a video stream");
    if (!(pCodecContext = avcodec_alloc_context3(pCodec))) return fail("failed to allocate AVCodecContext");
    if (avcodec_parameters_to_context(pCodecContext, pCodecParameters) < 0) return fail("failed to copy codec params to codec context");
    if (avcodec_open2(pCodecContext, pCodec, &options) < 0) return fail("failed to open codec through avcodec_open2");
    if (!(pFrame = av_frame_alloc()) || !(pPacket = av_packet_alloc())) return fail("failed to allocate AVFrame / AVPacket");
    const int w = pCodecContext->width, h = pCodecContext->height;
    conversion_rgb = sws_getContext(w, h, pCodecContext->pix_fmt, w, h, AV_PIX_FMT_BGR24, SWS_FAST_BILINEAR, NULL, NULL, NULL);
    conversion_grey = sws_getContext(w, h, pCodecContext->pix_fmt, w, h, AV_PIX_FMT_GRAY8, SWS_FAST_BILINEAR, NULL, NULL, NULL);
    mFPS = (float)av_q2d(pFormatContext->streams[video_stream_index]->r_frame_rate);
    // the reference's deque lets a record reach back qlen-1 frames; ref <= qlen-2 keeps its indices in range
    queues()[this] = new movfe_shim::RasterQueue(w, h, std::max(0, std::min(qlen - 2, 10)));
    return true;
}

int VideoDecoder::GetWidth() { return pCodecContext->width; }
int VideoDecoder::GetHeight() { return pCodecContext->height; }

shared_ptr<MotionVectorImage> VideoDecoder::NextImage(bool mv) {
    movfe_shim::RasterQueue *rq = queues()[this];
    if (!rq) return nullptr;
    bool eof = false;
    // fill the look-ahead: decode until qlen frames are pending (VideoDecoder.cc:163)
    while (rq->pending() < qlen && !eof) {
        if (av_read_frame(pFormatContext, pPacket) < 0) {
            eof = true;
            break;
        }
        bool got = false;
        while (!got) {  // a packet may need to be re-sent until the decoder hands out a picture (:166-178)
            const int sent = avcodec_send_packet(pCodecContext, pPacket);
            if (sent < 0) {
                std::cerr << "Error while sending a packet to the decoder " << std::endl;
                eof = true;
                break;
            }
            if (pPacket->stream_index != video_stream_index) break;
            const int rc = avcodec_receive_frame(pCodecContext, pFrame);
            if (rc == AVERROR(EAGAIN) || rc == AVERROR_EOF) continue;
            if (rc < 0) {
                std::cout << "Error while receiving a frame from the decoder " << std::endl;
                eof = true;
                break;
            }
            got = true;
        }
        if (got) {
            frames++;
            shared_ptr<MotionVectorImage> smv(new MotionVectorImage(pFrame->width, pFrame->height));
            smv->frame = frames;
            smv->ft = pFrame->pict_type != AV_PICTURE_TYPE_I ? FrameType::P_FRAME : FrameType::I_FRAME;  // :193
            smv->coverageArea = 0.0;
            // colour conversion (VideoBase.h:50-68): the luma plane becomes imGray, BGR is kept for the viewer
            smv->imGray = cv::Mat(pFrame->height, pFrame->width, CV_8UC1);
            int ls[1] = {(int)(size_t)smv->imGray.step};
            sws_scale(conversion_grey, pFrame->data, pFrame->linesize, 0, pFrame->height, &smv->imGray.data, ls);
            smv->imRGB = cv::Mat(pFrame->height, pFrame->width, CV_8UC3);
            int lc[1] = {(int)(size_t)smv->imRGB.step};
            sws_scale(conversion_rgb, pFrame->data, pFrame->linesize, 0, pFrame->height, &smv->imRGB.data, lc);
            // the side data goes to the GPU instead of through the MV loop (:198-351)
            AVFrameSideData *sd = av_frame_get_side_data(pFrame, AV_FRAME_DATA_MOTION_VECTORS);
            const int n = (sd && mv) ? (int)(sd->size / sizeof(AVMotionVector)) : 0;
            if (!rq->push(smv, n ? sd->data : nullptr, n, sd && mv)) return nullptr;
        }
        av_packet_unref(pPacket);
    }
    // hand out the oldest pending frame; at end of stream no more look-ahead will come (:362-369)
    return rq->pop(eof || rq->pending() >= qlen);
}

}  // namespace MOV_SLAM
