// oracle/raster.cc — TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of the motion-vector loop of VideoDecoder::NextImage, src/VideoDecoder.cc:198-351.
// Follows the reference statement by statement: same float expressions, same truncations, same
// int-vs-float comparisons, same push_back order. Build with -ffp-contract=off.
#include "oracle.h"

#include <cstdlib>
#include <vector>

namespace {

// VideoImage / MotionVectorImage, include/Frame.h:109-156 (grey/RGB planes live outside the raster).
struct Frame {
    std::vector<int32_t>    mvi;  // cv::Mat(height,width,CV_32SC4, Scalar(-1,-1,-1,-1)), Frame.h:123
    std::vector<movfe_rect> kps;  // Frame.h:115
    std::vector<movfe_hop>  mvs;  // Frame.h:116
    double coverageArea = 0.0;
};

}  // namespace

struct orc_clip {
    int width = 0, height = 0;
    std::vector<Frame> frames;
    int64_t bad_ref = 0;
};

extern "C" orc_clip *orc_raster_clip(int width, int height, int n_frames, const movfe_mv_record *recs,
                                     const int64_t *rec_off, const uint8_t *frame_flags, int max_ref) {
    orc_clip *clip = new orc_clip;
    clip->width = width;
    clip->height = height;
    clip->frames.resize(n_frames);

    for (int n = 0; n < n_frames; n++) {
        // "vqueue" == frames[0..n): vqueue.size() == n, vqueue.back() == frames[n-1] (VideoDecoder.cc:163,353).
        Frame &smv = clip->frames[n];
        smv.mvi.assign((size_t)width * height * 4, -1);  // Frame.h:123
        smv.kps.reserve(3000);                           // Frame.h:124-125
        smv.mvs.reserve(3000);
        const int vq = n;

        if (!(frame_flags[n] & MOVFE_FRAME_MV)) continue;  // VideoDecoder.cc:200 "if (sd && mv)"

        const int num_mv = (int)(rec_off[n + 1] - rec_off[n]);  // :203
        float coverage = 0;                                      // :204
        float mb_h, mb_w, mb_h_half, mb_w_half, mv_x, mv_y, dst_x, dst_y, d_x_top, d_y_top, d_x_bottom,
            d_y_bottom, src_x, src_y, s_x_top, s_y_top, s_x_bottom, s_y_bottom;
        int sMB_size;
        int dIndx = -1;

        for (int i = 0; i < num_mv; i++) {  // :211
            const movfe_mv_record *mv = &recs[rec_off[n] + i];

            if (mv->ref > max_ref) {  // not in the reference: refs beyond the configured look-ahead are rejected
                clip->bad_ref++;
                continue;
            }

            mb_h = mv->h * 1;  // :215-218
            mb_w = mv->w * 1;
            mb_h_half = mb_h / 2;
            mb_w_half = mb_w / 2;

            mv_x = mv->dst_x - mv->src_x;  // :220-221 (int arithmetic, then int -> float)
            mv_y = mv->dst_y - mv->src_y;

            mv_x = mv_x / (mv->ref + 1);  // :223-224 (float / int)
            mv_y = mv_y / (mv->ref + 1);

            // Calculate dst mb (:227-228)
            dst_x = mv->ref > 0 && mv->source < 0 ? mv->src_x : mv->dst_x;
            dst_y = mv->ref > 0 && mv->source < 0 ? mv->src_y : mv->dst_y;

            d_x_top = dst_x - mb_w_half;  // :230-241
            if (d_x_top < 0) d_x_top = 0;
            d_y_top = dst_y - mb_h_half;
            if (d_y_top < 0) d_y_top = 0;
            d_x_bottom = dst_x + mb_w_half;
            if (d_x_bottom >= width) continue;
            d_y_bottom = dst_y + mb_h_half;
            if (d_y_bottom >= height) continue;

            dIndx = -1;  // :243
            // cv::Rect dMB(d_x_top, d_y_top, mb_w, mb_h): float -> int truncation (:244)
            movfe_rect dMB = {(int16_t)(int)d_x_top, (int16_t)(int)d_y_top, (int16_t)(int)mb_w, (int16_t)(int)mb_h};
            if (mv->ref > 0 && mv->source < 0) {  // :245-248
                const int q = (vq - 1) - mv->ref;
                if (q >= 0) clip->frames[q].kps.push_back(dMB);  // q < 0: deque under-run in the reference
            } else {  // :249-253
                smv.kps.push_back(dMB);
                dIndx = (int)smv.kps.size() - 1;
            }

            if (mv->source > 0) {
                // B frames (:255-286): entries go to bmap[frames], which nothing reads. No hop, no coverage.
            } else {  // P frames (:287-348)
                for (int j = (mv->ref + 1); j > 0; j--) {
                    src_x = mv->dst_x + (mv_x * j * -1);  // :291-292
                    src_y = mv->dst_y + (mv_y * j * -1);

                    // Calculate src mb (:295-306)
                    s_x_top = src_x - mb_w_half;
                    if (s_x_top < 0) s_x_top = 0;
                    s_y_top = src_y - mb_h_half;
                    if (s_y_top < 0) s_y_top = 0;
                    s_x_bottom = src_x + mb_w_half;
                    if (s_x_bottom >= width) s_x_bottom = width - 1;
                    s_y_bottom = src_y + mb_h_half;
                    if (s_y_bottom >= height) s_y_bottom = height - 1;

                    movfe_hop mvc;  // :310-313
                    mvc.mv_x = mv_x;
                    mvc.mv_y = mv_y;
                    mvc.d_indx = dIndx;
                    mvc._pad = 0;

                    Frame *sp;  // :315-323
                    if (j == 1) {
                        sp = &smv;
                    } else {
                        const int q = vq - (j - 1);
                        if (q < 0) continue;  // deque under-run in the reference
                        sp = &clip->frames[q];
                    }

                    sp->mvs.push_back(mvc);  // :325
                    sMB_size = (int)sp->mvs.size() - 1;

                    // Used for forward predicting (:330-345). int h vs float bound, inclusive.
                    for (int h = s_y_top; h <= s_y_bottom; h++) {
                        int32_t *p = &sp->mvi[(size_t)h * width * 4];
                        for (int w = s_x_top; w <= s_x_bottom; w++) {
                            int32_t *v = p + (size_t)w * 4;
                            if (v[0] == -1)
                                v[0] = sMB_size;
                            else if (v[1] == -1)
                                v[1] = sMB_size;
                            else if (v[2] == -1)
                                v[2] = sMB_size;
                            else
                                v[3] = sMB_size;
                        }
                    }
                }
                coverage += (int)dMB.w * (int)dMB.h;  // :347 dMB.area()
            }
        }
        smv.coverageArea = coverage / (double)(width * height);  // :350
    }
    return clip;
}

extern "C" void orc_clip_free(orc_clip *c) { delete c; }
extern "C" int orc_clip_n_hops(const orc_clip *c, int f) { return (int)c->frames[f].mvs.size(); }
extern "C" int orc_clip_n_kps(const orc_clip *c, int f) { return (int)c->frames[f].kps.size(); }
extern "C" double orc_clip_coverage(const orc_clip *c, int f) { return c->frames[f].coverageArea; }
extern "C" int64_t orc_clip_bad_ref(const orc_clip *c) { return c->bad_ref; }
extern "C" const int32_t *orc_clip_grid(const orc_clip *c, int f) { return c->frames[f].mvi.data(); }
extern "C" const movfe_hop *orc_clip_hops(const orc_clip *c, int f) { return c->frames[f].mvs.data(); }
extern "C" const movfe_rect *orc_clip_kps(const orc_clip *c, int f) { return c->frames[f].kps.data(); }
