// oracle/express.cc — TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of include/EXPRESS.h:20-192 on (image, stride, ROI) instead of cv::Mat.
// The uint8 wrap-around of the band limits, the p++-before-read off-by-one of the row scans and the
// (row=center_col, col=center_row) argument order of compute_center are reproduced, not fixed.
#include "oracle.h"

#include <cmath>
#include <cstring>

namespace {

// EXPRESS.h:20-38 tabulates, for the four supported shapes, the length (_L), start row (_S) and start column
// (_R[direction]) of every diagonal d = 0 .. rows+cols-2. The tables are this closed form (checked entry by
// entry against the header when the oracle was written):
//   _L[d]    = min(d+1, rows, cols, rows+cols-1-d)
//   _S[d]    = max(rows-1-d, 0)                       (walk starts on the left/right edge, then the top edge)
//   _R[1][d] = max(0, d-(rows-1))                     direction=1: from the bottom-left corner, step down-right
//   _R[0][d] = cols-1 - _R[1][d]                      direction=0: from the bottom-right corner, step down-left
inline int imin(int a, int b) { return a < b ? a : b; }
inline int imax(int a, int b) { return a > b ? a : b; }
inline int diag_len(int rows, int cols, int d) { return imin(imin(d + 1, rows), imin(cols, rows + cols - 1 - d)); }
inline int diag_row(int rows, int d) { return imax(rows - 1 - d, 0); }
inline int diag_col(int rows, int cols, int d, bool direction) {
    const int r1 = imax(0, d - (rows - 1));
    return direction ? r1 : cols - 1 - r1;
}

// A strided 1-D view, the result of diagonal() (EXPRESS.h:40-77).
struct Diag {
    const uint8_t *data;
    long step;
    int rows;
};

// EXPRESS.h:40-77. `roi` points at ROI(0,0); returns rows = 0 for unsupported shapes (the reference prints a
// warning and reads an uninitialised length there, i.e. undefined behaviour).
Diag diagonal(const uint8_t *roi, int stride, int rows, int cols, int d, bool direction) {
    Diag m = {roi, stride, 0};
    const bool supported = (rows == 8 || rows == 16) && (cols == 8 || cols == 16);  // :46-69
    if (supported) {
        m.rows = diag_len(rows, cols, d);
        m.data += (long)stride * diag_row(rows, d) + diag_col(rows, cols, d, direction);
    }
    m.step += direction ? 1 : -1;  // :72
    return m;
}

// EXPRESS.h:79-88. Note at(row = center_col, col = center_row).
uint8_t compute_center(const uint8_t *roi, int stride, int rows, int cols) {
    uint8_t center_row = rows / 2;
    uint8_t center_col = cols / 2;
    auto at = [&](int r, int c) -> int { return roi[(long)r * stride + c]; };
    return (at(center_col, center_row) + at(center_col - 1, center_row - 1) + at(center_col, center_row - 1) +
            at(center_col - 1, center_row)) /
           4;
}

}  // namespace

extern "C" int orc_express_center(const uint8_t *img, int stride, int x0, int y0, int cols, int rows) {
    return compute_center(img + (long)y0 * stride + x0, stride, rows, cols);
}

// EXPRESS.h:90-110
extern "C" void orc_express_descriptor(const uint8_t *img, int stride, int x0, int y0, int cols, int rows,
                                       int threshold, uint32_t desc[8]) {
    const uint8_t *roi = img + (long)y0 * stride + x0;
    uint8_t center = compute_center(roi, stride, rows, cols);
    uint8_t low_bounds = center - threshold;
    uint8_t high_bounds = center + threshold;

    std::memset(desc, 0, 32);  // desc.reset()

    for (int y = 0; y < rows; ++y) {
        const uint8_t *p = roi + (long)y * stride;  // img.ptr(y)
        for (int x = 0; x < cols; ++x) {
            p++;  // increment BEFORE the read: the tested pixel is (y, x+1)
            if ((low_bounds > *p) || (high_bounds < *p)) {
                const int bit = (y * rows) + x;  // desc.set((y * img.rows) + x, true)
                // bitset<256>::set throws beyond 255 (blocks larger than 16x16 never reach here in the reference)
                if (bit < 256) desc[bit >> 5] |= 1u << (bit & 31);
            }
        }
    }
}

// EXPRESS.h:112-115
extern "C" int orc_express_distance(const uint32_t a[8], const uint32_t b[8]) {
    int d = 0;
    for (int i = 0; i < 8; i++) d += __builtin_popcount(a[i] ^ b[i]);
    return d;
}

// EXPRESS.h:117-192
extern "C" int orc_express_test(const uint8_t *img, int stride, int x0, int y0, int cols, int rows,
                                int threshold) {
    const uint8_t *roi = img + (long)y0 * stride + x0;
    uint8_t center = compute_center(roi, stride, rows, cols);
    uint8_t low_bounds = center - threshold;
    uint8_t high_bounds = center + threshold;
    uint8_t precheck = (rows * cols * .125);

    uint8_t f = 0;
    for (int row = 0; row < rows; ++row) {
        const uint8_t *p = roi + (long)row * stride;
        for (int col = 0; col < cols; ++col) {
            p++;
            if (low_bounds > *p || high_bounds < *p) f++;
        }
        if (f >= precheck) break;
    }

    if (f < precheck) return 0;
    // Shapes other than 8x8/16x8/8x16/16x16 make diagonal() read an uninitialised length in the reference
    // (:66-69, undefined behaviour). Defined here as "not a feature block".
    if (!((rows == 8 || rows == 16) && (cols == 8 || cols == 16))) return 0;

    uint8_t slices = rows + cols - 1;
    uint8_t rounds = std::round(slices * .25);
    uint8_t u_rounds = slices - rounds;
    uint8_t wins, losses, win, loss;

    for (int a = 0; a < 2; a++) {
        wins = 0;
        losses = 0;
        for (int i = 0; i < slices; i++) {
            Diag diag = diagonal(roi, stride, rows, cols, i, a == 0);
            win = 0;
            loss = 0;
            for (int r = 0; r < diag.rows; r++) {
                const uint8_t v = diag.data[(long)r * diag.step];
                if (low_bounds > v || high_bounds < v) {
                    win++;
                } else {
                    loss++;
                }
            }
            if (wins < rounds) {
                if (win >= loss)
                    wins++;
                else
                    wins = 0;
            }

            if (losses < rounds) {
                if (loss > win)
                    losses++;
                else
                    losses = 0;
            }
            if (i > u_rounds && (wins == 0 || losses == 0)) break;
        }
        if (wins >= rounds && losses >= rounds) {
            return 1;
        }
    }
    return 0;
}
