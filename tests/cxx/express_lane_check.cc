// CPU check of mov-slam_b200/csrc/express_lane.cuh (the thread-level EXPRESS used by the propagation kernels) against the oracle's
// restatement of include/EXPRESS.h. Built and run by tests/test_express_lane.py; prints "OK <cases> <passes>" or the first mismatch.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mov-slam_b200/csrc/express_lane.cuh"

extern "C" {
void orc_express_descriptor(const uint8_t *img, int stride, int x0, int y0, int cols, int rows, int threshold, uint32_t desc[8]);
int orc_express_test(const uint8_t *img, int stride, int x0, int y0, int cols, int rows, int threshold);
int orc_express_center(const uint8_t *img, int stride, int x0, int y0, int cols, int rows);
}

static uint64_t rng_state = 0x5EED1234ABCDull;
static uint64_t rnd() {  // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void stage(const uint8_t *img, int pitch, int xw, int y0, uint32_t *win) {
    for (int r = 0; r < 16; r++) std::memcpy(win + r * xl::ROW_WORDS, img + (size_t)(y0 + r) * pitch + xw, xl::ROW_BYTES);
}

int main(int argc, char **argv) {
    const int n_cases = argc > 1 ? std::atoi(argv[1]) : 200000;
    const int W = 256, H = 192, P = 320;
    std::vector<uint8_t> img((size_t)P * (H + 16) + 64);
    long cases = 0, passes = 0, alls = 0;
    for (int im = 0; im < 6; im++) {
        // textures of different kinds: blobs, edges, noise, near-black and near-white (band limits that wrap)
        for (int y = 0; y < H + 16; y++)
            for (int x = 0; x < P; x++) {
                int v;
                const int blob = (((x / 5) * 7 + (y / 3) * 13) ^ ((x / 11) * (y / 7))) & 0xff;
                switch (im) {
                    case 0: v = blob; break;
                    case 1: v = ((x + y) & 16) ? 200 : 40; v += (int)(rnd() % 9) - 4; break;
                    case 2: v = (int)(rnd() & 0xff); break;
                    case 3: v = (int)(rnd() % 24); break;
                    case 4: v = 255 - (int)(rnd() % 24); break;
                    default: v = (x * 3 + y * 2) & 0xff; v = (v + (blob >> 2)) & 0xff; break;
                }
                img[(size_t)y * P + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        for (int it = 0; it < n_cases / 6; it++) {
            const int cols = (rnd() & 1) ? 16 : 8, rows = (rnd() & 1) ? 16 : 8;
            const int mx = (int)(rnd() % (W - cols - 1)), my = (int)(rnd() % (H - rows - 1));
            const int thr = (it % 7 == 0) ? (int)(rnd() % 128) : 5 + (int)(rnd() % 40);
            uint32_t win[xl::WIN_WORDS];
            uint32_t want[8], got[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            // candidate descriptor: window from column mx + 1
            {
                const int xw = (mx + 1) & ~7;
                stage(img.data(), P, xw, my, win);
                const int c = xl::centre_of(win, mx - xw, rows, cols);
                if (c != orc_express_center(img.data(), P, mx, my, cols, rows)) { std::printf("centre mismatch\n"); return 1; }
                const xl::Band b = xl::band_of(c, thr);
                if (b.k4 == 0x80808080u) alls++;
                for (int half = 0; half < (rows == 16 ? 2 : 1); half++) {
                    uint32_t d[4];
                    xl::half_descriptor(win, mx + 1 - xw, rows, cols, half, b, d);
                    for (int k = 0; k < 4; k++) got[4 * half + k] = d[k];
                }
                orc_express_descriptor(img.data(), P, mx, my, cols, rows, thr, want);
                if (std::memcmp(want, got, 32)) {
                    std::printf("descriptor mismatch im=%d mx=%d my=%d %dx%d thr=%d\n", im, mx, my, cols, rows, thr);
                    return 1;
                }
            }
            // compute_express + descriptor: window from column mx
            {
                const int xw = mx & ~7;
                stage(img.data(), P, xw, my, win);
                uint32_t d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const bool ok = xl::block_express(win, mx - xw, rows, cols, thr, d);
                const int ref = orc_express_test(img.data(), P, mx, my, cols, rows, thr);
                if ((int)ok != ref) {
                    std::printf("compute_express mismatch im=%d mx=%d my=%d %dx%d thr=%d got %d want %d\n", im, mx, my, cols, rows, thr, (int)ok, ref);
                    return 1;
                }
                if (ok && std::memcmp(want, d, 32)) {
                    std::printf("birth descriptor mismatch im=%d mx=%d my=%d %dx%d thr=%d\n", im, mx, my, cols, rows, thr);
                    return 1;
                }
                passes += ok;
            }
            cases++;
        }
    }
    std::printf("OK %ld %ld %ld\n", cases, passes, alls);
    return 0;
}
