// test_pool.cc — CPU check of the decoder pool and the trajectory writers (no GPU: nothing is pushed). Streams are decoded from the
// fake libav back-end (one clip per stream, "fake://clip/<id>"); the windows the pool assembles on its worker threads must equal a
// plain serial packing of the same clips, byte for byte.  usage: test_pool <clip dir> <n streams> <frames per window> <threads>
// The clip directory holds recs.bin (40-byte records), off.bin (int64, n+1), flags.bin (uint8, n), grey.bin (n planes), meta.txt "W H n".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "decoder_pool.h"
extern "C" {
#include "standin/libav_standin.h"
}

template <typename T>
static std::vector<T> slurp(const std::string &p) {
    std::ifstream f(p, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> v(raw.size() / sizeof(T));
    memcpy(v.data(), raw.data(), v.size() * sizeof(T));
    return v;
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const std::string dir = argv[1];
    const int S = atoi(argv[2]), F = atoi(argv[3]), NT = atoi(argv[4]);
    int W = 0, H = 0, N = 0;
    {
        std::ifstream m(dir + "/meta.txt");
        m >> W >> H >> N;
    }
    const auto recs = slurp<movfe_mv_record>(dir + "/recs.bin");
    const auto off = slurp<int64_t>(dir + "/off.bin");
    const auto flags = slurp<uint8_t>(dir + "/flags.bin");
    const auto grey = slurp<uint8_t>(dir + "/grey.bin");
    const size_t plane = (size_t)W * H;
    // stream s plays the clip from frame s on (so the streams differ) and is N - S + 1 frames long ... except the last, one shorter
    std::vector<std::vector<const uint8_t *>> luma(S), side(S);
    std::vector<std::vector<int>> side_bytes(S);
    std::vector<std::vector<uint8_t>> isp(S);
    std::vector<int> len(S);
    for (int s = 0; s < S; s++) {
        len[s] = N - S + 1 - (s == S - 1 ? 1 : 0);
        for (int f = 0; f < len[s]; f++) {
            const int g = s + f;
            luma[s].push_back(grey.data() + (size_t)g * plane);
            side[s].push_back(off[g + 1] > off[g] ? reinterpret_cast<const uint8_t *>(recs.data() + off[g]) : nullptr);
            side_bytes[s].push_back((int)((off[g + 1] - off[g]) * (int64_t)sizeof(movfe_mv_record)));
            isp[s].push_back((flags[g] & MOVFE_FRAME_P) ? 1 : 0);
        }
        fake_av_clip c = {W, H, len[s], isp[s].data(), luma[s].data(), side[s].data(), side_bytes[s].data()};
        fake_av_install_clip(s, &c);
    }
    std::vector<std::string> paths;
    for (int s = 0; s < S; s++) paths.push_back("fake://clip/" + std::to_string(s));
    movfe_shim::DecoderPool pool(paths, NT);
    if (!pool.ok() || pool.width() != W || pool.height() != H) {
        printf("FAIL open\n");
        return 1;
    }
    int done = 0, windows = 0;
    long total_recs = 0;
    movfe_shim::HostWindow w;
    for (;;) {
        const int n = pool.next_window(F, w);
        if (n == 0) break;
        windows++;
        // serial packing of the same frames
        if ((int)w.flags.size() != S * n || (int)w.rec_off.size() != S * n + 1 || w.grey.size() != (size_t)S * n * plane) {
            printf("FAIL sizes\n");
            return 1;
        }
        size_t at = 0;
        for (int s = 0; s < S; s++)
            for (int f = 0; f < n; f++) {
                const int g = s + done + f;
                const int64_t cnt = off[g + 1] - off[g];
                if (w.rec_off[(size_t)s * n + f] != (int64_t)at || w.rec_off[(size_t)s * n + f + 1] != (int64_t)at + cnt) {
                    printf("FAIL offsets s=%d f=%d\n", s, f);
                    return 1;
                }
                std::vector<movfe_packed_record> want((size_t)cnt);
                movfe_pack_records(recs.data() + off[g], cnt, want.data());
                if (cnt && memcmp(want.data(), w.recs.data() + at, (size_t)cnt * sizeof(movfe_packed_record))) {
                    printf("FAIL records s=%d f=%d\n", s, f);
                    return 1;
                }
                const uint8_t fl = (uint8_t)((flags[g] & MOVFE_FRAME_P) | (cnt > 0 ? MOVFE_FRAME_MV : 0));
                if (w.flags[(size_t)s * n + f] != fl) {
                    printf("FAIL flags s=%d f=%d got %u want %u\n", s, f, w.flags[(size_t)s * n + f], fl);
                    return 1;
                }
                if (memcmp(w.grey.data() + ((size_t)s * n + f) * plane, grey.data() + (size_t)g * plane, plane)) {
                    printf("FAIL luma s=%d f=%d\n", s, f);
                    return 1;
                }
                at += (size_t)cnt;
                total_recs += cnt;
            }
        done += n;
    }
    // all streams advance in step: the shortest stream (the last one) ends the run
    if (done != len[S - 1]) {
        printf("FAIL frames %d want %d\n", done, len[S - 1]);
        return 1;
    }
    // trajectories: two poses with known inverses
    movfe_pose P[3];
    memset(P, 0, sizeof P);
    for (int i = 0; i < 3; i++) P[i].R[0] = P[i].R[4] = P[i].R[8] = 1.0;
    P[1].R[0] = 0, P[1].R[1] = -1, P[1].R[3] = 1, P[1].R[4] = 0;  // 90 degrees about z
    P[1].t[0] = 1, P[1].t[1] = 2, P[1].t[2] = 3;
    P[2].t[0] = -0.5;
    const double ts[3] = {0.0, 0.033333, 0.066667};
    const uint8_t lost[3] = {0, 0, 1};
    if (!movfe_shim::write_trajectory_tum(dir + "/traj_tum.txt", ts, P, lost, 3) || !movfe_shim::write_trajectory_kitti(dir + "/traj_kitti.txt", nullptr, P, lost, 3)) {
        printf("FAIL trajectory files\n");
        return 1;
    }
    printf("OK %d windows %d frames %ld records\n", windows, done, total_recs);
    return 0;
}
