// Renders one DrawTracker scene with include/vins_fast_b200/draw.hpp.
//   test_draw <scene.bin> <out.rgb>
// scene.bin: int32 w, h, n, n_right; u8 left[h*w]; u8 right[h*w]; f32 kps[2n]; i32 track_count[n]; f32 right_kps[2 n_right];
//            f32 prev_xy[2n]; u8 has_prev[n].   out.rgb: h x 2w x 3 bytes.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "vins_fast_b200/draw.hpp"

template <typename T>
static bool rd(FILE* f, std::vector<T>& v, size_t n) { v.resize(n ? n : 1); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) return 4;
    const int w = hdr[0], h = hdr[1], n = hdr[2], nr = hdr[3];
    std::vector<unsigned char> left, right, has_prev;
    std::vector<float> kps, rkps, prev;
    std::vector<int32_t> cnt;
    if (!rd(f, left, (size_t)w * h) || !rd(f, right, (size_t)w * h) || !rd(f, kps, 2 * (size_t)n) || !rd(f, cnt, (size_t)n) ||
        !rd(f, rkps, 2 * (size_t)nr) || !rd(f, prev, 2 * (size_t)n) || !rd(f, has_prev, (size_t)n)) return 5;
    fclose(f);
    std::vector<unsigned char> out((size_t)h * 2 * w * 3);
    const vf_draw::Canvas canvas = {out.data(), 2 * w, h, (size_t)2 * w * 3};
    vf_draw::DrawTrackImage(left.data(), (size_t)w, right.data(), (size_t)w, w, h, kps.data(), cnt.data(), n, rkps.data(), nr, prev.data(),
                            has_prev.data(), canvas);
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 6;
    fwrite(out.data(), 1, out.size(), o);
    fclose(o);
    return 0;
}
