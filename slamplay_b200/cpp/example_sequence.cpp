// example_sequence.cpp — the reference's driver loop (dense_mapping/test_monocular_mapping.cpp:264-305)
// on a synthetic sequence, through the C++ shim.  Build:
//   g++ -std=c++17 -O2 example_sequence.cpp -o example_sequence -L.. -ldmf -ldmf_synth_cpu -Wl,-rpath,'$ORIGIN/..'
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../include/dmf_synth.h"
#include "dense_mono_update.hpp"

using slamplay_b200::Mat;
using slamplay_b200::SE3d;

int main(int argc, char **argv) {
    const int width = 640, height = 480, n_frames = argc > 1 ? std::atoi(argv[1]) : 10;
    dmf_params p;
    dmf_default_params(&p, width, height, 0);
    dmf_synth_scene scene{2.0, 0.15, 1.0, 1.5 * 2.0 / std::fabs(p.fx), 5, 0x5EED0000u, 12, 2};
    auto render = [&](double tx, std::vector<uint8_t> &img) {
        dmf_synth_camera cam{width, height, p.fx, p.fy, p.cx, p.cy, {0, 0, 0, 1}, {tx, 0, 0}};
        img.resize(size_t(width) * height);
        dmf_synth_render_host(&scene, &cam, img.data(), width, nullptr, 0);
    };
    std::vector<uint8_t> ref_img, cur_img;
    render(0.0, ref_img);
    Mat ref(height, width, slamplay_b200::kType8UC1, ref_img.data(), width);

    std::vector<double> depth_buf(size_t(width) * height, 3.0), cov_buf(size_t(width) * height, 3.0);  // :270-278
    Mat depth(height, width, slamplay_b200::kType64F, depth_buf.data(), width * sizeof(double));
    Mat depth_cov2(height, width, slamplay_b200::kType64F, cov_buf.data(), width * sizeof(double));

    for (int index = 1; index < n_frames; index++) {  // :285
        const double tx = 0.004 * index;              // T_WC(index) = translation along x
        render(tx, cur_img);
        Mat curr(height, width, slamplay_b200::kType8UC1, cur_img.data(), width);
        SE3d T_C_R;                                   // T_WC(index)^-1 * T_WC(0): identity rotation, t = -tx
        T_C_R.t[0] = -tx;
        slamplay_b200::update(ref, curr, T_C_R, depth, depth_cov2);  // :291
        double s = 0; long n = 0;
        for (int y = p.border; y < height - p.border; y++)
            for (int x = p.border; x < width - p.border; x++)
                if (cov_buf[size_t(y) * width + x] < 0.5) { s += depth_buf[size_t(y) * width + x]; n++; }
        std::printf("*** loop %d ***  pixels with cov2 < 0.5: %ld, mean depth %.4f\n", index, n, n ? s / n : 0.0);
    }
    return 0;
}
