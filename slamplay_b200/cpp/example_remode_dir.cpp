// example_remode_dir.cpp — the reference's whole driver (dense_mapping/test_monocular_mapping.cpp:253-310) on a
// REMODE-layout directory, through the C++ shim: readDatasetFiles (ref:256), reference image (ref:264), state init
// (ref:270-278), per frame T_C_R = T_WC(i)^-1 * T_WC(0) (ref:289-290), update() (ref:291) and evaludateDepth (ref:292).
// Images are binary PGM (P5) here: the reference decodes with cv::imread, which this stand-alone example does not link.
//
//   g++ -std=c++17 -O2 example_remode_dir.cpp -o example_remode_dir -L.. -ldmf -Wl,-rpath,'$ORIGIN/..'
//   ./example_remode_dir <dataset dir> [--dump-poses]
#include <cmath>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "dense_mono_update.hpp"

using slamplay_b200::Mat;
using slamplay_b200::SE3d;

static bool read_pgm(const std::string &file, int width, int height, std::vector<uint8_t> &img) {
    std::ifstream f(file, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    if (!(f >> magic >> w >> h >> maxv) || magic != "P5" || w != width || h != height || maxv != 255) return false;
    f.get();
    img.resize(size_t(w) * h);
    f.read(reinterpret_cast<char *>(img.data()), std::streamsize(img.size()));
    return bool(f);
}

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <dataset dir> [--dump-poses]\n", argv[0]); return 2; }
    const int width = 640, height = 480;  // ref:73-74
    std::vector<std::string> files;
    std::vector<SE3d> poses_TWC;
    Mat ref_depth;
    if (!slamplay_b200::readDatasetFiles(argv[1], files, poses_TWC, ref_depth, width, height)) {  // ref:256
        std::printf("Reading image files failed!\n");
        return 1;
    }
    std::printf("read total %zu files.\n", files.size());
    if (argc > 2 && std::string(argv[2]) == "--dump-poses") {  // what the reader and the pose chain produce, bit-exact (hex floats)
        for (size_t i = 0; i < poses_TWC.size(); i++) {
            double q[4], t[3], rq[4], rt[3];
            slamplay_b200::pose_of(poses_TWC[i], q, t);
            slamplay_b200::pose_of(slamplay_b200::relative_pose(poses_TWC[0], poses_TWC[i]), rq, rt);
            std::printf("pose %zu %a %a %a %a %a %a %a | T_C_R %a %a %a %a %a %a %a\n", i, q[0], q[1], q[2], q[3], t[0], t[1], t[2], rq[0],
                        rq[1], rq[2], rq[3], rt[0], rt[1], rt[2]);
        }
        double s = 0;
        for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) s += ref_depth.ptr<double>(y)[x];
        std::printf("ref_depth sum %a first %a last %a\n", s, ref_depth.ptr<double>(0)[0], ref_depth.ptr<double>(height - 1)[width - 1]);
        return 0;
    }
    std::vector<uint8_t> ref_img, cur_img;
    if (!read_pgm(files[0], width, height, ref_img)) { std::printf("cannot read %s\n", files[0].c_str()); return 1; }
    Mat ref(height, width, slamplay_b200::kType8UC1, ref_img.data(), width);                      // ref:264
    const double init_depth = 3.0, init_cov2 = 3.0;                                               // ref:270,274
    Mat depth = slamplay_b200::make_mat64(height, width), depth_cov2 = slamplay_b200::make_mat64(height, width);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) { depth.ptr<double>(y)[x] = init_depth; depth_cov2.ptr<double>(y)[x] = init_cov2; }  // ref:277-278
    const double good_cov = 2.0 * 0.01 * 0.01;                                                    // ref:85-89
    for (size_t index = 1; index < files.size(); index++) {                                       // ref:285
        std::printf("*** loop %zu ***\n", index);
        if (!read_pgm(files[index], width, height, cur_img)) continue;                            // ref:288
        Mat curr(height, width, slamplay_b200::kType8UC1, cur_img.data(), width);
        const SE3d T_C_R = slamplay_b200::relative_pose(poses_TWC[0], poses_TWC[index]);          // ref:289-290
        slamplay_b200::update(ref, curr, T_C_R, depth, depth_cov2);                               // ref:291
        // evaludateDepth(ref_depth, depth, depth_cov2, good_cov) ref:292,569-590 on the host maps update() returned
        double sq = 0; long cnt = 0;
        for (int y = 20; y < height - 20; y++)
            for (int x = 20; x < width - 20; x++) {
                if (depth_cov2.ptr<double>(y)[x] >= good_cov) continue;
                const double e = ref_depth.ptr<double>(y)[x] - depth.ptr<double>(y)[x];
                sq += e * e; cnt++;
            }
        std::printf("Average error (RMS) = %.10g over %ld pixels\n", cnt ? std::sqrt(sq / cnt) : 0.0, cnt);
    }
    std::printf("estimation returns\n");
    return 0;
}
