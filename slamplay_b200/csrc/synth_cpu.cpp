// synth_cpu.cpp — CPU renderer of the synthetic sequences (include/dmf_synth.h).
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (see slamplay_b200/build.py).
#include "../../include/dmf_synth.h"

extern "C" int dmf_synth_render_host(const dmf_synth_scene *scene, const dmf_synth_camera *cam, uint8_t *img,
                                     size_t step, double *dist, size_t dist_step) {
    if (!scene || !cam || !img || cam->width <= 0 || cam->height <= 0 || step < (size_t)cam->width) return -1;
    const dmf_synth_scene s = *scene;
    const dmf_synth_camera c = *cam;
#pragma omp parallel for schedule(dynamic, 4)
    for (int v = 0; v < c.height; ++v) {
        uint8_t *row = img + (size_t)v * step;
        double *drow = dist ? reinterpret_cast<double *>(reinterpret_cast<char *>(dist) + (size_t)v * dist_step) : nullptr;
        for (int u = 0; u < c.width; ++u) {
            row[u] = dmf_synth::shade_pixel(s, c, u, v);
            if (drow) drow[u] = dmf_synth::pixel_distance(s, c, u, v);
        }
    }
    return 0;
}
