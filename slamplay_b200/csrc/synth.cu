// synth.cu — CUDA renderer of the synthetic sequences (include/dmf_synth.h).
// Compiled with -fmad=false so it is bit-identical with synth_cpu.cpp.
#include "../../include/dmf_synth.h"
#include <cuda_runtime.h>

namespace {
__global__ void __launch_bounds__(256) render_kernel(const dmf_synth_scene s, const dmf_synth_camera c, uint8_t *img,
                                                     size_t pitch, double *dist, size_t dist_pitch) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    int v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= c.width || v >= c.height) return;
    img[(size_t)v * pitch + u] = dmf_synth::shade_pixel(s, c, u, v);
    if (dist) *reinterpret_cast<double *>(reinterpret_cast<char *>(dist) + (size_t)v * dist_pitch + (size_t)u * 8) =
        dmf_synth::pixel_distance(s, c, u, v);
}
}  // namespace

extern "C" int dmf_synth_render_device(const dmf_synth_scene *scene, const dmf_synth_camera *cam, uint8_t *img_dev,
                                       size_t pitch, double *dist_dev, size_t dist_pitch, void *stream) {
    if (!scene || !cam || !img_dev || cam->width <= 0 || cam->height <= 0 || pitch < (size_t)cam->width) return -1;
    dim3 blk(32, 8), grid((cam->width + 31) / 32, (cam->height + 7) / 8);
    render_kernel<<<grid, blk, 0, (cudaStream_t)stream>>>(*scene, *cam, img_dev, pitch, dist_dev, dist_pitch);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
