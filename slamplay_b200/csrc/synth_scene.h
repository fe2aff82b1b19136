// synth_scene.h — deterministic synthetic textured-relief sequences (SURVEY.md §8d
// "Synthetic inputs").  Shared by the CPU renderer (synth_cpu.cpp, used by the CPU test
// suite and the golden-fixture script) and the CUDA renderer (synth.cu, used by bench.py
// for the 1080p / 4K sequences).  The reference ships no data for this path (REMODE is a
// network download, scripts/download_dataset_remode_test_data.sh:18), so every benchmark
// and parity input comes from here.
//
// Scene: a relief surface  Z = plane_z + relief_amp * (2*vnoise(X/L, Y/L) - 1)  in the world
// frame, carrying a band-limited value-noise texture.  A pinhole camera with pose T_WC
// (unit quaternion + translation, camera looks along +Z_c) ray-casts it by fixed-point
// iteration.  Only +,-,*,/ and floor are used and both builds disable FMA contraction
// (g++ -ffp-contract=off, nvcc -fmad=false), so the CPU and GPU renderers are bit-identical.
#ifndef DMF_SYNTH_SCENE_H_
#define DMF_SYNTH_SCENE_H_

#include <stdint.h>

#ifdef __CUDACC__
#define DMF_HD __host__ __device__ __forceinline__
#else
#define DMF_HD inline
#endif

typedef struct dmf_synth_scene {
    double plane_z;        /* mean distance of the surface along world +Z [m] */
    double relief_amp;     /* relief amplitude [m] */
    double relief_period;  /* relief lattice spacing [m] */
    double tex_base;       /* finest texture lattice spacing [m] (≈1.5 image px on the surface) */
    int32_t tex_octaves;   /* number of texture octaves (spacing doubles per octave), <= 8 */
    uint32_t seed;
    int32_t ray_iters;     /* fixed-point iterations of the ray cast */
    int32_t supersample;   /* n x n sub-samples per pixel (1 or 2) */
} dmf_synth_scene;

typedef struct dmf_synth_camera {
    int32_t width, height;
    double fx, fy, cx, cy;
    double q[4];  /* T_WC rotation, unit quaternion (x,y,z,w) */
    double t[3];  /* T_WC translation */
} dmf_synth_camera;

namespace dmf_synth {

DMF_HD uint32_t hash3(uint32_t seed, int32_t ix, int32_t iy, int32_t k) {
    uint32_t h = seed ^ 0x9E3779B9u;
    h ^= (uint32_t)ix * 0x85EBCA77u; h = (h << 13) | (h >> 19); h *= 0xC2B2AE3Du;
    h ^= (uint32_t)iy * 0x27D4EB2Fu; h = (h << 15) | (h >> 17); h *= 0x165667B1u;
    h ^= (uint32_t)k * 0x9E3779B1u;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

DMF_HD double lattice(uint32_t seed, int32_t ix, int32_t iy, int32_t k) {
    return (double)(hash3(seed, ix, iy, k) >> 8) * (1.0 / 16777216.0);
}

DMF_HD double dfloor(double x) {
#ifdef __CUDA_ARCH__
    return floor(x);
#else
    return __builtin_floor(x);
#endif
}

// Quintic-faded value noise in [0,1).
DMF_HD double vnoise(uint32_t seed, double x, double y, int32_t k) {
    double fx0 = dfloor(x), fy0 = dfloor(y);
    int32_t ix = (int32_t)fx0, iy = (int32_t)fy0;
    double tx = x - fx0, ty = y - fy0;
    double sx = tx * tx * tx * (tx * (tx * 6.0 - 15.0) + 10.0);
    double sy = ty * ty * ty * (ty * (ty * 6.0 - 15.0) + 10.0);
    double v00 = lattice(seed, ix, iy, k), v10 = lattice(seed, ix + 1, iy, k);
    double v01 = lattice(seed, ix, iy + 1, k), v11 = lattice(seed, ix + 1, iy + 1, k);
    double a = v00 + sx * (v10 - v00);
    double b = v01 + sx * (v11 - v01);
    return a + sy * (b - a);
}

DMF_HD double relief(const dmf_synth_scene &s, double X, double Y) {
    return s.plane_z + s.relief_amp * (2.0 * vnoise(s.seed, X / s.relief_period, Y / s.relief_period, 100) - 1.0);
}

// Texture intensity in [0,1].
DMF_HD double texture(const dmf_synth_scene &s, double X, double Y) {
    double v = 0.5;
    double spacing = s.tex_base;
    for (int k = 0; k < s.tex_octaves; ++k) {
        double amp = (k < 2) ? 0.20 : ((k < 4) ? 0.15 : 0.10);
        v += amp * (2.0 * vnoise(s.seed, X / spacing, Y / spacing, k) - 1.0);
        spacing = spacing * 2.0;
    }
    if (v < 0.0) v = 0.0;
    if (v > 1.0) v = 1.0;
    return v;
}

struct Ray { double ox, oy, oz, dx, dy, dz; };

// World-frame ray through pixel (u,v): direction R_WC * ((u-cx)/fx, (v-cy)/fy, 1).
DMF_HD Ray pixel_ray(const dmf_synth_camera &c, double u, double v) {
    double x = (u - c.cx) / c.fx, y = (v - c.cy) / c.fy, z = 1.0;
    double qx = c.q[0], qy = c.q[1], qz = c.q[2], qw = c.q[3];
    // v + w*(2 q×v) + q×(2 q×v)
    double ux = qy * z - qz * y, uy = qz * x - qx * z, uz = qx * y - qy * x;
    ux = ux + ux; uy = uy + uy; uz = uz + uz;
    Ray r;
    r.dx = x + qw * ux + (qy * uz - qz * uy);
    r.dy = y + qw * uy + (qz * ux - qx * uz);
    r.dz = z + qw * uz + (qx * uy - qy * ux);
    r.ox = c.t[0]; r.oy = c.t[1]; r.oz = c.t[2];
    return r;
}

// Ray parameter s with  o + s*d  on the surface (fixed number of iterations).
DMF_HD double cast(const dmf_synth_scene &s, const Ray &r) {
    double par = (s.plane_z - r.oz) / r.dz;
    for (int i = 0; i < s.ray_iters; ++i) {
        double X = r.ox + par * r.dx, Y = r.oy + par * r.dy;
        par = (relief(s, X, Y) - r.oz) / r.dz;
    }
    return par;
}

// One u8 pixel (supersampled) of the view.
DMF_HD uint8_t shade_pixel(const dmf_synth_scene &s, const dmf_synth_camera &c, int u, int v) {
    int n = s.supersample < 1 ? 1 : s.supersample;
    double acc = 0.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double du = ((double)i + 0.5) / (double)n - 0.5, dv = ((double)j + 0.5) / (double)n - 0.5;
            Ray r = pixel_ray(c, (double)u + du, (double)v + dv);
            double par = cast(s, r);
            acc += texture(s, r.ox + par * r.dx, r.oy + par * r.dy);
        }
    double val = acc / (double)(n * n) * 255.0 + 0.5;
    double f = dfloor(val);
    if (f < 0.0) f = 0.0;
    if (f > 255.0) f = 255.0;
    return (uint8_t)(int)f;
}

// Ground-truth ray distance |OP| of the pixel-centre ray (the quantity the depth maps hold,
// dense_mapping/test_monocular_mapping.cpp:299).
DMF_HD double pixel_distance(const dmf_synth_scene &s, const dmf_synth_camera &c, int u, int v) {
    Ray r = pixel_ray(c, (double)u, (double)v);
    double par = cast(s, r);
    double x = (u - c.cx) / c.fx, y = (v - c.cy) / c.fy;
#ifdef __CUDA_ARCH__
    return par * sqrt(x * x + y * y + 1.0);
#else
    return par * __builtin_sqrt(x * x + y * y + 1.0);
#endif
}

}  // namespace dmf_synth

#endif  // DMF_SYNTH_SCENE_H_
