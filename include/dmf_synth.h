/*
 * dmf_synth.h — C ABI of the synthetic-sequence renderer (benchmark / test INPUT generator,
 * SURVEY.md §8d).  Not part of the reference's call surface: the reference reads the REMODE
 * set from disk (dense_mapping/test_monocular_mapping.cpp:317-352), which is a network
 * download and absent here, so sequences of the same shape are rendered instead.
 *
 * dmf_synth_render_host   lives in slamplay_b200/libdmf_synth_cpu.so (g++, OpenMP)
 * dmf_synth_render_device lives in slamplay_b200/libdmf_synth.so      (CUDA, sm_100a; NOT in the product library libdmf.so)
 * Both evaluate slamplay_b200/csrc/synth_scene.h with FMA contraction off and produce
 * bit-identical images.
 */
#ifndef DMF_SYNTH_H_
#define DMF_SYNTH_H_

#include <stddef.h>
#include <stdint.h>
#include "../slamplay_b200/csrc/synth_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Render the u8 view of `cam` into img (step bytes per row) and, when dist != NULL, the
 * ground-truth ray-distance map (doubles, dist_step bytes per row).  Returns 0 or -1. */
int dmf_synth_render_host(const dmf_synth_scene *scene, const dmf_synth_camera *cam, uint8_t *img, size_t step,
                          double *dist, size_t dist_step);

/* Same, into device memory on the current CUDA device, asynchronously on `stream`
 * (a cudaStream_t, may be NULL).  Returns 0 or a negative dmf_status. */
int dmf_synth_render_device(const dmf_synth_scene *scene, const dmf_synth_camera *cam, uint8_t *img_dev, size_t pitch,
                            double *dist_dev, size_t dist_pitch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DMF_SYNTH_H_ */
