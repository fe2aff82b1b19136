// dmf_internal.h — the few context internals other translation units of libdmf.so need (frame_ring.cu).
#pragma once
#include <stdint.h>

#include "../../include/dmf.h"

// thread-local + context error message (dmf_last_error)
void dmf_internal_set_error(const char *msg);

// One frame staged into the context's own double buffer by a copy the CALLER enqueues on the context's copy stream.
struct dmf_internal_stage {
    void *copy_stream;  // cudaStream_t
    uint8_t *dst;       // device buffer of the frame (pitched)
    int pitch;          // bytes per row of dst
    int buffer;         // which of the two buffers
};
// Picks the next buffer and makes the copy stream wait until the kernels of two updates ago have consumed it.
int dmf_internal_stage_begin(dmf_ctx *ctx, int width, int height, dmf_internal_stage *st);
// Marks the copy as complete (event on the copy stream) and launches one update() against the staged frame.
int dmf_internal_stage_launch(dmf_ctx *ctx, const dmf_internal_stage *st, const double q[4], const double t[3]);
