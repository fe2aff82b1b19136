// frame_ring.cu — per-frame distribution of the current image to the GPUs of one node on the COPY ENGINES.
//
// every rank needs the whole current frame (an epipolar segment can reach anywhere inside the border,
// ref:397-447), so each frame is "broadcast" from the rank that receives it to all ranks.  ncc_kernel keeps every SM
// busy with persistent CTAs, so a collective kernel (NCCL broadcast) cannot get an SM until the previous update has
// drained, which serialises transfer -> moments_kernel -> ncc_kernel (round 1: 26 ms of exposed moments_kernel per
// 4K step at 8 GPUs).  Here no SM is involved:
//
//   producer (rank 0)   H2D / D2D copy of frame k into slot k % S of a ring in ITS device memory (cudaMemcpy2DAsync on
//                       a copy stream), then a stream-ordered 32-bit flag write  filled[slot] = k + 1.
//   consumer (any rank) its context's copy stream waits for that flag (cuStreamWaitValue32), PULLS the slot over NVLink
//                       peer-to-peer (cudaIpcOpenMemHandle mapping, cudaMemcpyAsync: a copy engine) into the context's
//                       own double buffer, writes  released[consumer][slot] = k + 1, and the update is launched
//                       against the local copy (moments_kernel on its side stream, beside the previous ncc_kernel).
//   slot reuse          before frame k overwrites slot k % S the producer's stream waits for released[c][slot] >= k-S+1
//                       of every consumer c.
//
// The flags live in a POSIX shared-memory page that every process registers with CUDA (cudaHostRegisterMapped), so the
// stream memory operations of different processes meet on the same physical words; the hosts never block on each
// other and no host thread polls.  The driver entry points are looked up at run time (cudaGetDriverEntryPoint), so
// libdmf.so keeps loading on a machine without libcuda (the CPU-only ABI tests).
#include "../../include/dmf.h"
#include "dmf_internal.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

namespace {

constexpr uint32_t RING_MAGIC = 0x444d4652u;  // "DMFR"
constexpr int MAX_SLOTS = 16;
constexpr int MAX_CONSUMERS = 16;

struct RingCtl {  // one page of POSIX shared memory
    uint32_t magic, n_slots, n_consumers, pad0;
    uint64_t slot_bytes;
    uint32_t pitch, width, height, pad1;
    volatile uint32_t filled[MAX_SLOTS];                     // frame number + 1 held by the slot
    volatile uint32_t released[MAX_CONSUMERS][MAX_SLOTS];    // last frame number + 1 consumer c pulled out of the slot
};

struct Handle {  // DMF_RING_HANDLE_BYTES, plain bytes that travel between the processes
    uint32_t magic;
    int32_t pid;
    int32_t device;
    int32_t reserved;
    uint64_t dev_ptr;               // valid inside the creating process only
    cudaIpcMemHandle_t ipc;         // 64 bytes
    char shm_name[64];
};
static_assert(sizeof(Handle) <= DMF_RING_HANDLE_BYTES, "ring handle does not fit");

typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
wait32_fn p_wait32 = nullptr;
write32_fn p_write32 = nullptr;

bool load_memops(std::string &why) {
    if (p_wait32 && p_write32) return true;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        why = "cuStreamWaitValue32 is not available from this driver";
        return false;
    }
    p_wait32 = (wait32_fn)f;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) != cudaSuccess || !f || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        why = "cuStreamWriteValue32 is not available from this driver";
        return false;
    }
    p_write32 = (write32_fn)f;
    return true;
}

std::atomic<unsigned> g_ring_counter{0};

}  // namespace

struct dmf_ring {
    bool producer = false;
    int device = 0;
    int consumer = -1;              // index into released[][] (consumers only)
    RingCtl *ctl = nullptr;         // host mapping of the shared page
    CUdeviceptr ctl_dev = 0;        // the same page as the device sees it
    uint8_t *slots = nullptr;       // device memory of the producer (own allocation, IPC mapping, or same-process alias)
    bool slots_ipc = false, slots_owned = false;
    cudaStream_t stream = nullptr;  // producer: the stream the publications are enqueued on
    cudaEvent_t ev_ext = nullptr;
    uint32_t next = 0;              // next frame number to publish / consume
    char shm_name[64] = {0};
    std::string err;
};

namespace {

int rfail(dmf_ring *r, int code, const std::string &msg) {
    if (r) r->err = msg;
    dmf_internal_set_error(msg.c_str());
    return code;
}

#define RCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return rfail(r, DMF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

CUdeviceptr flag_addr(const dmf_ring *r, const volatile uint32_t *host_word) {
    return r->ctl_dev + (CUdeviceptr)((const char *)host_word - (const char *)r->ctl);
}

int map_ctl(dmf_ring *r, const char *name, bool create) {
    int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return rfail(r, DMF_ERR_STATE, std::string("shm_open(") + name + ") failed");
    const size_t bytes = (sizeof(RingCtl) + 4095) / 4096 * 4096;
    if (create && ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name); return rfail(r, DMF_ERR_NOMEM, "ftruncate on the ring control page failed"); }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { if (create) shm_unlink(name); return rfail(r, DMF_ERR_NOMEM, "mmap of the ring control page failed"); }
    r->ctl = (RingCtl *)p;
    std::strncpy(r->shm_name, name, sizeof(r->shm_name) - 1);
    if (create) std::memset(p, 0, bytes);
    RCU(cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    void *d = nullptr;
    RCU(cudaHostGetDevicePointer(&d, p, 0));
    r->ctl_dev = (CUdeviceptr)d;
    return DMF_OK;
}

}  // namespace

extern "C" {

int dmf_ring_create(int device, int n_slots, int width, int height, int n_consumers, dmf_ring **out, uint8_t *handle_out) {
    dmf_ring *r = nullptr;
    if (!out || !handle_out) return rfail(r, DMF_ERR_INVALID, "dmf_ring_create: NULL argument");
    *out = nullptr;
    if (n_slots < 2 || n_slots > MAX_SLOTS || n_consumers < 1 || n_consumers > MAX_CONSUMERS || width < 1 || height < 1)
        return rfail(r, DMF_ERR_INVALID, "dmf_ring_create: need 2 <= n_slots <= 16, 1 <= n_consumers <= 16");
    std::string why;
    RCU(cudaSetDevice(device));
    if (!load_memops(why)) return rfail(r, DMF_ERR_CUDA, "dmf_ring_create: " + why);
    r = new (std::nothrow) dmf_ring();
    if (!r) return rfail(r, DMF_ERR_NOMEM, "dmf_ring_create: out of host memory");
    r->producer = true;
    r->device = device;
    char name[64];
    std::snprintf(name, sizeof(name), "/dmf_ring_%d_%u", (int)getpid(), g_ring_counter.fetch_add(1));
    int rc = map_ctl(r, name, true);
    if (rc) { dmf_ring_close(r); return rc; }
    const uint32_t pitch = (uint32_t)((width + 15) / 16 * 16);
    const size_t slot_bytes = (size_t)pitch * height;
    cudaError_t e = cudaMalloc(&r->slots, slot_bytes * n_slots);
    if (e == cudaSuccess) e = cudaMemset(r->slots, 0, slot_bytes * n_slots);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_ext, cudaEventDisableTiming);
    Handle h{};
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h.ipc, r->slots);
    if (e != cudaSuccess) { rc = rfail(r, DMF_ERR_CUDA, std::string("dmf_ring_create: ") + cudaGetErrorString(e)); dmf_ring_close(r); return rc; }
    r->slots_owned = true;
    r->ctl->n_slots = (uint32_t)n_slots; r->ctl->n_consumers = (uint32_t)n_consumers; r->ctl->slot_bytes = slot_bytes;
    r->ctl->pitch = pitch; r->ctl->width = (uint32_t)width; r->ctl->height = (uint32_t)height;
    r->ctl->magic = RING_MAGIC;
    h.magic = RING_MAGIC; h.pid = (int32_t)getpid(); h.device = device; h.dev_ptr = (uint64_t)(uintptr_t)r->slots;
    std::strncpy(h.shm_name, name, sizeof(h.shm_name) - 1);
    std::memset(handle_out, 0, DMF_RING_HANDLE_BYTES);
    std::memcpy(handle_out, &h, sizeof(h));
    *out = r;
    return DMF_OK;
}

int dmf_ring_open(int device, const uint8_t *handle, int consumer, dmf_ring **out) {
    dmf_ring *r = nullptr;
    if (!out || !handle) return rfail(r, DMF_ERR_INVALID, "dmf_ring_open: NULL argument");
    *out = nullptr;
    Handle h;
    std::memcpy(&h, handle, sizeof(h));
    if (h.magic != RING_MAGIC) return rfail(r, DMF_ERR_INVALID, "dmf_ring_open: not a ring handle");
    if (consumer < 0 || consumer >= MAX_CONSUMERS) return rfail(r, DMF_ERR_INVALID, "dmf_ring_open: consumer index out of range");
    std::string why;
    RCU(cudaSetDevice(device));
    if (!load_memops(why)) return rfail(r, DMF_ERR_CUDA, "dmf_ring_open: " + why);
    r = new (std::nothrow) dmf_ring();
    if (!r) return rfail(r, DMF_ERR_NOMEM, "dmf_ring_open: out of host memory");
    r->device = device;
    r->consumer = consumer;
    h.shm_name[sizeof(h.shm_name) - 1] = 0;
    int rc = map_ctl(r, h.shm_name, false);
    if (rc) { dmf_ring_close(r); return rc; }
    if (r->ctl->magic != RING_MAGIC || (uint32_t)consumer >= r->ctl->n_consumers) {
        rc = rfail(r, DMF_ERR_INVALID, "dmf_ring_open: control page mismatch / consumer index >= n_consumers");
        dmf_ring_close(r);
        return rc;
    }
    if (h.pid == (int32_t)getpid()) {  // same process: an IPC handle cannot be opened by its creator
        r->slots = (uint8_t *)(uintptr_t)h.dev_ptr;
        if (device != h.device) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device, h.device);
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0); if (e != cudaSuccess) cudaGetLastError(); }
        }
    } else {
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { rc = rfail(r, DMF_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); dmf_ring_close(r); return rc; }
        r->slots = (uint8_t *)p;
        r->slots_ipc = true;
    }
    *out = r;
    return DMF_OK;
}

void dmf_ring_close(dmf_ring *r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
    if (r->ev_ext) cudaEventDestroy(r->ev_ext);
    if (r->slots_ipc && r->slots) cudaIpcCloseMemHandle(r->slots);
    if (r->slots_owned && r->slots) cudaFree(r->slots);
    if (r->ctl) {
        const size_t bytes = (sizeof(RingCtl) + 4095) / 4096 * 4096;
        cudaHostUnregister(r->ctl);
        munmap(r->ctl, bytes);
        if (r->producer) shm_unlink(r->shm_name);
    }
    cudaGetLastError();
    delete r;
}

int dmf_ring_info(const dmf_ring *r, int *n_slots, int *n_consumers, uint32_t *published_or_consumed) {
    if (!r) return rfail(nullptr, DMF_ERR_INVALID, "dmf_ring_info: NULL ring");
    if (n_slots) *n_slots = (int)r->ctl->n_slots;
    if (n_consumers) *n_consumers = (int)r->ctl->n_consumers;
    if (published_or_consumed) *published_or_consumed = r->next;
    return DMF_OK;
}

int dmf_ring_publish(dmf_ring *r, const uint8_t *frame, size_t step, void *wait_stream) {
    if (!r || !frame) return rfail(r, DMF_ERR_INVALID, "dmf_ring_publish: NULL argument");
    if (!r->producer) return rfail(r, DMF_ERR_STATE, "dmf_ring_publish: this handle was opened as a consumer");
    RingCtl *ctl = r->ctl;
    if (step < ctl->width) return rfail(r, DMF_ERR_INVALID, "dmf_ring_publish: step < width");
    RCU(cudaSetDevice(r->device));
    const uint32_t k = r->next, S = ctl->n_slots, s = k % S;
    if (wait_stream) {  // a device frame that work queued on the caller's stream is still producing
        RCU(cudaEventRecord(r->ev_ext, (cudaStream_t)wait_stream));
        RCU(cudaStreamWaitEvent(r->stream, r->ev_ext, 0));
    }
    if (k >= S)  // the slot still holds frame k - S until every consumer has pulled it
        for (uint32_t c = 0; c < ctl->n_consumers; ++c)
            if (p_wait32((CUstream)r->stream, flag_addr(r, &ctl->released[c][s]), k - S + 1, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return rfail(r, DMF_ERR_CUDA, "dmf_ring_publish: cuStreamWaitValue32 failed");
    if (step == ctl->pitch && ctl->width == ctl->pitch)  // contiguous frame: one linear DMA instead of one descriptor per row
        RCU(cudaMemcpyAsync(r->slots + (size_t)s * ctl->slot_bytes, frame, ctl->slot_bytes, cudaMemcpyDefault, r->stream));
    else
        RCU(cudaMemcpy2DAsync(r->slots + (size_t)s * ctl->slot_bytes, ctl->pitch, frame, step, ctl->width, ctl->height, cudaMemcpyDefault, r->stream));
    if (p_write32((CUstream)r->stream, flag_addr(r, &ctl->filled[s]), k + 1, 0) != CUDA_SUCCESS)
        return rfail(r, DMF_ERR_CUDA, "dmf_ring_publish: cuStreamWriteValue32 failed");
    r->next = k + 1;
    return DMF_OK;
}

int dmf_update_ring(dmf_ctx *ctx, dmf_ring *r, const double q[4], const double t[3]) {
    if (!ctx || !r || !q || !t) return rfail(r, DMF_ERR_INVALID, "dmf_update_ring: NULL argument");
    if (r->consumer < 0) return rfail(r, DMF_ERR_STATE, "dmf_update_ring: the ring must be opened with dmf_ring_open (also in the producer's process)");
    RingCtl *ctl = r->ctl;
    const uint32_t k = r->next, s = k % ctl->n_slots;
    dmf_internal_stage st;
    int rc = dmf_internal_stage_begin(ctx, (int)ctl->width, (int)ctl->height, &st);
    if (rc) return rc;
    RCU(cudaSetDevice(r->device));
    if (p_wait32((CUstream)st.copy_stream, flag_addr(r, &ctl->filled[s]), k + 1, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
        return rfail(r, DMF_ERR_CUDA, "dmf_update_ring: cuStreamWaitValue32 failed");
    const uint8_t *src = r->slots + (size_t)s * ctl->slot_bytes;
    if ((size_t)st.pitch == ctl->pitch) RCU(cudaMemcpyAsync(st.dst, src, ctl->slot_bytes, cudaMemcpyDefault, (cudaStream_t)st.copy_stream));
    else RCU(cudaMemcpy2DAsync(st.dst, st.pitch, src, ctl->pitch, ctl->width, ctl->height, cudaMemcpyDefault, (cudaStream_t)st.copy_stream));
    if (p_write32((CUstream)st.copy_stream, flag_addr(r, &ctl->released[r->consumer][s]), k + 1, 0) != CUDA_SUCCESS)
        return rfail(r, DMF_ERR_CUDA, "dmf_update_ring: cuStreamWriteValue32 failed");
    r->next = k + 1;
    return dmf_internal_stage_launch(ctx, &st, q, t);
}

}  // extern "C"
