"""ctypes handle on the copy-engine frame ring of include/dmf.h (dmf_ring_*): one producer process owns a ring of
frames in its HBM, every rank pulls each frame over NVLink peer-to-peer with a copy engine, ordered by stream memory
operations on a shared flag page (slamplay_b200/csrc/frame_ring.cu).  The reference has no multi-GPU code
(SURVEY.md §2); this is the transport behind ShardedDepthFilter."""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib

HANDLE_BYTES = 192


class RingError(RuntimeError):
    pass


class FrameRing:
    def __init__(self, ptr: C.c_void_p, handle: Optional[bytes], producer: bool):
        self._lib = _lib.load_dmf()
        self._ptr = ptr
        self.handle = handle
        self.producer = producer

    @classmethod
    def create(cls, device: int, n_slots: int, width: int, height: int, n_consumers: int) -> "FrameRing":
        lib = _lib.load_dmf()
        ptr = C.c_void_p()
        buf = (C.c_uint8 * HANDLE_BYTES)()
        rc = lib.dmf_ring_create(int(device), int(n_slots), int(width), int(height), int(n_consumers), C.byref(ptr), buf)
        if rc != 0:
            raise RingError(f"dmf_ring_create failed ({rc}): {lib.dmf_last_error(None).decode()}")
        return cls(ptr, bytes(buf), True)

    @classmethod
    def open(cls, device: int, handle: bytes, consumer: int) -> "FrameRing":
        lib = _lib.load_dmf()
        ptr = C.c_void_p()
        buf = (C.c_uint8 * HANDLE_BYTES).from_buffer_copy(handle)
        rc = lib.dmf_ring_open(int(device), buf, int(consumer), C.byref(ptr))
        if rc != 0:
            raise RingError(f"dmf_ring_open failed ({rc}): {lib.dmf_last_error(None).decode()}")
        return cls(ptr, handle, False)

    def publish(self, frame_ptr: int, step: int, wait_stream: Optional[int] = None) -> None:
        """Enqueue the copy of one frame (pinned host or device memory) into the next slot; asynchronous."""
        rc = self._lib.dmf_ring_publish(self._ptr, C.c_void_p(frame_ptr), step, C.c_void_p(wait_stream) if wait_stream else None)
        if rc != 0:
            raise RingError(f"dmf_ring_publish failed ({rc}): {self._lib.dmf_last_error(None).decode()}")

    def info(self) -> dict:
        a, b, n = C.c_int(), C.c_int(), C.c_uint32()
        self._lib.dmf_ring_info(self._ptr, C.byref(a), C.byref(b), C.byref(n))
        return {"n_slots": a.value, "n_consumers": b.value, "next_frame": n.value}

    def close(self) -> None:
        if self._ptr is not None and self._ptr.value:
            self._lib.dmf_ring_close(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedHostFrames:
    """A stack of frames (F, H, W) uint8 in POSIX shared memory, mapped and page-locked (dmf_host_register) by every rank
    of the node.  With the frames visible to all ranks, each rank can publish "its" frames (frame k by rank k mod N) into
    its own ring, so that the host-to-device uploads are spread over all PCIe links instead of rank 0's alone."""

    def __init__(self, name: str, shape, create: bool):
        import numpy as np
        self._lib = _lib.load_dmf()
        self.name, self.shape, self.created = name, tuple(int(v) for v in shape), create
        path = "/dev/shm/" + name
        self.array = np.memmap(path, dtype=np.uint8, mode="w+" if create else "r+", shape=self.shape)
        self.nbytes = int(np.prod(self.shape))
        self._ptr = self.array.ctypes.data
        rc = self._lib.dmf_host_register(C.c_void_p(self._ptr), self.nbytes)
        if rc != 0:
            raise RingError(f"dmf_host_register failed ({rc}): {self._lib.dmf_last_error(None).decode()}")
        self._registered = True

    def frame_ptr(self, i: int) -> int:
        return self._ptr + i * self.shape[1] * self.shape[2]

    def close(self) -> None:
        if getattr(self, "_registered", False):
            self._lib.dmf_host_unregister(C.c_void_p(self._ptr))
            self._registered = False
            del self.array
            if self.created:
                import os
                try:
                    os.unlink("/dev/shm/" + self.name)
                except OSError:
                    pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
