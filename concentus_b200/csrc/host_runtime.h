// host_runtime.h — host-side plumbing shared by the decoder half (opus_capi.cu), the encoder half (opus_enc_capi.cu) and the encoder
// pipeline (opus_enc_pipe.cu): grow-only device / pinned buffers, the slot registry entry, and the per-thread device selection.
//
// One process can drive several GPUs through the C ABI: every device has its own context in each half (state pool, streams,
// staging buffers, lock).  A host thread picks its device with opus_b200_init(device) (thread-local; threads that never call it
// use the device of the process's first opus_b200_init, or device 0); states are resident on the device of the thread that last
// ran a batch / span call on them.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

enum { kCbMaxDevices = 16 };

struct CbSlotInfo {
    const void *owner;
    uint64_t gen;
};

struct CbDevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) return false;
        cap = want;
        return true;
    }
};
struct CbPinBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) return false;
        cap = want;
        return true;
    }
};

// opus_capi.cu: the device the calling thread codes on (no initialisation, no lock), and its selection
extern "C" int opus_b200_current_device(void);
