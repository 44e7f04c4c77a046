// enc_pipe_host.h — internal interface between the encoder's C ABI (opus_enc_capi.cu) and the frame-synchronous encoder pipeline
// (opus_enc_pipe.cu, kernels over celt_enc_pipe.cuh).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "opus_state.h"

struct EncPipeCall {
    CbEncState *pool;        // the encoder state pool (device)
    const int *d_slots;      // [n] pool slot of pipeline stream k (device)
    const int *d_sidx;       // [n] index of pipeline stream k among the call's streams: its PCM / packet / ret rows (device)
    int n;                   // pipeline streams
    int F, f0, f1;           // row pitch in frames, and the frame range [f0, f1) this call codes
    int frame_size, channels, Fs;
    int max_bytes, stride;   // min(1276, max_data_bytes), packet slot pitch
    const int16_t *d_pcm;    // [streams][F][frame_size * channels]
    uint8_t *d_data;         // [streams][F][stride]
    int *d_rets;             // [streams][F]
    unsigned *d_ranges;      // [streams][F] or null
};

// Enqueue the whole span on `stream` (side streams are forked from and joined back into it).  Returns the number of kernel
// launches, or a negative OPUS error (allocation failure).  Must be called with the encoder context lock held.
int enc_pipe_enqueue(const EncPipeCall &c, cudaStream_t stream);
// 1 when the stream may go through the pipeline (host-side test on its ctl-visible configuration)
int enc_pipe_takes(const CbEncState *st, int frame_size, int out_data_bytes);
// statistics of the split band loop: leaves chain-X searched itself (its budget differed from chain-S's) / leaves listed
void enc_pipe_stats(long long *misses, long long *leaves);
