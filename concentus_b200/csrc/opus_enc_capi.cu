// opus_enc_capi.cu — encoder kernels and the encoder half of the C ABI of libconcentus_b200.so (include/opus_b200.h).
//
// Boundary: libopus's public encoder API (opus-fix/include/opus.h:171-328, src/opus_encoder.c:150-252,2007-2507) plus our
// batch / span calls.  Host side = argument checks, ctl, state residency and copies; everything from the Opus-layer rate
// decisions down to ec_enc_done runs on the device, one warp per stream, F frames per launch (the frames of a stream are
// serially dependent through the encoder state; streams are independent).
// There is NO CPU path: if CUDA is unusable every codec call returns OPUS_INTERNAL_ERROR.
#if !defined(CB_NO_PHASE_SYNC) && !defined(CB_PHASE_SYNC)
#define CB_PHASE_SYNC 1      // see the block-shape note below
#endif
#define CB_SMALL_CODE 1   // see celt_simt.cuh: the encoder kernel is instruction-cache bound when everything is inlined
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_set>
#include <vector>

#include "../../include/opus_b200.h"
#include "enc_pipe_host.h"
#include "host_runtime.h"
#include "opus_encoder_dev.cuh"

using namespace cb;

extern "C" int opus_b200_device_index(void);   // opus_capi.cu: the device opus_b200_init selected (initialises the runtime)

// Caller-visible state block: pointer-free and memcpy-able like the reference's (tests/test_opus_encode.c:198,214).
// Same residency protocol as OpusDecoder (opus_capi.cu).
struct OpusEncoder {
    uint32_t magic;
    int32_t slot;
    uint64_t gen;
    int32_t host_current;
    int32_t device;        // the device whose pool `slot` refers to
    CbEncState st;
};
static const uint32_t kEncMagic = 0x0B200E4Cu;

// Block shape (measured on B200, profiles/r1_enc_*): ONE block of 14 warps per SM, the warps phase-synchronised
// (celt_simt.cuh WarpTeam::phase).  The kernel is ~250 KB of SASS against a 32 KB L1.5 instruction cache: with free-running
// warps 35-70 % of the issue slots were lost to instruction fetch; letting the 14 streams of an SM walk the frame in step
// doubled the throughput.  2,072 streams are resident per B200, so 4,096 streams are two full waves.
#ifndef CB_ENC_WPB
#define CB_ENC_WPB 14         // warps (= streams) per block
#endif
#ifndef CB_ENC_MINBLOCKS
#define CB_ENC_MINBLOCKS 1
#endif
#if !defined(CB_NO_PHASE_SYNC) && !defined(CB_PHASE_SYNC)
#define CB_PHASE_SYNC 1
#endif

// Shared memory of one warp: the frame working set plus the head of the stream's state (scalars + band-energy histories),
// which stays on chip for the whole span and is written back once at the end.
struct EncWarpSmem {
    EncShared S;
    int head[CB_ENC_HEAD_BYTES / 4];
};

// The one-kernel path: takes every stream the frame-synchronous pipeline (opus_enc_pipe.cu) does not (OPUS_APPLICATION_AUDIO,
// 40 / 60 ms frames, "PLC frame" budgets) — and everything when CB200_ENC_PIPE=0.
// One warp per stream, frames 0..F-1 in order.  PCM of frame (s,f): pcm[(s*F+f)*frame_size*channels]; packet slot:
// data[(s*F+f)*stride], at most max_bytes are written; rets[s*F+f] = packet length or error.  One launch codes frames [f0, f1).
__global__ void __launch_bounds__(CB_ENC_WPB * 32, CB_ENC_MINBLOCKS)
encode_span_kernel(CbEncState *pool, const int *slots, const int *sidx, EncGlobal *scratch, const int16_t *pcm, uint8_t *data, int *rets,
                   unsigned *ranges, int n, int F, int f0, int f1, int frame_size, int max_bytes, int stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * CB_ENC_WPB + warp;
    if (s >= n) {
#if defined(CB_PHASE_SYNC)
        {
            WarpTeam idle{lane};
            for (int f = 0; f < (f1 - f0) * kEncPhases; f++) idle.phase();   // keep the block's phase barriers balanced
        }
#endif
        return;
    }
    EncWarpSmem &W = reinterpret_cast<EncWarpSmem *>(smem_raw)[warp];
    CbEncState *gst = pool + slots[s];
    for (int i = lane; i < CB_ENC_HEAD_BYTES / 4; i += 32) W.head[i] = reinterpret_cast<const int *>(gst)[i];
    __syncwarp();
    CbEncState *st = reinterpret_cast<CbEncState *>(W.head);   // only the head fields are valid through this pointer
    EncGlobal &G = scratch[s];
    const int channels = st->channels;
    WarpTeam tm{lane};
    const int srow = sidx ? sidx[s] : s;   // the stream's row of PCM / packets / rets (a launch may cover a subset of a call's streams)
    for (int f = f0; f < f1; f++) {
        const size_t k = (size_t)srow * F + f;
        const int r = opus_encode_frame(tm, st, gst, W.S, G, pcm + k * frame_size * channels, frame_size, data + k * stride, max_bytes);
        if (lane == 0) {
            rets[k] = r;
            if (ranges) ranges[k] = st->rangeFinal;   // OPUS_GET_FINAL_RANGE after this frame (what opus_demo stores per packet)
        }
        __syncwarp();
    }
    for (int i = lane; i < CB_ENC_HEAD_BYTES / 4; i += 32) reinterpret_cast<int *>(gst)[i] = W.head[i];
}

__global__ void enc_scatter_states_kernel(CbEncState *pool, const int *slots, const CbEncState *stage, int n) {
    const int words = sizeof(CbEncState) / 4;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int *src = reinterpret_cast<const int *>(stage + k);
        int *dst = reinterpret_cast<int *>(pool + slots[k]);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
}
__global__ void enc_gather_states_kernel(const CbEncState *pool, const int *slots, CbEncState *stage, int n) {
    const int words = sizeof(CbEncState) / 4;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int *src = reinterpret_cast<const int *>(pool + slots[k]);
        int *dst = reinterpret_cast<int *>(stage + k);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
}

namespace {

inline int hmin(int a, int b) { return a < b ? a : b; }

typedef CbSlotInfo SlotInfo;
typedef CbDevBuf DevBuf;
typedef CbPinBuf PinBuf;

struct EncCtx {
    std::mutex mu;
    bool tried = false, ok = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    enum { kMaxSub = 8 };
    cudaEvent_t ev_in[kMaxSub] = {}, ev_done[kMaxSub] = {}, ev_prev = nullptr;
    CbEncState *pool = nullptr;
    int pool_cap = 0;
    std::vector<SlotInfo> reg;
    std::vector<int> free_slots;
    DevBuf d_slots, d_pcm, d_data, d_rets, d_ranges, d_stage, d_scratch, d_split;
    PinBuf h_stage, h_slots, h_split[4];
    cudaStream_t legacy_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_legacy = nullptr;
    int use_pipe = 1;
    unsigned split_seq = 0;
    long long pipe_streams = 0, legacy_streams = 0;
    long long launches = 0;
    float last_ms = 0.f;
    double total_ms = 0;
    int device = 0;
    std::mutex deferred_mu;
    std::vector<std::pair<int, const void *>> deferred;   // slots of this pool whose owners moved to another device
};
// one context per device; the calling thread's device is the one it selected with opus_b200_init (host_runtime.h)
EncCtx e_ctxs[kCbMaxDevices];
#define e (e_ctxs[opus_b200_current_device()])
enum { kStageStates = 4096 };   // states per upload / download slice

bool ctx_init_locked() {
    if (e.tried) {
        if (e.ok) {
            cudaSetDevice(e.device);
            if (!e.deferred.empty()) {
                std::lock_guard<std::mutex> lk(e.deferred_mu);
                for (auto &pr : e.deferred)
                    if (pr.first >= 0 && pr.first < e.pool_cap && e.reg[pr.first].owner == pr.second) {
                        e.reg[pr.first].owner = nullptr;
                        e.reg[pr.first].gen++;
                        e.free_slots.push_back(pr.first);
                    }
                e.deferred.clear();
            }
        }
        return e.ok;
    }
    e.tried = true;
    const int dev = opus_b200_device_index();
    if (dev < 0) return false;
    e.device = dev;
    if (cudaSetDevice(dev) != cudaSuccess) return false;
    if (cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    cudaEventCreate(&e.ev0);
    cudaEventCreate(&e.ev1);
    for (int i = 0; i < EncCtx::kMaxSub; i++) {
        cudaEventCreateWithFlags(&e.ev_in[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e.ev_done[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&e.ev_prev, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&e.legacy_stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    cudaEventCreateWithFlags(&e.ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e.ev_legacy, cudaEventDisableTiming);
    if (const char *v = getenv("CB200_ENC_PIPE")) e.use_pipe = atoi(v);
    if (!e.h_stage.reserve(sizeof(CbEncState) * 64) || !e.d_stage.reserve(sizeof(CbEncState) * 64)) return false;
    cudaFuncSetAttribute(encode_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CB_ENC_WPB * sizeof(EncWarpSmem)));
#ifndef CB_ENC_CARVEOUT
#define CB_ENC_CARVEOUT 70   // % of the 228 KB: the block needs 157 KB of shared memory; the rest stays L1 for stacks and tables
#endif
    cudaFuncSetAttribute(encode_span_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, CB_ENC_CARVEOUT);
    e.ok = cudaGetLastError() == cudaSuccess;
    return e.ok;
}

bool pool_reserve_locked(int need_total) {
    if (need_total <= e.pool_cap) return true;
    int ncap = e.pool_cap ? e.pool_cap : 64;
    while (ncap < need_total) ncap *= 2;
    CbEncState *np = nullptr;
    if (cudaMalloc(&np, sizeof(CbEncState) * (size_t)ncap) != cudaSuccess) return false;
    if (e.pool) {
        cudaMemcpyAsync(np, e.pool, sizeof(CbEncState) * (size_t)e.pool_cap, cudaMemcpyDeviceToDevice, e.stream);
        cudaStreamSynchronize(e.stream);
        cudaFree(e.pool);
    }
    for (int i = ncap - 1; i >= e.pool_cap; i--) e.free_slots.push_back(i);
    e.reg.resize(ncap, SlotInfo{nullptr, 0});
    e.pool = np;
    e.pool_cap = ncap;
    return true;
}

inline bool resident(const OpusEncoder *d) {
    return d->device == e.device && d->slot >= 0 && d->slot < e.pool_cap && e.reg[d->slot].owner == d && e.reg[d->slot].gen == d->gen;
}
// the block's slot lives in another device's pool: that device's next call gives it back
void release_elsewhere(OpusEncoder *d) {
    if (d->slot >= 0 && d->device >= 0 && d->device < kCbMaxDevices && d->device != e.device) {
        EncCtx &o = e_ctxs[d->device];
        std::lock_guard<std::mutex> lk(o.deferred_mu);
        o.deferred.emplace_back(d->slot, (const void *)d);
        d->slot = -1;
    }
}

int make_host_current_locked(OpusEncoder *d) {
    if (d->host_current) return OPUS_OK;
    // (a state resident on another device has to be synchronised by a thread of that device: OPUS_INVALID_STATE here)
    if (d->device != e.device || d->slot < 0 || d->slot >= e.pool_cap || e.reg[d->slot].gen != d->gen) return OPUS_INVALID_STATE;
    if (cudaMemcpyAsync(&d->st, e.pool + d->slot, sizeof(CbEncState), cudaMemcpyDeviceToHost, e.stream) != cudaSuccess) return OPUS_INTERNAL_ERROR;
    if (cudaStreamSynchronize(e.stream) != cudaSuccess) return OPUS_INTERNAL_ERROR;
    d->host_current = 1;
    return OPUS_OK;
}

void release_slot_locked(OpusEncoder *d) {
    release_elsewhere(d);
    if (d->slot >= 0 && d->slot < e.pool_cap && e.reg[d->slot].owner == d) {
        e.reg[d->slot].owner = nullptr;
        e.reg[d->slot].gen++;
        e.free_slots.push_back(d->slot);
    }
    d->slot = -1;
}

int make_resident_locked(OpusEncoder **st, int n, int *h_slots) {
    int need_new = 0;
    for (int i = 0; i < n; i++) {
        OpusEncoder *d = st[i];
        if (!d || d->magic != kEncMagic) return OPUS_BAD_ARG;
        if (!resident(d)) need_new++;
    }
    if (n > 1) {   // the same state twice in one batch would race two warps on it
        std::unordered_set<const void *> seen;
        seen.reserve((size_t)n * 2);
        for (int i = 0; i < n; i++)
            if (!seen.insert(st[i]).second) return OPUS_BAD_ARG;
    }
    const int in_use = e.pool_cap - (int)e.free_slots.size();
    if (!pool_reserve_locked(in_use + need_new)) return OPUS_ALLOC_FAIL;
    std::vector<int> up_idx;
    for (int i = 0; i < n; i++) {
        OpusEncoder *d = st[i];
        if (resident(d)) {
            if (d->host_current) up_idx.push_back(i);
        } else {
            if (!d->host_current) {
                int rc = make_host_current_locked(d);
                if (rc != OPUS_OK) return rc;
            }
            release_elsewhere(d);
            d->slot = e.free_slots.back();
            e.free_slots.pop_back();
            d->device = e.device;
            e.reg[d->slot].owner = d;
            d->gen = ++e.reg[d->slot].gen;
            up_idx.push_back(i);
        }
        h_slots[i] = d->slot;
    }
    {
        const size_t slice = up_idx.size() < (size_t)kStageStates ? up_idx.size() : (size_t)kStageStates;
        if (slice > 0 && (!e.h_stage.reserve(sizeof(CbEncState) * slice) || !e.d_stage.reserve(sizeof(CbEncState) * slice))) return OPUS_ALLOC_FAIL;
    }
    CbEncState *hs = (CbEncState *)e.h_stage.p;
    if (!e.d_slots.reserve(sizeof(int) * (size_t)(n > kStageStates ? n : kStageStates))) return OPUS_ALLOC_FAIL;
    for (size_t base = 0; base < up_idx.size(); base += kStageStates) {
        const int cnt = (int)((up_idx.size() - base) < (size_t)kStageStates ? (up_idx.size() - base) : kStageStates);
        std::vector<int> sl(cnt);
        for (int k = 0; k < cnt; k++) {
            OpusEncoder *d = st[up_idx[base + k]];
            memcpy(&hs[k], &d->st, sizeof(CbEncState));
            sl[k] = d->slot;
        }
        cudaMemcpyAsync(e.d_stage.p, hs, sizeof(CbEncState) * (size_t)cnt, cudaMemcpyHostToDevice, e.stream);
        cudaMemcpyAsync(e.d_slots.p, sl.data(), sizeof(int) * (size_t)cnt, cudaMemcpyHostToDevice, e.stream);
        enc_scatter_states_kernel<<<cnt < 2048 ? cnt : 2048, 256, 0, e.stream>>>(e.pool, (const int *)e.d_slots.p, (const CbEncState *)e.d_stage.p, cnt);
        if (cudaStreamSynchronize(e.stream) != cudaSuccess) return OPUS_INTERNAL_ERROR;
    }
    return OPUS_OK;
}

void mark_device_newer_locked(OpusEncoder **st, int n) {
    for (int i = 0; i < n; i++) {
        OpusEncoder *d = st[i];
        d->gen = ++e.reg[d->slot].gen;
        d->host_current = 0;
    }
}

int sync_states_locked(OpusEncoder **st, int n, bool release) {
    std::vector<int> idx;
    for (int i = 0; i < n; i++) {
        OpusEncoder *d = st[i];
        if (!d || d->magic != kEncMagic) return OPUS_BAD_ARG;
        if (!d->host_current) {
            if (!resident(d)) {
                int rc = make_host_current_locked(d);
                if (rc != OPUS_OK) return rc;
            } else {
                idx.push_back(i);
            }
        }
    }
    {
        const size_t slice = idx.size() < (size_t)kStageStates ? idx.size() : (size_t)kStageStates;
        if (slice > 0 && (!e.h_stage.reserve(sizeof(CbEncState) * slice) || !e.d_stage.reserve(sizeof(CbEncState) * slice))) return OPUS_ALLOC_FAIL;
    }
    CbEncState *hs = (CbEncState *)e.h_stage.p;
    if (!e.d_slots.reserve(sizeof(int) * (size_t)kStageStates)) return OPUS_ALLOC_FAIL;
    for (size_t base = 0; base < idx.size(); base += kStageStates) {
        const int cnt = (int)((idx.size() - base) < (size_t)kStageStates ? (idx.size() - base) : kStageStates);
        std::vector<int> sl(cnt);
        for (int k = 0; k < cnt; k++) sl[k] = st[idx[base + k]]->slot;
        cudaMemcpyAsync(e.d_slots.p, sl.data(), sizeof(int) * (size_t)cnt, cudaMemcpyHostToDevice, e.stream);
        enc_gather_states_kernel<<<cnt < 2048 ? cnt : 2048, 256, 0, e.stream>>>(e.pool, (const int *)e.d_slots.p, (CbEncState *)e.d_stage.p, cnt);
        cudaMemcpyAsync(hs, e.d_stage.p, sizeof(CbEncState) * (size_t)cnt, cudaMemcpyDeviceToHost, e.stream);
        if (cudaStreamSynchronize(e.stream) != cudaSuccess) return OPUS_INTERNAL_ERROR;
        for (int k = 0; k < cnt; k++) {
            OpusEncoder *d = st[idx[base + k]];
            memcpy(&d->st, &hs[k], sizeof(CbEncState));
            d->host_current = 1;
        }
    }
    if (release)
        for (int i = 0; i < n; i++) release_slot_locked(st[i]);
    return OPUS_OK;
}

// Common checks of a span call; all streams must share Fs / channels, and frame_size must be the size every stream actually
// codes: opus_encode runs frame_size_select (opus_encoder.c:807-826,2007-2025) first, so a stream whose
// OPUS_SET_EXPERT_FRAME_DURATION selects another size, or a size that is no Opus frame, is an argument error here (the kernel
// would otherwise walk PCM rows of the wrong length).  The ctl-visible configuration never changes on the device, so a stale host
// block still holds it.
int check_span(OpusEncoder **st, int n, int frame_size) {
    for (int i = 0; i < n; i++) {
        if (!st[i] || st[i]->magic != kEncMagic) return OPUS_BAD_ARG;
        if (st[i]->st.channels != st[0]->st.channels || st[i]->st.Fs != st[0]->st.Fs) return OPUS_BAD_ARG;
        if (frame_size_select(frame_size, st[i]->st.variable_duration, st[i]->st.Fs) != frame_size) return OPUS_BAD_ARG;
    }
    return OPUS_OK;
}

// How a call's streams are divided between the two encoder paths (device arrays in e.d_split).
struct SpanSplit {
    int n_pipe = 0, n_legacy = 0;
    const int *d_pslots = nullptr, *d_psidx = nullptr, *d_lslots = nullptr, *d_lsidx = nullptr;
};

// Decide per stream (host side, ctl-visible configuration only) and upload the two index lists.
bool split_span(OpusEncoder **st, const int *h_slots, int n, int frame_size, int max_bytes, SpanSplit &sp) {
    PinBuf &hb = e.h_split[e.split_seq++ & 3];   // a small ring: an earlier call's upload may still be in flight
    if (!hb.reserve(sizeof(int) * 4 * (size_t)n) || !e.d_split.reserve(sizeof(int) * 4 * (size_t)n)) return false;
    int *h = (int *)hb.p;
    int *pslots = h, *psidx = h + n, *lslots = h + 2 * n, *lsidx = h + 3 * n;
    for (int i = 0; i < n; i++) {
        if (e.use_pipe && enc_pipe_takes(&st[i]->st, frame_size, max_bytes)) { pslots[sp.n_pipe] = h_slots[i]; psidx[sp.n_pipe++] = i; }
        else { lslots[sp.n_legacy] = h_slots[i]; lsidx[sp.n_legacy++] = i; }
    }
    cudaMemcpyAsync(e.d_split.p, h, sizeof(int) * 4 * (size_t)n, cudaMemcpyHostToDevice, e.stream);
    const int *d = (const int *)e.d_split.p;
    sp.d_pslots = d; sp.d_psidx = d + n; sp.d_lslots = d + 2 * n; sp.d_lsidx = d + 3 * n;
    e.pipe_streams += sp.n_pipe;
    e.legacy_streams += sp.n_legacy;
    return true;
}

int launch_frames(const SpanSplit &sp, const int16_t *d_pcm, uint8_t *d_data, int *d_rets, int F, int f0, int f1, int frame_size, int channels, int Fs,
                  int max_bytes, int stride, unsigned *d_ranges = nullptr) {
    if (sp.n_legacy > 0) {
        // the one-kernel path runs beside the pipeline on its own stream
        if (!e.d_scratch.reserve(sizeof(EncGlobal) * (size_t)sp.n_legacy)) return OPUS_ALLOC_FAIL;
        cudaStream_t ls = sp.n_pipe > 0 ? e.legacy_stream : e.stream;
        if (sp.n_pipe > 0) {
            cudaEventRecord(e.ev_fork, e.stream);
            cudaStreamWaitEvent(ls, e.ev_fork, 0);
        }
        encode_span_kernel<<<(sp.n_legacy + CB_ENC_WPB - 1) / CB_ENC_WPB, CB_ENC_WPB * 32, CB_ENC_WPB * sizeof(EncWarpSmem), ls>>>(
            e.pool, sp.d_lslots, sp.d_lsidx, (EncGlobal *)e.d_scratch.p, d_pcm, d_data, d_rets, d_ranges, sp.n_legacy, F, f0, f1, frame_size, max_bytes,
            stride);
        e.launches++;
        if (sp.n_pipe > 0) cudaEventRecord(e.ev_legacy, ls);
    }
    if (sp.n_pipe > 0) {
        EncPipeCall c;
        c.pool = e.pool; c.d_slots = sp.d_pslots; c.d_sidx = sp.d_psidx; c.n = sp.n_pipe;
        c.F = F; c.f0 = f0; c.f1 = f1; c.frame_size = frame_size; c.channels = channels; c.Fs = Fs;
        c.max_bytes = max_bytes; c.stride = stride;
        c.d_pcm = d_pcm; c.d_data = d_data; c.d_rets = d_rets; c.d_ranges = d_ranges;
        const int r = enc_pipe_enqueue(c, e.stream);
        if (r < 0) return r;
        e.launches += r;
        if (sp.n_legacy > 0) cudaStreamWaitEvent(e.stream, e.ev_legacy, 0);
    }
    return OPUS_OK;
}
int launch_span(const SpanSplit &sp, const int16_t *d_pcm, uint8_t *d_data, int *d_rets, int F, int frame_size, int channels, int Fs, int max_bytes,
                int stride, unsigned *d_ranges = nullptr) {
    cudaEventRecord(e.ev0, e.stream);
    const int rc = launch_frames(sp, d_pcm, d_data, d_rets, F, 0, F, frame_size, channels, Fs, max_bytes, stride, d_ranges);
    cudaEventRecord(e.ev1, e.stream);
    return rc;
}

int encode_span_host_locked(OpusEncoder **st, int n, int F, const int16_t *pcm, int frame_size, uint8_t *data, int max_bytes, int stride,
                            int *ret, bool keep_resident, uint32_t *ranges = nullptr) {
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    int rc = check_span(st, n, frame_size);
    if (rc != OPUS_OK) return rc;
    const int channels = st[0]->st.channels;
    if (!e.h_slots.reserve(sizeof(int) * (size_t)n)) return OPUS_ALLOC_FAIL;
    int *hsl = (int *)e.h_slots.p;
    rc = make_resident_locked(st, n, hsl);
    if (rc != OPUS_OK) return rc;
    const size_t NF = (size_t)n * F;
    const size_t pcm_bytes = NF * frame_size * channels * sizeof(int16_t);
    if (!e.d_pcm.reserve(pcm_bytes) || !e.d_data.reserve(NF * stride) || !e.d_rets.reserve(sizeof(int) * NF) ||
        (ranges && !e.d_ranges.reserve(sizeof(uint32_t) * NF)))
        return OPUS_ALLOC_FAIL;
    unsigned *d_ranges = ranges ? (unsigned *)e.d_ranges.p : nullptr;
    SpanSplit sp;
    if (!split_span(st, hsl, n, frame_size, max_bytes, sp)) return OPUS_ALLOC_FAIL;
    const int Fs = st[0]->st.Fs;
    // Large spans are cut into up to kMaxSub sub-spans of frames: the PCM of sub-span k+1 goes up and the packets of sub-span
    // k-1 come down on the copy stream while sub-span k is coded (the state stays resident between the launches).
    int nsub = 1;
    if (pcm_bytes >= ((size_t)32 << 20)) nsub = F / 8 < 1 ? 1 : (F / 8 > 5 ? 5 : F / 8);
    const size_t row_pcm = (size_t)frame_size * channels * sizeof(int16_t);   // one frame of one stream
    if (nsub == 1) {
        cudaMemcpyAsync(e.d_pcm.p, pcm, pcm_bytes, cudaMemcpyHostToDevice, e.stream);
        rc = launch_span(sp, (const int16_t *)e.d_pcm.p, (uint8_t *)e.d_data.p, (int *)e.d_rets.p, F, frame_size, channels, Fs, max_bytes, stride, d_ranges);
        if (rc != OPUS_OK) return rc;
        cudaMemcpyAsync(data, e.d_data.p, NF * stride, cudaMemcpyDeviceToHost, e.stream);
    } else {
        cudaEventRecord(e.ev_prev, e.stream);                 // earlier work on the device buffers (previous call) is done
        cudaStreamWaitEvent(e.copy_stream, e.ev_prev, 0);
        const int per = (F + nsub - 1) / nsub;
        for (int k = 0; k < nsub; k++) {
            const int f0 = k * per, f1 = hmin(F, f0 + per);
            cudaMemcpy2DAsync((char *)e.d_pcm.p + f0 * row_pcm, F * row_pcm, (const char *)pcm + f0 * row_pcm, F * row_pcm, (f1 - f0) * row_pcm,
                              (size_t)n, cudaMemcpyHostToDevice, e.copy_stream);
            cudaEventRecord(e.ev_in[k], e.copy_stream);
        }
        cudaEventRecord(e.ev0, e.stream);
        for (int k = 0; k < nsub; k++) {
            const int f0 = k * per, f1 = hmin(F, f0 + per);
            cudaStreamWaitEvent(e.stream, e.ev_in[k], 0);
            rc = launch_frames(sp, (const int16_t *)e.d_pcm.p, (uint8_t *)e.d_data.p, (int *)e.d_rets.p, F, f0, f1, frame_size, channels, Fs, max_bytes,
                               stride, d_ranges);
            if (rc != OPUS_OK) return rc;
            cudaEventRecord(e.ev_done[k], e.stream);
            cudaStreamWaitEvent(e.copy_stream, e.ev_done[k], 0);
            cudaMemcpy2DAsync(data + (size_t)f0 * stride, (size_t)F * stride, (const uint8_t *)e.d_data.p + (size_t)f0 * stride, (size_t)F * stride,
                              (size_t)(f1 - f0) * stride, (size_t)n, cudaMemcpyDeviceToHost, e.copy_stream);
        }
        cudaEventRecord(e.ev1, e.stream);
    }
    cudaMemcpyAsync(ret, e.d_rets.p, sizeof(int) * NF, cudaMemcpyDeviceToHost, e.stream);
    if (ranges) cudaMemcpyAsync(ranges, d_ranges, sizeof(uint32_t) * NF, cudaMemcpyDeviceToHost, e.stream);
    mark_device_newer_locked(st, n);
    if (cudaStreamSynchronize(e.stream) != cudaSuccess || cudaStreamSynchronize(e.copy_stream) != cudaSuccess) {
        fprintf(stderr, "concentus_b200: CUDA failure in encode span: %s\n", cudaGetErrorString(cudaGetLastError()));
        return OPUS_INTERNAL_ERROR;
    }
    cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1);
    e.total_ms += e.last_ms;
    if (!keep_resident) return sync_states_locked(st, n, true);
    return OPUS_OK;
}

}  // namespace

extern "C" {

long long opus_b200_enc_kernel_launches(void) { return e.launches; }
// streams x calls coded by the frame-synchronous pipeline / by the one-kernel path since start (tests check which path ran)
void opus_b200_enc_path_counts(long long *pipe, long long *legacy) {
    if (pipe) *pipe = e.pipe_streams;
    if (legacy) *legacy = e.legacy_streams;
}
// split band loop statistics (opus_enc_pipe.cu): leaves the exact chain searched itself / leaves the speculative chain listed
void opus_b200_enc_band_stats(long long *misses, long long *leaves) {
    std::lock_guard<std::mutex> lk(e.mu);
    if (e.ok) cudaDeviceSynchronize();
    enc_pipe_stats(misses, leaves);
}
// 1 (default): streams the pipeline takes go through it; 0: everything through the one-kernel path.  Returns the previous setting.
int opus_b200_enc_set_pipeline(int on) {
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    const int prev = e.use_pipe;
    e.use_pipe = on != 0;
    return prev;
}
#if defined(CB_PHASE_PROF)
// dev build only: cycles per phase slot of stream 0 (own work, then barrier wait); reset != 0 clears the counters
int opus_b200_enc_phase_prof(long long *cycles, long long *wait, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(cycles, cb::g_phase_cycles, sizeof(long long) * 64);
    cudaMemcpyFromSymbol(wait, cb::g_phase_wait, sizeof(long long) * 64);
    if (reset) {
        long long z[64] = {0};
        int zi = 0;
        cudaMemcpyToSymbol(cb::g_phase_cycles, z, sizeof(z));
        cudaMemcpyToSymbol(cb::g_phase_wait, z, sizeof(z));
        cudaMemcpyToSymbol(cb::g_phase_idx, &zi, sizeof(zi));
    }
    return 0;
}
#endif
void *opus_b200_enc_stream(void) {
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return nullptr;
    return (void *)e.stream;
}
float opus_b200_enc_last_kernel_ms(void) {
    std::lock_guard<std::mutex> lk(e.mu);
    return e.last_ms;
}
int opus_b200_enc_synchronize(void) {
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    if (cudaStreamSynchronize(e.stream) != cudaSuccess) return OPUS_INTERNAL_ERROR;
    if (e.launches > 0) cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1);
    return OPUS_OK;
}

// ---- lifecycle (opus_encoder.c:150-252,482-510) ----
int opus_encoder_get_size(int channels) {
    if (channels < 1 || channels > 2) return 0;
    return (int)sizeof(OpusEncoder);
}
int opus_encoder_init(OpusEncoder *st, opus_int32 Fs, int channels, int application) {
    if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2) ||
        (application != OPUS_APPLICATION_VOIP && application != OPUS_APPLICATION_AUDIO && application != OPUS_APPLICATION_RESTRICTED_LOWDELAY))
        return OPUS_BAD_ARG;
    {   // re-initialising a live block in place: give its pool slot back first
        std::lock_guard<std::mutex> lk(e.mu);
        if (e.ok && st->magic == kEncMagic && st->slot >= 0) release_slot_locked(st);
    }
    memset(st, 0, sizeof(OpusEncoder));
    st->magic = kEncMagic;
    st->slot = -1;
    st->gen = 0;
    st->host_current = 1;
    st->device = -1;
    if (enc_state_init(&st->st, Fs, channels, application) != 0) return OPUS_BAD_ARG;
    return OPUS_OK;
}
OpusEncoder *opus_encoder_create(opus_int32 Fs, int channels, int application, int *error) {
    if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2) ||
        (application != OPUS_APPLICATION_VOIP && application != OPUS_APPLICATION_AUDIO && application != OPUS_APPLICATION_RESTRICTED_LOWDELAY)) {
        if (error) *error = OPUS_BAD_ARG;
        return nullptr;
    }
    OpusEncoder *st = (OpusEncoder *)malloc(sizeof(OpusEncoder));
    if (!st) {
        if (error) *error = OPUS_ALLOC_FAIL;
        return nullptr;
    }
    int ret = opus_encoder_init(st, Fs, channels, application);
    if (error) *error = ret;
    if (ret != OPUS_OK) {
        free(st);
        st = nullptr;
    }
    return st;
}
void opus_encoder_destroy(OpusEncoder *st) {
    if (!st) return;
    {
        std::lock_guard<std::mutex> lk(e.mu);
        if (e.ok && st->magic == kEncMagic) release_slot_locked(st);
    }
    free(st);
}

// opus_encoder_ctl (opus_encoder.c:2031-2507): every supported request carries one opus_int32 or one pointer to it
int opus_encoder_ctl(OpusEncoder *st, int request, ...) {
    va_list ap;
    va_start(ap, request);
    {
        std::lock_guard<std::mutex> lk(e.mu);
        if (!st->host_current) {
            if (!ctx_init_locked()) { va_end(ap); return OPUS_INTERNAL_ERROR; }
            int rc = make_host_current_locked(st);
            if (rc != OPUS_OK) { va_end(ap); return rc; }
        }
    }
    int ret;
    if (request == 10024) {
        // OPUS_SET_LFE (celt.h:113, opus_encoder.c:2455): the multistream encoder's switch for a low-frequency-effects channel.  The
        // default (0) is accepted; the LFE analysis path itself is part of the multistream scope that is not built.
        ret = va_arg(ap, opus_int32) == 0 ? OPUS_OK : OPUS_UNIMPLEMENTED;
    } else if (request == 10026) {
        // OPUS_SET_ENERGY_MASK (celt.h:116, opus_encoder.c:2462): a surround masking curve; only "none" (NULL) is accepted
        ret = va_arg(ap, void *) == nullptr ? OPUS_OK : OPUS_UNIMPLEMENTED;
    } else if (request == OPUS_RESET_STATE) {
        ret = enc_ctl(&st->st, request, 0, nullptr);
    } else if (enc_ctl_is_get(request)) {
        opus_int32 *p = va_arg(ap, opus_int32 *);
        if (!p) ret = OPUS_BAD_ARG;
        else {
            int v = 0;
            ret = enc_ctl(&st->st, request, 0, &v);
            if (ret == OPUS_OK) *p = v;
        }
    } else {
        opus_int32 v = va_arg(ap, opus_int32);
        int dummy = 0;
        ret = enc_ctl(&st->st, request, v, &dummy);
    }
    va_end(ap);
    return ret;
}

// ---- encode ----
int opus_encode_span_device(OpusEncoder **st, int n, int F, const opus_int16 *d_pcm, int frame_size, unsigned char *d_data,
                            opus_int32 max_data_bytes, opus_int32 *d_ret) {
    if (!st || n <= 0 || F <= 0 || frame_size <= 0 || max_data_bytes <= 0) return OPUS_BAD_ARG;
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    int rc = check_span(st, n, frame_size);
    if (rc != OPUS_OK) return rc;
    if (!e.h_slots.reserve(sizeof(int) * (size_t)n)) return OPUS_ALLOC_FAIL;
    int *hsl = (int *)e.h_slots.p;
    rc = make_resident_locked(st, n, hsl);
    if (rc != OPUS_OK) return rc;
    SpanSplit sp;
    if (!split_span(st, hsl, n, frame_size, hmin(max_data_bytes, 1276), sp)) return OPUS_ALLOC_FAIL;
    rc = launch_span(sp, d_pcm, d_data, d_ret, F, frame_size, st[0]->st.channels, st[0]->st.Fs, hmin(max_data_bytes, 1276), max_data_bytes);
    if (rc != OPUS_OK) return rc;
    mark_device_newer_locked(st, n);
    if (cudaGetLastError() != cudaSuccess) return OPUS_INTERNAL_ERROR;
    return OPUS_OK;
}

int opus_encode_span(OpusEncoder **st, int n, int F, const opus_int16 *pcm, int frame_size, unsigned char *data, opus_int32 max_data_bytes,
                     opus_int32 *ret) {
    if (!st || n <= 0 || F <= 0 || frame_size <= 0 || max_data_bytes <= 0 || !pcm || !data || !ret) return OPUS_BAD_ARG;
    std::lock_guard<std::mutex> lk(e.mu);
    return encode_span_host_locked(st, n, F, pcm, frame_size, data, hmin(max_data_bytes, 1276), max_data_bytes, ret, true);
}

int opus_encode_span_ranges(OpusEncoder **st, int n, int F, const opus_int16 *pcm, int frame_size, unsigned char *data,
                            opus_int32 max_data_bytes, opus_int32 *ret, opus_uint32 *final_range) {
    if (!st || n <= 0 || F <= 0 || frame_size <= 0 || max_data_bytes <= 0 || !pcm || !data || !ret || !final_range) return OPUS_BAD_ARG;
    std::lock_guard<std::mutex> lk(e.mu);
    return encode_span_host_locked(st, n, F, pcm, frame_size, data, hmin(max_data_bytes, 1276), max_data_bytes, ret, true, final_range);
}

int opus_encode_batch(OpusEncoder **st, const opus_int16 *const *pcm, int frame_size, unsigned char *const *data, opus_int32 max_data_bytes,
                      opus_int32 *ret, int n) {
    if (!st || !pcm || !data || !ret || n <= 0) return OPUS_BAD_ARG;
    if (frame_size <= 0 || max_data_bytes <= 0) {
        for (int i = 0; i < n; i++) ret[i] = OPUS_BAD_ARG;
        return OPUS_OK;
    }
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    const int stride = hmin(max_data_bytes, 1276);
    for (int i = 0; i < n; i++)
        if (!st[i] || st[i]->magic != kEncMagic || !pcm[i] || !data[i]) ret[i] = OPUS_BAD_ARG;
    // group by (Fs, channels): the span kernel wants uniform rows
    std::vector<char> done(n, 0);
    for (int i = 0; i < n; i++) {
        if (done[i] || !st[i] || st[i]->magic != kEncMagic || !pcm[i] || !data[i]) continue;
        std::vector<int> idx;
        for (int j = i; j < n; j++)
            if (!done[j] && st[j] && st[j]->magic == kEncMagic && pcm[j] && data[j] && st[j]->st.channels == st[i]->st.channels &&
                st[j]->st.Fs == st[i]->st.Fs) {
                idx.push_back(j);
                done[j] = 1;
            }
        const int m = (int)idx.size(), ch = st[i]->st.channels;
        // OPUS_SET_EXPERT_FRAME_DURATION: the frame actually coded (opus_encode -> compute_frame_size, opus_encoder.c:2007-2025).
        // Streams of one group that select different sizes are launched separately.
        std::vector<int> fsz(m);
        for (int k = 0; k < m; k++) {
            if (!st[idx[k]]->host_current && st[idx[k]]->st.variable_duration != 5000) { /* config never changes on the device */ }
            fsz[k] = frame_size_select(frame_size, st[idx[k]]->st.variable_duration, st[idx[k]]->st.Fs);
        }
        std::vector<char> sub_done(m, 0);
        for (int k0 = 0; k0 < m; k0++) {
            if (sub_done[k0]) continue;
            if (fsz[k0] < 0) { ret[idx[k0]] = OPUS_BAD_ARG; sub_done[k0] = 1; continue; }
            std::vector<int> sub;
            for (int k = k0; k < m; k++)
                if (!sub_done[k] && fsz[k] == fsz[k0]) { sub.push_back(idx[k]); sub_done[k] = 1; }
            const int q = (int)sub.size(), fs = fsz[k0];
            std::vector<OpusEncoder *> sts(q);
            std::vector<int16_t> in((size_t)q * fs * ch);
            std::vector<uint8_t> out((size_t)q * stride);
            std::vector<int> rets(q);
            for (int k = 0; k < q; k++) {
                sts[k] = st[sub[k]];
                memcpy(in.data() + (size_t)k * fs * ch, pcm[sub[k]], (size_t)fs * ch * sizeof(int16_t));
            }
            int rc = encode_span_host_locked(sts.data(), q, 1, in.data(), fs, out.data(), stride, stride, rets.data(), true);
            if (rc != OPUS_OK) return rc;
            for (int k = 0; k < q; k++) {
                ret[sub[k]] = rets[k];
                if (rets[k] > 0) memcpy(data[sub[k]], out.data() + (size_t)k * stride, (size_t)rets[k]);
            }
        }
    }
    return OPUS_OK;
}

int opus_encoder_sync(OpusEncoder **st, int n) {
    if (!st || n <= 0) return OPUS_BAD_ARG;
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    return sync_states_locked(st, n, true);
}

// Scalar call = batch of one; the host block is left current so it stays memcpy-able (opus_encoder.c:2007-2025).
opus_int32 opus_encode(OpusEncoder *st, const opus_int16 *pcm, int analysis_frame_size, unsigned char *data, opus_int32 max_data_bytes) {
    if (!st || st->magic != kEncMagic || !pcm || !data) return OPUS_BAD_ARG;
    if (max_data_bytes <= 0) return OPUS_BAD_ARG;
    std::lock_guard<std::mutex> lk(e.mu);
    if (!ctx_init_locked()) return OPUS_INTERNAL_ERROR;
    // the ctl-visible configuration never changes on the device, so a stale host block still holds it
    const int frame_size = frame_size_select(analysis_frame_size, st->st.variable_duration, st->st.Fs);
    if (frame_size < 0) return OPUS_BAD_ARG;
    const int stride = hmin(max_data_bytes, 1276);
    std::vector<uint8_t> out((size_t)stride);
    int r = 0;
    OpusEncoder *one = st;
    int rc = encode_span_host_locked(&one, 1, 1, pcm, frame_size, out.data(), stride, stride, &r, false);
    if (rc != OPUS_OK) return rc;
    if (r > 0) memcpy(data, out.data(), (size_t)r);
    return r;
}

}  // extern "C"
