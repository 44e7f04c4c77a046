// opus_enc_pipe.cu — kernels and host orchestration of the frame-synchronous encoder pipeline (celt_enc_pipe.cuh has the design
// and the per-stage device functions; enc_pipe_host.h is the interface opus_enc_capi.cu calls).
//
// A launch's streams advance frame by frame through K1..K5 on the main CUDA stream, with the front end (P0 / FE1 / FE2) of the
// NEXT chunk of frames on a side stream.  Every frame-step kernel is ONE wave holding all streams of the launch (4,096 streams =
// 27.7 warps per SM at 72 registers = the whole register file), so a kernel's duration is a stream's dependent-instruction chain
// and the span's time is the sum of its kernels — measured: work on the side stream costs its full stand-alone time, stream groups
// on separate CUDA streams (CB200_ENC_GROUPS) and per-stream dataflow launches (CB200_ENC_FLOW) are slower
// (profiles/r2_encoder_ab.md).  What pays is a shorter chain per kernel: the knobs below default to what measured best.
#if !defined(CB_PIPE_BIG_CODE)
#define CB_SMALL_CODE 1   // celt_simt.cuh: medium helpers as real calls, loops not unrolled (A/B: -DCB_PIPE_BIG_CODE)
#if !defined(CB_PIPE_NO_TINY)
#define CB_TINY_CODE 1    // ... and the range coder's renormalisation / tell / the rotation chain as shared calls
#endif
#if !defined(CB_PIPE_NO_ROT_LUT)
#define CB_ROT_LUT 1      // celt_enc_bands.cuh: the spreading rotation's parameters from tables filled at start-up
#endif
#endif
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "celt_enc_pipe.cuh"
#include "enc_pipe_host.h"
#include "host_runtime.h"

using namespace cb;

#ifndef CB_PIPE_WPB
#define CB_PIPE_WPB 4          // warps per block of the warp-per-item kernels
#endif

namespace {

struct PipeBufs {              // device views of one group's buffers
    int16_t *D;                // [n][Fc][fsz*CC]
    int *P[2];                 // [n][CC][pstride], double-buffered over chunks
    EncPlan *plans[2];         // [n][Fc]
    FeFrame *fe[2];            // [n][Fc]
    int *m0;                   // [n][2] pre-emphasis memory entering the chunk
    EncPipeCtx *ctx;           // [n]
    EncPipeBuf *buf;           // [n]
};

// ---- per-stream dataflow between the kernels of a frame step ---------------------------------------------------------------------
// Stream-ordered launches make every kernel wait for the SLOWEST stream of its predecessor (a stream's cost varies with its content:
// the average warp of the band walk is done at 78 % of the kernel's time), and the thread-per-stream stages leave the machine idle.
// With Flow the kernels of a chunk are launched with programmatic stream serialisation: a kernel's blocks may start as soon as every
// block of the previous kernel has STARTED, and what a stream's stage really needs — the previous stage of the SAME stream — is
// tracked in a per-stream progress counter: a stage waits until the counter shows its predecessor done (acquire) and adds its own
// unit when its results are written (fence + atomic).  Every earlier kernel is fully resident by the time a block can spin, so the
// spin always ends; it is bounded anyway, and a time-out only raises a flag the host turns into OPUS_INTERNAL_ERROR.
struct Flow {
    int *prog;       // [n] units completed per stream since the call began; nullptr: plain stream-ordered launches
    int need;        // units that must be complete before this stage may touch the stream
    int *err;        // raised when a wait timed out (never in a correct run)
    int open;        // let the next kernel of the stream start early (griddepcontrol.launch_dependents)
};
__device__ __forceinline__ void flow_open(const Flow &fl) {
#if defined(__CUDA_ARCH__)
    if (fl.open) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void flow_wait_one(const Flow &fl, int s) {
    const int *p = fl.prog + s;
    int v, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= fl.need) break;
        __nanosleep(200);
        if (++spins > (1 << 21) || *(volatile int *)fl.err) { atomicExch(fl.err, 1); break; }
    }
}
// warp-per-item kernels: lane 0 waits for the warp's stream
__device__ __forceinline__ void flow_wait_warp(const Flow &fl, int s) {
    if (!fl.prog) return;
    if ((threadIdx.x & 31) == 0) flow_wait_one(fl, s);
    __syncwarp();
    __threadfence();
}
__device__ __forceinline__ void flow_done_warp(const Flow &fl, int s) {
    if (!fl.prog) return;
    __threadfence();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicAdd(fl.prog + s, 1);
}
// block-per-group-of-streams kernels: thread k waits for / signals stream s0 + k
__device__ __forceinline__ void flow_wait_block(const Flow &fl, int s0, int nvalid) {
    if (!fl.prog) return;
    if ((int)threadIdx.x < nvalid) flow_wait_one(fl, s0 + threadIdx.x);
    __syncthreads();
    __threadfence();
}
__device__ __forceinline__ void flow_done_block(const Flow &fl, int s0, int nvalid) {
    if (!fl.prog) return;
    __threadfence();
    __syncthreads();
    if ((int)threadIdx.x < nvalid) atomicAdd(fl.prog + s0 + threadIdx.x, 1);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- staging for the thread-per-stream kernels ------------------------------------------------------------------------------------
// A scalar stage walks a few KB of per-stream data with dependent accesses; from global memory every one of them costs an L2 round
// trip (the first version of K4 ran 21 cycles per instruction).  So a block takes 32 streams: all its 128 threads copy the streams'
// structs into shared memory (coalesced, several loads in flight), the stage runs on the copies and the block writes back what the
// stage changed.  The stages are divergent (every stream is somewhere else in its own partition tree), and a warp serialises its
// divergent lanes, so a stage's latency grows with the streams per warp: only L (1, 2 or 4, CB200_ENC_SCALAR_L) lanes of a warp
// take a stream each.  The machine is nowhere near issue bound in these stages; the idle lanes cost nothing that matters.  Struct strides are chosen against bank conflicts: an odd number of words, or
// 2 * odd where the struct holds an 8-byte pointer.
enum { kScalarThreads = 256 };   // 8 warps; a block takes T = 8 * L streams, L streams per warp (lanes 0..L-1 work)
constexpr int odd_stride(int words) { return words | 1; }
constexpr int even_odd_stride(int words) { return ((words + 1) / 2 % 2 == 1) ? (words + 1) / 2 * 2 : (words + 1) / 2 * 2 + 2; }
constexpr int kHeadWords = CB_ENC_HEAD_BYTES / 4, kHeadStride = odd_stride(kHeadWords);
constexpr int kCtxWords = (int)((sizeof(EncPipeCtx) + 3) / 4), kCtxStride = even_odd_stride(kCtxWords);
constexpr int kPrepWords = (int)((sizeof(BandPrep) + 3) / 4), kPrepStride = odd_stride(kPrepWords);
constexpr int kLeafWords = (int)((sizeof(LeafList) + 3) / 4), kLeafStride = odd_stride(kLeafWords);
constexpr int kHeadCtxWords = (int)(offsetof(EncPipeCtx, spread_sum) / 4);   // what K1 produces
constexpr int kLeafTaskWords = (int)(offsetof(LeafList, index) / 4);   // what chain-S produces
static_assert(sizeof(EncPipeCtx) % 8 == 0 && sizeof(BandPrep) % 4 == 0 && sizeof(LeafList) % 4 == 0, "struct sizes");

__device__ __forceinline__ void stage_copy_in(int *sm, int stride, int words, int *const *ptrs, int nvalid) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < nvalid; k += kScalarThreads / 32) {
        const int *src = ptrs[k];
        int *dst = sm + k * stride;
#pragma unroll 4
        for (int w = lane; w < words; w += 32) dst[w] = src[w];
    }
}
// Of the state head a frame step owns [rangeFinal, preemph_memE) and [vbr_reservoir, end of head): the Opus-layer fields, hp_mem and
// preemph_memE belong to the prepass, which is working on the NEXT chunk on another stream at the same time — a frame step that
// wrote its whole copy of the head back would undo the prepass's updates.
constexpr int kHeadOwn0 = (int)(offsetof(CbEncState, rangeFinal) / 4), kHeadOwn1 = (int)(offsetof(CbEncState, preemph_memE) / 4);
constexpr int kHeadOwn2 = (int)(offsetof(CbEncState, vbr_reservoir) / 4);
__device__ __forceinline__ void stage_copy_out_head(const int *sm, int *const *ptrs, int nvalid) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < nvalid; k += kScalarThreads / 32) {
        int *dst = ptrs[k];
        const int *src = sm + k * kHeadStride;
        for (int w = kHeadOwn0 + lane; w < kHeadWords; w += 32)
            if (w < kHeadOwn1 || w >= kHeadOwn2) dst[w] = src[w];
    }
}
__device__ __forceinline__ void stage_copy_out(const int *sm, int stride, int words, int *const *ptrs, int nvalid) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < nvalid; k += kScalarThreads / 32) {
        int *dst = ptrs[k];
        const int *src = sm + k * stride;
#pragma unroll 4
        for (int w = lane; w < words; w += 32) dst[w] = src[w];
    }
}

// ---- P0: the Opus layer of a chunk --------------------------------------------------------------------------------------------
// A block takes 32 streams.  The sample pass (compute_stereo_width sums, dc_reject, stereo_fade: all order dependent along time,
// independent across streams and — but for the pointwise fade — channels) runs with one lane per (stream, channel) on tiles of
// kPreTile samples that the whole block moves between global and shared memory with coalesced 32-bit accesses; a thread walking
// its own row in global memory cost 32 transactions per load (first version: 1 ms per frame step).
enum { kPreTile = 64, kPreRow = kPreTile * 2 + 2 };   // int16 per stream row of a tile: an odd number of words
__global__ void __launch_bounds__(128)
pipe_prepass_kernel(CbEncState *pool, const int *slots, const int *sidx, PipeGeom g, const int16_t *pcm, int fbase, int nfr, int16_t *D,
                    EncPlan *plans, int *m0out) {
    __shared__ __align__(16) int16_t tile[32][kPreRow];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * 32;
    const int nvalid = g.n - t0 < 32 ? g.n - t0 : 32;
    const int CC = g.CC, fsz = g.fsz;
    const size_t row = (size_t)fsz * CC;
    // worker threads: tid < 64, stream ws = tid >> 1, channel wc = tid & 1
    const int ws = tid >> 1, wc = tid & 1;
    const bool worker = tid < 64;
    const bool live = worker && ws < nvalid && wc < CC;
    CbEncState *st = worker && ws < nvalid ? pool + slots[t0 + ws] : nullptr;
    int hm0 = 0, hm1 = 0, last_v = 0, any_coded = 0;
    if (live) {
        hm0 = st->hp_mem[2 * wc];
        hm1 = st->hp_mem[2 * wc + 1];
        m0out[2 * (t0 + ws) + wc] = st->preemph_memE[wc];
    } else if (worker && ws < nvalid) {
        m0out[2 * (t0 + ws) + wc] = 0;
    }
    for (int fi = 0; fi < nfr; fi++) {
        // ---- decisions: the channel-0 lane of every stream, shared with its partner ----
        PlanPre pp;
        pp.code = 0; pp.ret = 0; pp.want_width = 0; pp.fade = 0; pp.g1 = 0; pp.g2 = 0; pp.dc_shift = 0;
        if (worker && ws < nvalid && wc == 0) pipe_plan_pre(st, fsz, g.max_bytes, pp);
        int code = pp.code, want_width = pp.want_width, fade = pp.fade, fg1 = pp.g1, fg2 = pp.g2, dshift = pp.dc_shift;
        if (worker) {
            code = __shfl_sync(0xffffffffu, code, lane & ~1);
            want_width = __shfl_sync(0xffffffffu, want_width, lane & ~1);
            fade = __shfl_sync(0xffffffffu, fade, lane & ~1);
            fg1 = __shfl_sync(0xffffffffu, fg1, lane & ~1);
            fg2 = __shfl_sync(0xffffffffu, fg2, lane & ~1);
            dshift = __shfl_sync(0xffffffffu, dshift, lane & ~1);
        }
        int acc_a = 0, acc_b = 0;      // channel 0: xx, xy; channel 1: yy
        int part_a = 0, part_b = 0;
        for (int pos = 0; pos < fsz; pos += kPreTile) {
            const int nt = fsz - pos < kPreTile ? fsz - pos : kPreTile;
            const int words = nt * CC / 2;   // fsz * CC is even for every Opus frame size
            // ---- tile in: warp w moves streams w, w+4, ... ----
            for (int k = warp; k < nvalid; k += 4) {
                const int *src = reinterpret_cast<const int *>(pcm + ((size_t)sidx[t0 + k] * g.F + fbase + fi) * row + (size_t)pos * CC);
                int *dst = reinterpret_cast<int *>(&tile[k][0]);
                for (int w = lane; w < words; w += 32) dst[w] = src[w];
            }
            __syncthreads();
            // ---- the samples, one lane per (stream, channel) ----
            if (worker) {
                for (int i = 0; i < nt; i++) {
                    const int x = live ? (int)tile[ws][CC * i + wc] : 0;
                    const int xo = __shfl_xor_sync(0xffffffffu, x, 1);
                    int v = 0;
                    if (live && code) {
                        if (want_width) {
                            const int gi = pos + i;
                            if ((gi & ~3) < fsz - 3) {
                                if (wc == 0) { part_a += mul16_16(x, x) >> 2; part_b += mul16_16(x, xo) >> 2; }
                                else part_a += mul16_16(x, x) >> 2;
                                if ((gi & 3) == 3) {
                                    acc_a = wadd(acc_a, part_a >> 10);
                                    acc_b = wadd(acc_b, part_b >> 10);
                                    part_a = part_b = 0;
                                }
                            }
                        }
                        v = dc_reject_step(x, hm0, hm1, dshift);
                    }
                    const int vo = __shfl_xor_sync(0xffffffffu, v, 1);
                    if (live && code) {
                        if (fade) {
                            const int gq = stereo_fade_gain(pos + i, fg1, fg2, g.Fs);
                            const int l = wc == 0 ? v : vo, r = wc == 0 ? vo : v;
                            int diff = s16((l - r) >> 1);
                            diff = mul16_16_q15(gq, diff);
                            v = wc == 0 ? (int)(int16_t)(l - diff) : (int)(int16_t)(r + diff);
                        }
                        tile[ws][CC * i + wc] = (int16_t)v;
                        last_v = v;
                    }
                }
            }
            __syncthreads();
            // ---- tile out ----
            for (int k = warp; k < nvalid; k += 4) {
                int *dst = reinterpret_cast<int *>(D + ((size_t)(t0 + k) * g.Fc + fi) * row + (size_t)pos * CC);
                const int *src = reinterpret_cast<const int *>(&tile[k][0]);
                for (int w = lane; w < words; w += 32) dst[w] = src[w];
            }
            __syncthreads();
        }
        // ---- commit ----
        if (worker) {
            const int yy = __shfl_xor_sync(0xffffffffu, acc_a, 1);
            if (ws < nvalid && wc == 0) {
                EncPlan pl;
                pipe_plan_post(st, pp, fsz, acc_a, acc_b, yy, pl);
                plans[(size_t)(t0 + ws) * g.Fc + fi] = pl;
            }
            if (code) any_coded = 1;
        }
    }
    if (live && any_coded) {
        st->hp_mem[2 * wc] = hm0;
        st->hp_mem[2 * wc + 1] = hm1;
        st->preemph_memE[wc] = g.upsample == 1 ? mul16_16(kPreemphCoef0, last_v) >> 3 : 0;
    }
}

// ---- P0, second version: one THREAD per stream, no tiles ------------------------------------------------------------------------
// The first version staged every 64 samples through shared memory with three block barriers (147 us per frame, and since the front
// end's kernels are NOT hidden behind the frame steps — the machine is full, profiles/r2_encoder_ab.md — all of it was span time).
// A thread now walks its own PCM row with 16-byte loads (a 128-byte line serves eight of them out of L1, the next lines are
// prefetched), runs both channels' dc_reject chains side by side (independent, so they interleave in the pipeline), and stores
// 16-byte vectors.  Needs whole vectors per frame row (every 48 / 24 / 16 / 8 kHz frame size; not 12 kHz 2.5 ms).
__global__ void __launch_bounds__(32)
pipe_prepass2_kernel(CbEncState *pool, const int *slots, const int *sidx, PipeGeom g, const int16_t *pcm, int fbase, int nfr, int16_t *D,
                     EncPlan *plans, int *m0out) {
    const int t = blockIdx.x * 32 + threadIdx.x;
    if (t >= g.n) return;
    CbEncState *st = pool + slots[t];
    const int CC = g.CC, fsz = g.fsz;
    const size_t row = (size_t)fsz * CC;
    int hm0 = st->hp_mem[0], hm1 = st->hp_mem[1], hm2 = 0, hm3 = 0;
    m0out[2 * t] = st->preemph_memE[0];
    if (CC == 2) {
        hm2 = st->hp_mem[2];
        hm3 = st->hp_mem[3];
        m0out[2 * t + 1] = st->preemph_memE[1];
    } else {
        m0out[2 * t + 1] = 0;
    }
    int last_l = 0, last_r = 0, any_coded = 0;
    for (int fi = 0; fi < nfr; fi++) {
        PlanPre pp;
        pp.code = 0; pp.ret = 0; pp.want_width = 0; pp.fade = 0; pp.g1 = 0; pp.g2 = 0; pp.dc_shift = 0;
        pipe_plan_pre(st, fsz, g.max_bytes, pp);
        int xx = 0, xy = 0, yy = 0;
        if (pp.code) {
            const int4 *src = reinterpret_cast<const int4 *>(pcm + ((size_t)sidx[t] * g.F + fbase + fi) * row);
            int4 *dst = reinterpret_cast<int4 *>(D + ((size_t)t * g.Fc + fi) * row);
            const int nv = (int)(row >> 3);          // 16-byte vectors in the row
            const int shift = pp.dc_shift;
            int4 v = src[0];
            if (CC == 2) {
                for (int k = 0; k < nv; k++) {
                    if ((k & 7) == 0) prefetch_l1(src + (k + 16 < nv ? k + 16 : nv - 1));
                    const int4 nx = src[k + 1 < nv ? k + 1 : k];
                    const int w[4] = {v.x, v.y, v.z, v.w};
                    int o[4];
                    int pxx = 0, pxy = 0, pyy = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int xl = s16(w[j]), xr = w[j] >> 16;
                        if (pp.want_width) {
                            pxx += mul16_16(xl, xl) >> 2;
                            pxy += mul16_16(xl, xr) >> 2;
                            pyy += mul16_16(xr, xr) >> 2;
                        }
                        int vl = dc_reject_step(xl, hm0, hm1, shift);
                        int vr = dc_reject_step(xr, hm2, hm3, shift);
                        if (pp.fade) {
                            const int gq = stereo_fade_gain(4 * k + j, pp.g1, pp.g2, g.Fs);
                            int diff = s16((vl - vr) >> 1);
                            diff = mul16_16_q15(gq, diff);
                            const int l = vl, r = vr;
                            vl = (int)(int16_t)(l - diff);
                            vr = (int)(int16_t)(r + diff);
                        }
                        o[j] = (vl & 0xffff) | (vr << 16);
                        last_l = vl;
                        last_r = vr;
                    }
                    if (pp.want_width) {
                        xx = wadd(xx, pxx >> 10);
                        xy = wadd(xy, pxy >> 10);
                        yy = wadd(yy, pyy >> 10);
                    }
                    dst[k] = make_int4(o[0], o[1], o[2], o[3]);
                    v = nx;
                }
            } else {
                for (int k = 0; k < nv; k++) {
                    if ((k & 7) == 0) prefetch_l1(src + (k + 16 < nv ? k + 16 : nv - 1));
                    const int4 nx = src[k + 1 < nv ? k + 1 : k];
                    const int w[4] = {v.x, v.y, v.z, v.w};
                    int o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int a = dc_reject_step(s16(w[j]), hm0, hm1, shift);
                        const int b = dc_reject_step(w[j] >> 16, hm0, hm1, shift);
                        o[j] = (a & 0xffff) | (b << 16);
                        last_l = b;
                    }
                    dst[k] = make_int4(o[0], o[1], o[2], o[3]);
                    v = nx;
                }
            }
            any_coded = 1;
        }
        EncPlan pl;
        pipe_plan_post(st, pp, fsz, xx, xy, yy, pl);
        plans[(size_t)t * g.Fc + fi] = pl;
    }
    if (any_coded) {
        st->hp_mem[0] = hm0;
        st->hp_mem[1] = hm1;
        st->preemph_memE[0] = g.upsample == 1 ? mul16_16(kPreemphCoef0, last_l) >> 3 : 0;
        if (CC == 2) {
            st->hp_mem[2] = hm2;
            st->hp_mem[3] = hm3;
            st->preemph_memE[1] = g.upsample == 1 ? mul16_16(kPreemphCoef0, last_r) >> 3 : 0;
        }
    }
}

// ---- FE1: pre-emphasis + maxima, one warp per (stream, frame); the warp of frame 0 also installs the 1024-sample history ----
__global__ void __launch_bounds__(CB_PIPE_WPB * 32)
pipe_fe1_kernel(const CbEncState *pool, const int *slots, PipeGeom g, int nfr, const int16_t *D, const EncPlan *plans, const int *m0, int *P,
                const int *Pprev, int prev_nfr, FeFrame *fe) {
    const int w = blockIdx.x * CB_PIPE_WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= g.n * nfr) return;
    const int s = w / nfr, fi = w - s * nfr;
    FreeWarpTeam tm{{lane}};
    int *Prow = P + (size_t)s * g.CC * g.pstride;
    if (fi == 0) {
        for (int c = 0; c < g.CC; c++) {
            int *h = Prow + c * g.pstride;
            if (Pprev) {
                const int *src = Pprev + ((size_t)s * g.CC + c) * g.pstride + prev_nfr * g.N;
                for (int i = lane; i < kPipeHist; i += 32) h[i] = src[i];
            } else {
                const int *src = pool[slots[s]].prefilter_mem + c * kCombMaxPeriod;
                for (int i = lane; i < kPipeHist; i += 32) h[i] = src[i];
            }
        }
    }
    const EncPlan &pl = plans[(size_t)s * g.Fc + fi];
    if (!pl.code) return;
    const size_t row = (size_t)g.fsz * g.CC;
    const int16_t *d = D + ((size_t)s * g.Fc + fi) * row;
    int mi[2];
    for (int c = 0; c < g.CC; c++) mi[c] = fi == 0 ? m0[2 * s + c] : pipe_preemph_mem_after(g, d - row, c);
    pipe_preemph_frame(tm, g, pl, d, Prow, fi, mi, fe[(size_t)s * g.Fc + fi]);
}

// ---- FE2: pitch analysis, one warp per (stream, frame) ---------------------------------------------------------------------
__global__ void __launch_bounds__(CB_PIPE_WPB * 32)
pipe_fe2_kernel(PipeGeom g, int nfr, const EncPlan *plans, const int *P, FeFrame *fe) {
    __shared__ PitchScratch sm[CB_PIPE_WPB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * CB_PIPE_WPB + wib;
    if (w >= g.n * nfr) return;
    const int s = w / nfr, fi = w - s * nfr;
    FreeWarpTeam tm{{lane}};
    const int *Prow = P + (size_t)s * g.CC * g.pstride;
    pipe_pitch_frame(tm, g, plans[(size_t)s * g.Fc + fi], Prow + fi * g.N, Prow + g.pstride + fi * g.N, sm[wib], fe[(size_t)s * g.Fc + fi]);
}

// ---- K1 ---------------------------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kScalarThreads)
pipe_head_kernel(CbEncState *pool, const int *slots, const int *sidx, PipeGeom g, int f, int fi, const EncPlan *plans, const FeFrame *fe,
                 EncPipeCtx *ctx, uint8_t *data, Flow fl) {
    extern __shared__ __align__(16) int sm[];
    constexpr int kScalarT = 8 * L;
    __shared__ int *p_head[kScalarT], *p_ctx[kScalarT];
    int *sm_head = sm, *sm_ctx = sm + kScalarT * kHeadStride + (kScalarT * kHeadStride & 1);
    const int t0 = blockIdx.x * kScalarT;
    const int tid = (threadIdx.x & 31) < L ? (int)(threadIdx.x >> 5) * L + (int)(threadIdx.x & 31) : kScalarT;   // the stream this thread works on, if any
    const int nvalid = g.n - t0 < kScalarT ? g.n - t0 : kScalarT;
    flow_open(fl);
    flow_wait_block(fl, t0, nvalid);
    if (tid < nvalid) { p_head[tid] = reinterpret_cast<int *>(pool + slots[t0 + tid]); p_ctx[tid] = reinterpret_cast<int *>(ctx + t0 + tid); }
    __syncthreads();
    stage_copy_in(sm_head, kHeadStride, kHeadWords, p_head, nvalid);
    __syncthreads();
    if (tid < nvalid) {
        const int t = t0 + tid;
        uint8_t *out = data + ((size_t)sidx[t] * g.F + f) * g.stride;
        pipe_head(reinterpret_cast<CbEncState *>(sm_head + tid * kHeadStride), g, plans[(size_t)t * g.Fc + fi], fe[(size_t)t * g.Fc + fi],
                  *reinterpret_cast<EncPipeCtx *>(sm_ctx + tid * kCtxStride), out);
    }
    __syncthreads();
    stage_copy_out_head(sm_head, p_head, nvalid);
    stage_copy_out(sm_ctx, kCtxStride, kHeadCtxWords, p_ctx, nvalid);   // K1 writes the leading scalars of the context only
    flow_done_block(fl, t0, nvalid);
}

// ---- K2a: comb pre-filter, one warp per (stream, channel) -----------------------------------------------------------------------
__global__ void __launch_bounds__(CB_PIPE_WPB * 32)
pipe_comb_kernel(CbEncState *pool, const int *slots, PipeGeom g, int fi, const int *P, EncPipeCtx *ctx, EncPipeBuf *buf, Flow fl) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * CB_PIPE_WPB + wib;
    flow_open(fl);
    if (w >= g.n * g.CC) return;
    const int s = w / g.CC, c = w - s * g.CC;
    flow_wait_warp(fl, s);
    FreeWarpTeam tm{{lane}};
    const int *pre = P + ((size_t)s * g.CC + c) * g.pstride + fi * g.N;
    pipe_comb_channel(tm, pool + slots[s], g, ctx[s], pre, buf[s].in + c * (g.N + kOverlap), c, nullptr, nullptr);
    flow_done_warp(fl, s);
}

// ---- K2b: transient_analysis, one THREAD per (stream, channel) ----------------------------------------------------------------
// The analysis is three recurrences along time (a second-order high-pass with a floor in its feedback, two one-pole followers):
// order dependent within a channel, independent across channels.  With a warp per channel a single lane did all of it (64 K warp
// instructions per stream and frame, the most expensive stage of the pipeline); here a block takes 32 channels, streams them
// through shared-memory tiles (coalesced loads by the whole block) and lane r of warp 0 runs channel r.
enum { kTrTile = 64, kTrTileRow = kTrTile + 1, kTrRowWords = (kMaxFrame + kOverlap) / 2 + 1 };
__global__ void __launch_bounds__(128)
pipe_transient_kernel(PipeGeom g, EncPipeCtx *ctx, const EncPipeBuf *buf) {
    extern __shared__ __align__(16) int smt[];
    int *tile = smt;                                                      // [32][kTrTileRow]
    int16_t *rows = reinterpret_cast<int16_t *>(smt + 32 * kTrTileRow);   // [32][2 * kTrRowWords]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p0 = blockIdx.x * 32;
    const int npairs = g.n * g.CC;
    const int nvalid = npairs - p0 < 32 ? npairs - p0 : 32;
    const int len = g.N + kOverlap;
    const bool worker = warp == 0 && lane < nvalid;
    int ws = 0, wc = 0;
    bool act = false;
    if (worker) {
        ws = (p0 + lane) / g.CC;
        wc = (p0 + lane) - ws * g.CC;
        act = ctx[ws].code && ctx[ws].cfg.complexity >= 1;
    }
    TransientHp hp;
    hp.reset();
    int16_t *row = rows + lane * (2 * kTrRowWords);
    for (int pos = 0; pos < len; pos += kTrTile) {
        const int nt = len - pos < kTrTile ? len - pos : kTrTile;
        for (int k = warp; k < nvalid; k += 4) {
            const int s = (p0 + k) / g.CC, c = (p0 + k) - s * g.CC;
            const int *src = buf[s].in + c * (g.N + kOverlap) + pos;
            for (int i = lane; i < nt; i += 32) tile[k * kTrTileRow + i] = src[i];
        }
        __syncthreads();
        if (worker && act) {
            const int *t = tile + lane * kTrTileRow;
            for (int i = 0; i < nt; i++) row[pos + i] = (int16_t)hp.step(t[i] >> 12, pos + i);
        }
        __syncthreads();
    }
    if (worker && act) ctx[ws].v.mask_metric[wc] = transient_finish_row(row, len, hp.mx, hp.mn);
}

// ---- K2b, second version: the same analysis with the chains kept short ------------------------------------------------------------
// The first version spent 190 us per frame step on 8,192 threads: every step of its recurrences waited for a shared-memory load and
// a tile hand-over.  Here a block still takes 32 channels, but
//   A  lane r of warp 0 runs channel r's high-pass straight from global memory with 16-byte loads one step ahead of the chain
//      (a 128-byte line serves 32 steps out of L1) and stores pairs of int16 results;
//   B  all 128 threads square and add the pairs (four threads per channel) — position-parallel, the mean is a wrapping sum;
//   C  lane r of warp 0 runs the forward and the backward follower on 32-bit words (two values per load), unrolled so that the
//      loads are issued ahead of the chain;
//   D  all threads: the unmasking sum, four threads per channel.
enum { kTr2Row = kMaxFrame + kOverlap + 2, kTr2E = (kMaxFrame + kOverlap) / 2 + 2 };   // int16 strides: 2 * odd words
static_assert((kTr2Row / 2) % 2 == 1 && (kTr2E / 2) % 2 == 1, "row strides must be an odd number of words");
constexpr int kSmemTransient2 = 32 * (kTr2Row + kTr2E) * 2;
__global__ void __launch_bounds__(128)
pipe_transient2_kernel(PipeGeom g, EncPipeCtx *ctx, const EncPipeBuf *buf, Flow fl) {
    extern __shared__ __align__(16) int smt[];
    int16_t *rows = reinterpret_cast<int16_t *>(smt);          // [32][kTr2Row]
    int16_t *Es = rows + 32 * kTr2Row;                          // [32][kTr2E]
    __shared__ int s_mx[32], s_mn[32], s_act[32], s_maxE[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p0 = blockIdx.x * 32;
    const int npairs = g.n * g.CC;
    const int nvalid = npairs - p0 < 32 ? npairs - p0 : 32;
    const int len = g.N + kOverlap, len2 = len / 2;
    flow_open(fl);
    if (fl.prog) {   // thread k waits for the stream of channel p0 + k
        if (tid < nvalid) flow_wait_one(fl, (p0 + tid) / g.CC);
        __syncthreads();
        __threadfence();
    }
    // ---- A ----
    if (warp == 0) {
        bool act = false;
        int mx = 0, mn = 0;
        if (lane < nvalid) {
            const int ws = (p0 + lane) / g.CC, wc = (p0 + lane) - ws * g.CC;
            act = ctx[ws].code && ctx[ws].cfg.complexity >= 1;
            if (act) {
                const int4 *src = reinterpret_cast<const int4 *>(buf[ws].in + wc * len);
                int *row = reinterpret_cast<int *>(rows + lane * kTr2Row);
                int mem0 = 0, mem1 = 0;
                const int nq = len >> 2;
                int4 v = src[0];
#pragma unroll 2
                for (int k = 0; k < nq; k++) {
                    if ((k & 7) == 0) prefetch_l1(src + (k + 16 < nq ? k + 16 : nq - 1));   // two 128-byte lines ahead of the chain
                    const int4 nx = src[k + 1 < nq ? k + 1 : k];
                    int t[4];
                    const int xs[4] = {v.x >> 12, v.y >> 12, v.z >> 12, v.w >> 12};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int x = xs[j];
                        const int y = wadd(mem0, x);
                        mem0 = wsub(wadd(mem1, y), shl32(x, 1));
                        mem1 = wsub(x, y >> 1);
                        t[j] = 4 * k + j < 12 ? 0 : s16(y >> 2);
                        mx = imax(mx, t[j]);
                        mn = imin(mn, t[j]);
                    }
                    row[2 * k] = (t[0] & 0xffff) | (t[1] << 16);
                    row[2 * k + 1] = (t[2] & 0xffff) | (t[3] << 16);
                    v = nx;
                }
            }
        }
        s_act[lane] = act;
        s_mx[lane] = mx;
        s_mn[lane] = mn;
    }
    __syncthreads();
    // ---- B ----
    const int r = tid >> 2, q = tid & 3;
    const bool act = s_act[r] != 0;
    int mean = 0;
    if (act) {
        const int shift = 14 - celt_ilog2(1 + imax(s_mx[r], -s_mn[r]));
        const int *row = reinterpret_cast<const int *>(rows + r * kTr2Row);
        int16_t *E = Es + r * kTr2E;
        for (int i = q; i < len2; i += 4) {
            const int w = row[i];
            int a = s16(w), b = w >> 16;
            if (shift != 0) {   // SHL16 with a NEGATIVE count (maxabs == 32768): what the reference's C expression does on x86
                a = (int16_t)((unsigned)(uint16_t)a << (shift & 31));
                b = (int16_t)((unsigned)(uint16_t)b << (shift & 31));
            }
            const int x2 = s16(pshr32(wadd(mul16_16(a, a), mul16_16(b, b)), 16));
            E[i] = (int16_t)x2;
            mean = wadd(mean, x2);
        }
    }
    mean = wadd(mean, __shfl_xor_sync(0xffffffffu, mean, 1));
    mean = wadd(mean, __shfl_xor_sync(0xffffffffu, mean, 2));
    __syncthreads();
    // ---- C ----
    if (warp == 0 && s_act[lane]) {
        int *Ew = reinterpret_cast<int *>(Es + lane * kTr2E);
        const int nw = len2 >> 1;
        int mem0 = 0;
#pragma unroll 4
        for (int k = 0; k < nw; k++) {
            const int w = Ew[k];
            mem0 = s16(mem0 + pshr32(s16(w) - mem0, 4));
            const int lo = mem0;
            mem0 = s16(mem0 + pshr32((w >> 16) - mem0, 4));
            Ew[k] = (lo & 0xffff) | (mem0 << 16);
        }
        mem0 = 0;
        int maxE = 0;
#pragma unroll 4
        for (int k = nw - 1; k >= 0; k--) {
            const int w = Ew[k];
            mem0 = s16(mem0 + pshr32((w >> 16) - mem0, 3));
            const int hi = mem0;
            maxE = imax(maxE, mem0);
            mem0 = s16(mem0 + pshr32(s16(w) - mem0, 3));
            maxE = imax(maxE, mem0);
            Ew[k] = (mem0 & 0xffff) | (hi << 16);
        }
        s_maxE[lane] = maxE;
    }
    __syncthreads();
    // ---- D ----
    int unmask = 0;
    if (act) {
        const int m = mul16_16(celt_sqrt(mean), celt_sqrt(mul16_16(s_maxE[r], len2 >> 1)));
        const int norm = shl32(len2, 6 + 14) / wadd(1, m >> 1);
        const int16_t *E = Es + r * kTr2E;
        for (int i = 12 + 4 * q; i < len2 - 5; i += 16) {
            const int id = imax(0, imin(127, mul16_32_q15(E[i] + 1, norm)));
            unmask += kInvTable[id];
        }
    }
    unmask += __shfl_xor_sync(0xffffffffu, unmask, 1);
    unmask += __shfl_xor_sync(0xffffffffu, unmask, 2);
    if (q == 0 && r < nvalid) {
        const int ws = (p0 + r) / g.CC, wc = (p0 + r) - ws * g.CC;
        if (act) ctx[ws].v.mask_metric[wc] = 64 * unmask * 4 / (6 * (len2 - 17));
        if (fl.prog) {
            __threadfence();
            atomicAdd(fl.prog + ws, 1);
        }
    }
}

// ---- K3: one warp per stream ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_PIPE_WPB * 32, 7)
pipe_transform_kernel(CbEncState *pool, const int *slots, PipeGeom g, EncPipeCtx *ctx, EncPipeBuf *buf, Flow fl) {
    __shared__ __align__(16) TransformScratch sm[CB_PIPE_WPB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * CB_PIPE_WPB + wib;
    flow_open(fl);
    if (s >= g.n) return;
    flow_wait_warp(fl, s);
    FreeWarpTeam tm{{lane}};
    pipe_transform(tm, pool + slots[s], g, ctx[s], buf[s], sm[wib]);
    flow_done_warp(fl, s);
}

// ---- K4 ---------------------------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kScalarThreads)
pipe_decide_kernel(CbEncState *pool, const int *slots, PipeGeom g, EncPipeCtx *ctx, Flow fl) {
    extern __shared__ __align__(16) int sm[];
    constexpr int kScalarT = 8 * L;
    __shared__ int *p_head[kScalarT], *p_ctx[kScalarT];
    int *sm_head = sm, *sm_ctx = sm + kScalarT * kHeadStride + (kScalarT * kHeadStride & 1);
    const int t0 = blockIdx.x * kScalarT;
    const int tid = (threadIdx.x & 31) < L ? (int)(threadIdx.x >> 5) * L + (int)(threadIdx.x & 31) : kScalarT;   // the stream this thread works on, if any
    const int nvalid = g.n - t0 < kScalarT ? g.n - t0 : kScalarT;
    flow_open(fl);
    flow_wait_block(fl, t0, nvalid);
    if (tid < nvalid) { p_head[tid] = reinterpret_cast<int *>(pool + slots[t0 + tid]); p_ctx[tid] = reinterpret_cast<int *>(ctx + t0 + tid); }
    __syncthreads();
    stage_copy_in(sm_head, kHeadStride, kHeadWords, p_head, nvalid);
    stage_copy_in(sm_ctx, kCtxStride, kCtxWords, p_ctx, nvalid);
    __syncthreads();
    if (tid < nvalid)
        pipe_decide(reinterpret_cast<CbEncState *>(sm_head + tid * kHeadStride), g, *reinterpret_cast<EncPipeCtx *>(sm_ctx + tid * kCtxStride));
    __syncthreads();
    stage_copy_out_head(sm_head, p_head, nvalid);
    stage_copy_out(sm_ctx, kCtxStride, kCtxWords, p_ctx, nvalid);
    flow_done_block(fl, t0, nvalid);
}

// ---- K5: one warp per stream ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CB_PIPE_WPB * 32)
pipe_bands_kernel(CbEncState *pool, const int *slots, const int *sidx, PipeGeom g, int f, int fi, const EncPlan *plans, EncPipeCtx *ctx,
                  EncPipeBuf *buf, uint8_t *data, int *rets, unsigned *ranges) {
    __shared__ __align__(16) BandScratch sm[CB_PIPE_WPB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * CB_PIPE_WPB + wib;
    if (s >= g.n) return;
    FreeWarpTeam tm{{lane}};
    CbEncState *st = pool + slots[s];
    const size_t k = (size_t)sidx[s] * g.F + f;
    const int r = pipe_bands(tm, st, g, plans[(size_t)s * g.Fc + fi], ctx[s], buf[s], sm[wib], data + k * g.stride);
    if (lane == 0) {
        rets[k] = r;
        if (ranges) ranges[k] = st->rangeFinal;
    }
}

// ---- K5a..K5d: the band loop as prep / chain-S / leaves / chain-X (celt_enc_bandpipe.cuh) ----------------------------------------
__global__ void __launch_bounds__(CB_PIPE_WPB * 32, 7)
pipe_prep_kernel(const CbEncState *pool, const int *slots, PipeGeom g, const EncPipeCtx *ctx, const EncPipeBuf *buf, BandPrep *prep, int16_t *xall,
                 Flow fl) {
    __shared__ __align__(16) PrepScratch sm[CB_PIPE_WPB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * CB_PIPE_WPB + wib;
    flow_open(fl);
    if (s >= g.n) return;
    flow_wait_warp(fl, s);
    FreeWarpTeam tm{{lane}};
    pipe_band_prep(tm, pool + slots[s], g, ctx[s], buf[s], prep[s], xall + (size_t)s * kXallStride, sm[wib]);
    flow_done_warp(fl, s);
}
template <int L>
__global__ void __launch_bounds__(kScalarThreads)
pipe_spec_kernel(const CbEncState *pool, const int *slots, PipeGeom g, EncPipeCtx *ctx, BandPrep *prep, LeafList *leaves) {
    extern __shared__ __align__(16) int sm[];
    constexpr int kScalarT = 8 * L;
    __shared__ int *p_prep[kScalarT], *p_leaf[kScalarT];
    int *sm_prep = sm, *sm_leaf = sm + kScalarT * kPrepStride;
    const int t0 = blockIdx.x * kScalarT;
    const int tid = (threadIdx.x & 31) < L ? (int)(threadIdx.x >> 5) * L + (int)(threadIdx.x & 31) : kScalarT;   // the stream this thread works on, if any
    const int nvalid = g.n - t0 < kScalarT ? g.n - t0 : kScalarT;
    if (tid < nvalid) { p_prep[tid] = reinterpret_cast<int *>(prep + t0 + tid); p_leaf[tid] = reinterpret_cast<int *>(leaves + t0 + tid); }
    __syncthreads();
    stage_copy_in(sm_prep, kPrepStride, kPrepWords, p_prep, nvalid);
    __syncthreads();
    if (tid < nvalid)
        pipe_band_spec(pool + slots[t0 + tid], g, ctx[t0 + tid], *reinterpret_cast<const BandPrep *>(sm_prep + tid * kPrepStride),
                       *reinterpret_cast<LeafList *>(sm_leaf + tid * kLeafStride));
    __syncthreads();
    stage_copy_out(sm_leaf, kLeafStride, kLeafTaskWords, p_leaf, nvalid);
}
__global__ void __launch_bounds__(CB_PIPE_WPB * 32)
pipe_leaves_kernel(const CbEncState *pool, const int *slots, PipeGeom g, const EncPipeCtx *ctx, LeafList *leaves, const int16_t *xall) {
    __shared__ __align__(16) LeafScratch sm[CB_PIPE_WPB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * CB_PIPE_WPB + wib;
    if (s >= g.n) return;
    FreeWarpTeam tm{{lane}};
    pipe_leaves(tm, pool + slots[s], ctx[s], leaves[s], xall + (size_t)s * kXallStride, sm[wib]);
}
template <int L>
__global__ void __launch_bounds__(kScalarThreads)
pipe_exact_kernel(CbEncState *pool, const int *slots, const int *sidx, PipeGeom g, int f, int fi, const EncPlan *plans, EncPipeCtx *ctx,
                  BandPrep *prep, LeafList *leaves, int16_t *xall, uint8_t *data, int *rets, unsigned *ranges, int *misses) {
    extern __shared__ __align__(16) int sm[];
    constexpr int kScalarT = 8 * L;
    __shared__ int *p_head[kScalarT], *p_prep[kScalarT], *p_leaf[kScalarT];
    int *sm_head = sm, *sm_prep = sm_head + kScalarT * kHeadStride, *sm_leaf = sm_prep + kScalarT * kPrepStride;
    const int t0 = blockIdx.x * kScalarT;
    const int tid = (threadIdx.x & 31) < L ? (int)(threadIdx.x >> 5) * L + (int)(threadIdx.x & 31) : kScalarT;   // the stream this thread works on, if any
    const int nvalid = g.n - t0 < kScalarT ? g.n - t0 : kScalarT;
    if (tid < nvalid) {
        p_head[tid] = reinterpret_cast<int *>(pool + slots[t0 + tid]);
        p_prep[tid] = reinterpret_cast<int *>(prep + t0 + tid);
        p_leaf[tid] = reinterpret_cast<int *>(leaves + t0 + tid);
    }
    __syncthreads();
    stage_copy_in(sm_head, kHeadStride, kHeadWords, p_head, nvalid);
    stage_copy_in(sm_prep, kPrepStride, kPrepWords, p_prep, nvalid);
    stage_copy_in(sm_leaf, kLeafStride, kLeafWords, p_leaf, nvalid);
    __syncthreads();
    if (tid < nvalid) {
        const int t = t0 + tid;
        CbEncState *st = reinterpret_cast<CbEncState *>(sm_head + tid * kHeadStride);
        const LeafList &L = *reinterpret_cast<const LeafList *>(sm_leaf + tid * kLeafStride);
        const size_t k = (size_t)sidx[t] * g.F + f;
        int miss = 0;
        const int r = pipe_band_exact_finish(st, g, plans[(size_t)t * g.Fc + fi], ctx[t], *reinterpret_cast<const BandPrep *>(sm_prep + tid * kPrepStride), L,
                                             xall + (size_t)t * kXallStride, data + k * g.stride, &miss);
        rets[k] = r;
        if (ranges) ranges[k] = st->rangeFinal;
        if (miss) atomicAdd(misses, miss);
        atomicAdd(misses + 1, L.count);
    }
    __syncthreads();
    stage_copy_out_head(sm_head, p_head, nvalid);
}

// ---- K5 as prep + one inline walk: one warp per stream, all of a launch's streams resident at once (<= 72 registers) ----------------
// SYNC: the warps of a block meet at every band (SyncWarpTeam), so a block's warps run the same region of the walk's code.
template <int WPB, bool SYNC>
__global__ void __launch_bounds__(WPB * 32, 28 / WPB)
pipe_walk_kernel(CbEncState *pool, const int *slots, const int *sidx, PipeGeom g, int f, int fi, const EncPlan *plans, EncPipeCtx *ctx,
                 const BandPrep *prep, const int16_t *xall, uint8_t *data, int *rets, unsigned *ranges, Flow fl, unsigned sync_mask) {
    extern __shared__ __align__(16) unsigned char smw[];
    WalkScratch *sm = reinterpret_cast<WalkScratch *>(smw);
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * WPB + wib;
    flow_open(fl);
    if (s >= g.n) {
        if (SYNC)
            for (int i = 0; i < kNbEBands; i++)
                if ((sync_mask >> i) & 1u) __syncthreads();
        return;
    }
    flow_wait_warp(fl, s);
    // prep's results are produced while this kernel is already running (dataflow launches): they must not be read through the
    // non-coherent path the compiler picks for memory a kernel provably never writes — so the kernel "may" write them.
    if (g.n < 0) {
        const_cast<BandPrep *>(prep)[s].hasB[0] = 0;
        const_cast<int16_t *>(xall)[0] = 0;
    }
    CbEncState *st = pool + slots[s];
    const size_t k = (size_t)sidx[s] * g.F + f;
    int r;
    if (SYNC) {
        SyncWarpTeam tm{{lane}};
        r = pipe_band_inline_finish(tm, st, g, plans[(size_t)s * g.Fc + fi], ctx[s], prep[s], xall + (size_t)s * kXallStride, sm[wib], data + k * g.stride,
                                    sync_mask);
    } else {
        FreeWarpTeam tm{{lane}};
        r = pipe_band_inline_finish(tm, st, g, plans[(size_t)s * g.Fc + fi], ctx[s], prep[s], xall + (size_t)s * kXallStride, sm[wib], data + k * g.stride);
    }
    if (lane == 0) {
        if (fl.prog && *(volatile int *)fl.err) r = -3;   // OPUS_INTERNAL_ERROR: a dataflow wait timed out somewhere in this launch
        rets[k] = r;
        if (ranges) ranges[k] = st->rangeFinal;
    }
    flow_done_warp(fl, s);
}

#if defined(CB_ROT_LUT)
__global__ void rot_lut_kernel() {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3 * kRotLen * kRotK) {
        const int K = t % kRotK, len = (t / kRotK) % kRotLen, spread = t / (kRotK * kRotLen) + 1;
        uint32_t w = 0;
        if (len >= 2 && 2 * K < len) {
            int c, s;
            rotation_params(len, K, spread, c, s);
            w = ((uint32_t)c & 0xffffu) | ((uint32_t)s << 16);
        }
        g_rot_cs[t] = w;
    }
    if (t < kRotLen * 8) g_rot_s2[t] = (uint8_t)rotation_stride2(t >> 3, 1 << (t & 7));
    if (t < 32) g_inv16[t] = t < 2 ? 65536u : (65536u + (uint32_t)t - 1u) / (uint32_t)t;
}
#endif

__global__ void b2p_lut_kernel() {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (kMaxLM + 2) * kNbEBands * kB2pBits) return;
    const int bits = t % kB2pBits, row = t / kB2pBits;
    const int band = row % kNbEBands, LM = row / kNbEBands - 1;
    g_b2p_lut[t] = (uint8_t)bits2pulses(band, LM, bits);
}

// ---- end of the span: the last 1024 pre-emphasised samples go back into the state -------------------------------------------------
__global__ void pipe_epilogue_kernel(CbEncState *pool, const int *slots, PipeGeom g, const int *P, int last_nfr) {
    const int s = blockIdx.x;
    if (s >= g.n) return;
    CbEncState *st = pool + slots[s];
    for (int c = 0; c < g.CC; c++) {
        const int *src = P + ((size_t)s * g.CC + c) * g.pstride + last_nfr * g.N;
        for (int i = threadIdx.x; i < kCombMaxPeriod; i += blockDim.x) st->prefilter_mem[c * kCombMaxPeriod + i] = src[i];
    }
}

constexpr int kSmemTransient = (32 * kTrTileRow + 32 * kTrRowWords) * 4;
constexpr int smem_head_ctx(int T) { return (T * kHeadStride + 1 + T * kCtxStride) * 4; }
constexpr int smem_spec(int T) { return (T * kPrepStride + T * kLeafStride) * 4; }
constexpr int smem_exact(int T) { return (T * kHeadStride + T * kPrepStride + T * kLeafStride) * 4; }

typedef CbDevBuf DevBuf;

enum { kMaxGroups = 8 };
struct Group {
    cudaStream_t main = nullptr, side = nullptr;
    cudaEvent_t ev_fe[2] = {nullptr, nullptr}, ev_steps[2] = {nullptr, nullptr}, ev_done = nullptr;
    DevBuf D, P[2], plans[2], fe[2], m0, ctx, buf, prep, leaves, xall, prog;
};
struct PipeCtx {
    bool init = false;
    int groups = 1;             // stream groups on separate CUDA streams (measured: no gain, the stages are issue bound with all streams resident)
    int first_chunk = 2;        // frames in the first chunk of a span (its front end is exposed); 0: same as the others
    int chunk = 8;              // frames per front-end chunk (measured 4..50: 8 is best, profiles/r2_encoder_ab.md)
    int walk_mode = 3;          // band-walk kernel: 0 = 4 free-running warps per block; 1..4 = 4 / 7 / 14 / 28 warps meeting at every band (measured: 3)
    unsigned walk_sync_mask = 0x1fffffu;   // bands at whose start the walk's block meets (CB200_ENC_WALK_SYNC, hex)
    int scalar_l = 2;           // streams per warp in the thread-per-stream stages
    int flow = 0;               // per-stream dataflow between the kernels of a frame step (struct Flow); 0: stream-ordered launches.
                                // Measured (profiles/r2_encoder_ab.md): 8 % SLOWER than stream order — spinning blocks hold the slots ready
                                // blocks need — and the band walk faults when it overlaps its predecessor; kept as an A/B knob only.
    int prepass_v = 2;                  // Opus-layer prepass: 2 = thread per stream with vector loads, 1 = the tiled first version (A/B)
    int probe_fe2 = 1;                  // dev probe: run the pitch stage this many times
    int flow_noopen = 0;                // debugging: kernels of a frame step that do not let their successor start early
    int flow_mask = -1, flow_seq = 0;   // debugging: which of a frame step's seven kernels get the early launch
    int *d_flow_err = nullptr;  // raised by a timed-out dataflow wait
    int transient_v = 2;        // transient analysis kernel: 2 = short chains (pipe_transient2_kernel), 1 = the tiled first version (A/B)
    int split_bands = 2;        // the band loop as 2: prep + one inline walk; 1: prep / chain-S / leaves / chain-X; 0: one stage (A/B)
    Group g[kMaxGroups];
    cudaEvent_t ev_fork = nullptr;
    int *d_stats = nullptr;     // [0] leaves chain-X had to search itself, [1] leaves listed by chain-S
};
PipeCtx pcs[kCbMaxDevices];     // one per device (host_runtime.h)
#define pc (pcs[opus_b200_current_device()])

template <int L>
void set_scalar_smem() {
    cudaFuncSetAttribute(pipe_head_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_head_ctx(8 * L));
    cudaFuncSetAttribute(pipe_decide_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_head_ctx(8 * L));
    cudaFuncSetAttribute(pipe_spec_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_spec(8 * L));
    cudaFuncSetAttribute(pipe_exact_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_exact(8 * L));
}

bool pipe_init() {
    if (pc.init) return true;
    if (const char *e = getenv("CB200_ENC_GROUPS")) pc.groups = atoi(e);
    if (pc.groups < 1) pc.groups = 1;
    if (pc.groups > kMaxGroups) pc.groups = kMaxGroups;
    if (const char *e = getenv("CB200_ENC_CHUNK")) pc.chunk = atoi(e);
    if (pc.chunk < 1) pc.chunk = 1;
    if (const char *e = getenv("CB200_ENC_FIRST_CHUNK")) pc.first_chunk = atoi(e);
    if (const char *e = getenv("CB200_ENC_SPLIT_BANDS")) pc.split_bands = atoi(e);
    if (cudaMalloc(&pc.d_stats, 2 * sizeof(int)) != cudaSuccess) return false;
    cudaMemset(pc.d_stats, 0, 2 * sizeof(int));
    if (cudaMalloc(&pc.d_flow_err, sizeof(int)) != cudaSuccess) return false;
    cudaMemset(pc.d_flow_err, 0, sizeof(int));
    if (const char *e = getenv("CB200_ENC_FLOW")) pc.flow = atoi(e);
    if (const char *e = getenv("CB200_ENC_FLOW_MASK")) pc.flow_mask = atoi(e);
    if (const char *e = getenv("CB200_ENC_FLOW_NOOPEN")) pc.flow_noopen = atoi(e);
    if (const char *e = getenv("CB200_ENC_PROBE_FE2")) pc.probe_fe2 = atoi(e);
    if (const char *e = getenv("CB200_ENC_PREPASS")) pc.prepass_v = atoi(e);
    if (const char *e = getenv("CB200_ENC_STACK")) cudaDeviceSetLimit(cudaLimitStackSize, (size_t)atoi(e));
    int prio_lo = 0, prio_hi = 0, prio = 1;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (const char *e = getenv("CB200_ENC_PRIO")) prio = atoi(e);
    for (int i = 0; i < kMaxGroups; i++) {
        Group &G = pc.g[i];
        // The frame steps are one-wave kernels that want every SM: their blocks go first; the front end of the next chunk (many
        // short blocks) fills what is left instead of pushing a frame step's last blocks into a second wave.
        if (cudaStreamCreateWithPriority(&G.main, cudaStreamNonBlocking, prio ? prio_hi : 0) != cudaSuccess) return false;
        if (cudaStreamCreateWithPriority(&G.side, cudaStreamNonBlocking, prio ? prio_lo : 0) != cudaSuccess) return false;
        for (int b = 0; b < 2; b++) {
            cudaEventCreateWithFlags(&G.ev_fe[b], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&G.ev_steps[b], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&G.ev_done, cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&pc.ev_fork, cudaEventDisableTiming);
    if (const char *e = getenv("CB200_ENC_WALK")) pc.walk_mode = atoi(e);
    if (const char *e = getenv("CB200_ENC_WALK_SYNC")) pc.walk_sync_mask = (unsigned)strtoul(e, nullptr, 16);
    cudaFuncSetAttribute(pipe_walk_kernel<14, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(14 * sizeof(WalkScratch)));
    cudaFuncSetAttribute(pipe_walk_kernel<28, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(28 * sizeof(WalkScratch)));
    b2p_lut_kernel<<<((kMaxLM + 2) * kNbEBands * kB2pBits + 255) / 256, 256>>>();
#if defined(CB_ROT_LUT)
    rot_lut_kernel<<<(3 * kRotLen * kRotK + 255) / 256, 256>>>();
#endif
    cudaDeviceSynchronize();
    cudaFuncSetAttribute(pipe_transient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTransient);
    cudaFuncSetAttribute(pipe_transient2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTransient2);
    if (const char *e = getenv("CB200_ENC_TRANSIENT")) pc.transient_v = atoi(e);
    if (const char *e = getenv("CB200_ENC_SCALAR_L")) pc.scalar_l = atoi(e);
    if (pc.scalar_l != 1 && pc.scalar_l != 2 && pc.scalar_l != 4) pc.scalar_l = 2;
    set_scalar_smem<1>(); set_scalar_smem<2>(); set_scalar_smem<4>();
    pc.init = cudaGetLastError() == cudaSuccess;
    return pc.init;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// A launch that may begin before the previous kernel of the stream has finished (programmatic stream serialisation) when `early`.
template <typename... KArgs, typename... Args>
void launch_flow(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool early, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = early && pc.flow == 1 ? 1 : 0;   // CB200_ENC_FLOW=2: the counters without the early launches (debugging)
    if (early && pc.flow == 1 && pc.flow_mask >= 0) cfg.numAttrs = (pc.flow_mask >> (pc.flow_seq % 7)) & 1;
    pc.flow_seq++;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// launch a thread-per-stream stage with the configured number of streams per warp
#define CB_SCALAR_LAUNCH(KERNEL, SMEMFN, STREAM, ...)                                                                         \
    switch (pc.scalar_l) {                                                                                                    \
    case 4: KERNEL<4><<<cdiv(n, 32), kScalarThreads, SMEMFN(32), STREAM>>>(__VA_ARGS__); break;                               \
    case 2: KERNEL<2><<<cdiv(n, 16), kScalarThreads, SMEMFN(16), STREAM>>>(__VA_ARGS__); break;                               \
    default: KERNEL<1><<<cdiv(n, 8), kScalarThreads, SMEMFN(8), STREAM>>>(__VA_ARGS__); break;                                \
    }
#define CB_SCALAR_LAUNCH_FLOW(KERNEL, SMEMFN, STREAM, EARLY, ...)                                                             \
    switch (pc.scalar_l) {                                                                                                    \
    case 4: launch_flow(KERNEL<4>, cdiv(n, 32), kScalarThreads, SMEMFN(32), STREAM, EARLY, __VA_ARGS__); break;               \
    case 2: launch_flow(KERNEL<2>, cdiv(n, 16), kScalarThreads, SMEMFN(16), STREAM, EARLY, __VA_ARGS__); break;               \
    default: launch_flow(KERNEL<1>, cdiv(n, 8), kScalarThreads, SMEMFN(8), STREAM, EARLY, __VA_ARGS__); break;                \
    }

// one group: streams [k0, k0+n) of the pipeline's list
int enqueue_group(Group &G, const EncPipeCall &c, int k0, int n, PipeGeom g) {
    g.n = n;
    const int nframes = c.f1 - c.f0;
    const int Fc = g.Fc;
    const size_t row = (size_t)g.fsz * g.CC;
    if (!G.D.reserve((size_t)n * Fc * row * sizeof(int16_t)) || !G.m0.reserve((size_t)n * 2 * sizeof(int)) ||
        !G.ctx.reserve((size_t)n * sizeof(EncPipeCtx)) || !G.buf.reserve((size_t)n * sizeof(EncPipeBuf)) ||
        !G.prep.reserve((size_t)n * sizeof(BandPrep)) || !G.leaves.reserve((size_t)n * sizeof(LeafList)) ||
        !G.xall.reserve((size_t)n * kXallStride * sizeof(int16_t)))
        return -7;
    for (int b = 0; b < 2; b++)
        if (!G.P[b].reserve((size_t)n * g.CC * g.pstride * sizeof(int)) || !G.plans[b].reserve((size_t)n * Fc * sizeof(EncPlan)) ||
            !G.fe[b].reserve((size_t)n * Fc * sizeof(FeFrame)))
            return -7;
    const int *slots = c.d_slots + k0, *sidx = c.d_sidx + k0;
    int launches = 0;
    const bool flow = pc.flow && pc.split_bands == 2 && pc.walk_mode == 0 && pc.transient_v == 2;

    if (flow) {
        if (!G.prog.reserve((size_t)n * sizeof(int))) return -7;
        cudaMemsetAsync(G.prog.p, 0, (size_t)n * sizeof(int), G.main);
    }
    // The front end of the FIRST chunk overlaps nothing (the frame steps wait for it), and its prepass is a chain over the chunk's
    // frames: the first chunk is short (CB200_ENC_FIRST_CHUNK frames), the rest have Fc.
    int nchunks = 0, done = 0;
    int prev_nfr = 0;
    for (int k = 0; done < nframes; k++) {
        const int b = k & 1;
        const int fbase = c.f0 + done;
        const int want = k == 0 && pc.first_chunk > 0 && pc.first_chunk < Fc ? pc.first_chunk : Fc;
        const int nfr = nframes - done < want ? nframes - done : want;
        // ---- front end of chunk k on the side stream (buffers b are free once the frame steps of chunk k-2 are done) ----
        if (k >= 2) cudaStreamWaitEvent(G.side, G.ev_steps[b], 0);
        if (pc.prepass_v == 2 && (row * sizeof(int16_t)) % 16 == 0 && (g.CC == 2 || g.CC == 1))
            pipe_prepass2_kernel<<<cdiv(n, 32), 32, 0, G.side>>>(c.pool, slots, sidx, g, c.d_pcm, fbase, nfr, (int16_t *)G.D.p, (EncPlan *)G.plans[b].p,
                                                                (int *)G.m0.p);
        else
            pipe_prepass_kernel<<<cdiv(n, 32), 128, 0, G.side>>>(c.pool, slots, sidx, g, c.d_pcm, fbase, nfr, (int16_t *)G.D.p, (EncPlan *)G.plans[b].p,
                                                                  (int *)G.m0.p);
        pipe_fe1_kernel<<<cdiv(n * nfr, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.side>>>(c.pool, slots, g, nfr, (const int16_t *)G.D.p,
                                                                                    (const EncPlan *)G.plans[b].p, (const int *)G.m0.p, (int *)G.P[b].p,
                                                                                    k > 0 ? (const int *)G.P[b ^ 1].p : nullptr, prev_nfr,
                                                                                    (FeFrame *)G.fe[b].p);
        pipe_fe2_kernel<<<cdiv(n * nfr, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.side>>>(g, nfr, (const EncPlan *)G.plans[b].p, (const int *)G.P[b].p,
                                                                                    (FeFrame *)G.fe[b].p);
        if (pc.probe_fe2 > 1)   // dev probe: how much of the front end's cost reaches the span's time (fe2 is idempotent)
            for (int r = 1; r < pc.probe_fe2; r++)
                pipe_fe2_kernel<<<cdiv(n * nfr, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.side>>>(g, nfr, (const EncPlan *)G.plans[b].p, (const int *)G.P[b].p,
                                                                                            (FeFrame *)G.fe[b].p);
        cudaEventRecord(G.ev_fe[b], G.side);
        launches += 3;
        // ---- frame steps of chunk k on the main stream ----
        cudaStreamWaitEvent(G.main, G.ev_fe[b], 0);
        for (int fi = 0; fi < nfr; fi++) {
            const int f = fbase + fi;
            if (flow) {
                // units a stream completes per frame: head 1, comb CC, transient CC, transform 1, decide 1, prep 1, walk 1
                const int U = 5 + 2 * g.CC, base = (done + fi) * U;
                int *prog = (int *)G.prog.p;
                int kidx = 0;
                auto FL = [&](int need) {
                    Flow fl;
                    fl.prog = prog; fl.need = base + need; fl.err = pc.d_flow_err;
                    fl.open = pc.flow == 1 && !((pc.flow_noopen >> kidx++) & 1);
                    return fl;
                };
                const bool chained = fi > 0;   // the chunk's first kernel waits for the front end's event: launched in stream order
                CB_SCALAR_LAUNCH_FLOW(pipe_head_kernel, smem_head_ctx, G.main, chained, c.pool, slots, sidx, g, f, fi, (const EncPlan *)G.plans[b].p,
                                      (const FeFrame *)G.fe[b].p, (EncPipeCtx *)G.ctx.p, c.d_data, FL(0))
                launch_flow(pipe_comb_kernel, cdiv(n * g.CC, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main, true, c.pool, slots, g, fi, (const int *)G.P[b].p,
                            (EncPipeCtx *)G.ctx.p, (EncPipeBuf *)G.buf.p, FL(1));
                launch_flow(pipe_transient2_kernel, cdiv(n * g.CC, 32), 128, (size_t)kSmemTransient2, G.main, true, g, (EncPipeCtx *)G.ctx.p,
                            (const EncPipeBuf *)G.buf.p, FL(1 + g.CC));
                launch_flow(pipe_transform_kernel, cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main, true, c.pool, slots, g, (EncPipeCtx *)G.ctx.p,
                            (EncPipeBuf *)G.buf.p, FL(1 + 2 * g.CC));
                CB_SCALAR_LAUNCH_FLOW(pipe_decide_kernel, smem_head_ctx, G.main, true, c.pool, slots, g, (EncPipeCtx *)G.ctx.p, FL(2 + 2 * g.CC))
                launch_flow(pipe_prep_kernel, cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main, true, (const CbEncState *)c.pool, slots, g,
                            (const EncPipeCtx *)G.ctx.p, (const EncPipeBuf *)G.buf.p, (BandPrep *)G.prep.p, (int16_t *)G.xall.p, FL(3 + 2 * g.CC));
                launch_flow(pipe_walk_kernel<4, false>, cdiv(n, 4), 4 * 32, 4 * sizeof(WalkScratch), G.main, true, c.pool, slots, sidx, g, f, fi,
                            (const EncPlan *)G.plans[b].p, (EncPipeCtx *)G.ctx.p, (const BandPrep *)G.prep.p, (const int16_t *)G.xall.p, c.d_data,
                            c.d_rets, c.d_ranges, FL(4 + 2 * g.CC), 0u);
                launches += 7;
                continue;
            }
            const Flow nofl{nullptr, 0, nullptr, 0};
            CB_SCALAR_LAUNCH(pipe_head_kernel, smem_head_ctx, G.main, c.pool, slots, sidx, g, f, fi, (const EncPlan *)G.plans[b].p,
                             (const FeFrame *)G.fe[b].p, (EncPipeCtx *)G.ctx.p, c.d_data, nofl)
            pipe_comb_kernel<<<cdiv(n * g.CC, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main>>>(c.pool, slots, g, fi, (const int *)G.P[b].p, (EncPipeCtx *)G.ctx.p,
                                                                                          (EncPipeBuf *)G.buf.p, nofl);
            if (pc.transient_v == 2)
                pipe_transient2_kernel<<<cdiv(n * g.CC, 32), 128, kSmemTransient2, G.main>>>(g, (EncPipeCtx *)G.ctx.p, (const EncPipeBuf *)G.buf.p, nofl);
            else
                pipe_transient_kernel<<<cdiv(n * g.CC, 32), 128, kSmemTransient, G.main>>>(g, (EncPipeCtx *)G.ctx.p, (const EncPipeBuf *)G.buf.p);
            pipe_transform_kernel<<<cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main>>>(c.pool, slots, g, (EncPipeCtx *)G.ctx.p, (EncPipeBuf *)G.buf.p, nofl);
            CB_SCALAR_LAUNCH(pipe_decide_kernel, smem_head_ctx, G.main, c.pool, slots, g, (EncPipeCtx *)G.ctx.p, nofl)
            if (!pc.split_bands) {
                pipe_bands_kernel<<<cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main>>>(c.pool, slots, sidx, g, f, fi, (const EncPlan *)G.plans[b].p,
                                                                                        (EncPipeCtx *)G.ctx.p, (EncPipeBuf *)G.buf.p, c.d_data, c.d_rets,
                                                                                        c.d_ranges);
                launches += 6;
            } else if (pc.split_bands == 2) {
                pipe_prep_kernel<<<cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main>>>(c.pool, slots, g, (const EncPipeCtx *)G.ctx.p,
                                                                                       (const EncPipeBuf *)G.buf.p, (BandPrep *)G.prep.p, (int16_t *)G.xall.p, nofl);
#define CB_WALK_LAUNCH(WPB, SYNC)                                                                                                         \
    pipe_walk_kernel<WPB, SYNC><<<cdiv(n, WPB), WPB * 32, WPB * sizeof(WalkScratch), G.main>>>(                                            \
        c.pool, slots, sidx, g, f, fi, (const EncPlan *)G.plans[b].p, (EncPipeCtx *)G.ctx.p, (const BandPrep *)G.prep.p, (const int16_t *)G.xall.p, \
        c.d_data, c.d_rets, c.d_ranges, nofl, pc.walk_sync_mask)
                switch (pc.walk_mode) {
                case 1: CB_WALK_LAUNCH(4, true); break;
                case 2: CB_WALK_LAUNCH(7, true); break;
                case 3: CB_WALK_LAUNCH(14, true); break;
                case 4: CB_WALK_LAUNCH(28, true); break;
                default: CB_WALK_LAUNCH(4, false); break;
                }
                launches += 7;
            } else {
                pipe_prep_kernel<<<cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main>>>(c.pool, slots, g, (const EncPipeCtx *)G.ctx.p,
                                                                                       (const EncPipeBuf *)G.buf.p, (BandPrep *)G.prep.p, (int16_t *)G.xall.p, nofl);
                CB_SCALAR_LAUNCH(pipe_spec_kernel, smem_spec, G.main, c.pool, slots, g, (EncPipeCtx *)G.ctx.p, (BandPrep *)G.prep.p, (LeafList *)G.leaves.p)
                pipe_leaves_kernel<<<cdiv(n, CB_PIPE_WPB), CB_PIPE_WPB * 32, 0, G.main>>>(c.pool, slots, g, (const EncPipeCtx *)G.ctx.p, (LeafList *)G.leaves.p,
                                                                                         (const int16_t *)G.xall.p);
                CB_SCALAR_LAUNCH(pipe_exact_kernel, smem_exact, G.main, c.pool, slots, sidx, g, f, fi, (const EncPlan *)G.plans[b].p, (EncPipeCtx *)G.ctx.p,
                                 (BandPrep *)G.prep.p, (LeafList *)G.leaves.p, (int16_t *)G.xall.p, c.d_data, c.d_rets, c.d_ranges, pc.d_stats)
                launches += 9;
            }
        }
        cudaEventRecord(G.ev_steps[b], G.main);
        prev_nfr = nfr;
        done += nfr;
        nchunks = k + 1;
    }
    pipe_epilogue_kernel<<<n, 128, 0, G.main>>>(c.pool, slots, g, (const int *)G.P[(nchunks - 1) & 1].p, prev_nfr);
    launches++;
    return launches;
}

}  // namespace

// leaves the exact chain had to search itself / leaves listed by the speculative chain, since start (reads after a device sync)
void enc_pipe_stats(long long *misses, long long *leaves) {
    int h[2] = {0, 0};
    if (pc.init) cudaMemcpy(h, pc.d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    if (misses) *misses = h[0];
    if (leaves) *leaves = h[1];
}

int enc_pipe_takes(const CbEncState *st, int frame_size, int out_data_bytes) { return enc_pipe_eligible(st, frame_size, out_data_bytes); }

int enc_pipe_enqueue(const EncPipeCall &c, cudaStream_t stream) {
    if (!pipe_init()) return -3;
    if (c.n <= 0 || c.f1 <= c.f0) return 0;
    PipeGeom g;
    g.n = c.n; g.CC = c.channels; g.Fs = c.Fs; g.upsample = 48000 / c.Fs; g.fsz = c.frame_size; g.N = c.frame_size * g.upsample;
    for (g.LM = 0; g.LM <= kMaxLM; g.LM++)
        if ((kShortMdct << g.LM) == g.N) break;
    if (g.LM > kMaxLM) return -1;
    g.F = c.F;
    const int nframes = c.f1 - c.f0;
    g.Fc = nframes < pc.chunk ? nframes : pc.chunk;
    g.max_bytes = c.max_bytes; g.stride = c.stride;
    g.pstride = kPipeHist + g.Fc * g.N;
    // groups: enough streams each to fill the warp-per-stream kernels
    int G = pc.groups;
    while (G > 1 && c.n / G < 256) G--;
    cudaEventRecord(pc.ev_fork, stream);
    int launches = 0;
    const int per = (c.n + G - 1) / G;
    for (int i = 0; i < G; i++) {
        const int k0 = i * per, n = c.n - k0 < per ? c.n - k0 : per;
        if (n <= 0) break;
        Group &Gr = pc.g[i];
        cudaStreamWaitEvent(Gr.main, pc.ev_fork, 0);
        cudaStreamWaitEvent(Gr.side, pc.ev_fork, 0);
        const int r = enqueue_group(Gr, c, k0, n, g);
        if (r < 0) return r;
        launches += r;
        cudaEventRecord(Gr.ev_done, Gr.main);
        cudaStreamWaitEvent(stream, Gr.ev_done, 0);
    }
    {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            fprintf(stderr, "concentus_b200: encoder pipeline launch failed: %s\n", cudaGetErrorString(e));
            return -3;
        }
    }
    return launches;
}
