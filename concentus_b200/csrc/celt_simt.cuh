// celt_simt.cuh — execution-model shim shared by every codec header.
//
// The decoder is a two-stage pipeline:
//   stage A "parse"  — one THREAD per run of packets: range decoder, allocation, band loop (PVQ).  Purely
//                      scalar code; frame-level parallelism comes from the grid.
//   stage B "synth"  — one TEAM per stream (a warp on the GPU): energy prediction, anti-collapse,
//                      IMDCT, post-filter, de-emphasis.  Data-parallel loops are strided over the team.
// Team-parallel code is templated on a team type: WarpTeam (32 lanes, shuffles, __syncwarp) on the
// device, SoloTeam (1 lane, no-ops) for the scalar stage and for the host simulation.  The same headers
// compile under g++ (tests/hostsim, a TEST TOOL that is never linked into the product library).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_DEV __device__ __forceinline__
#define CB_MEM __device__ __forceinline__
#define CB_MEM_NOINLINE __device__ __noinline__
#define CB_DEV_NOINLINE static __device__ __noinline__
#define CB_TABLE static __device__ const
#define CB_CLZ(x) __clz((int)(x))
// Medium-sized helpers (polynomial math, allocation look-ups): inlined where a kernel is dominated by them (decoder stages),
// real calls where instruction-cache footprint matters more (CB_SMALL_CODE: the encoder kernel, profiles/r1_enc_*).
#if defined(CB_SMALL_CODE)
#define CB_MATH static __device__ __noinline__
#define CB_NOUNROLL _Pragma("unroll 1")
#else
#define CB_MATH __device__ __forceinline__
#define CB_NOUNROLL
#endif
// CB_TINY_CODE (the encoder pipeline's kernels, on top of CB_SMALL_CODE): the band-walk kernel's hot set was 46 KB of SASS against a
// 32 KB instruction cache per SM (33 % of its warp samples "no instruction"), so the range coder's renormalisation, ec_tell_frac
// and the rotation chain become real calls shared by their callers.
#if defined(CB_TINY_CODE)
#define CB_MEM_TINY __device__ __noinline__
#define CB_DEV_TINY static __device__ __noinline__
#else
#define CB_MEM_TINY __device__ __forceinline__
#define CB_DEV_TINY __device__ __forceinline__
#endif
#else
#define CB_DEV static inline
#define CB_MEM inline
#define CB_MEM_NOINLINE inline
#define CB_DEV_NOINLINE static
#define CB_TABLE static const
#define CB_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#define CB_MATH static inline
#define CB_NOUNROLL
#define CB_MEM_TINY inline
#define CB_DEV_TINY static inline
struct int4 { int x, y, z, w; };   // host simulation stand-ins for the CUDA vector types
struct int2 { int x, y; };
#endif

namespace cb {

struct SoloTeam {
    static constexpr int W = 1;
    CB_MEM int lane() const { return 0; }
    CB_MEM void sync() const {}
    CB_MEM int sum(int v) const { return v; }
    CB_MEM int max(int v) const { return v; }
    CB_MEM unsigned bor(unsigned v) const { return v; }
    CB_MEM int bcast(int v, int) const { return v; }
    CB_MEM int exscan(int) const { return 0; }   // exclusive prefix sum over lanes (wrapping)
    CB_MEM int shfl_xor(int v, int) const { return v; }
    CB_MEM void phase() const {}
};

#if defined(__CUDACC__)
#if defined(CB_PHASE_PROF)
static __device__ long long g_phase_cycles[64], g_phase_wait[64], g_phase_last;
static __device__ int g_phase_idx;
#endif
struct WarpTeam {
    static constexpr int W = 32;
    int lane_;
    CB_MEM int lane() const { return lane_; }
    CB_MEM void sync() const { __syncwarp(); }
    // warp-wide reductions: one redux.sync each (wrapping 32-bit add, signed max, or) instead of a five-step shuffle tree
#if !defined(CB_NO_REDUX)
    CB_MEM int sum(int v) const { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
    CB_MEM int max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    CB_MEM unsigned bor(unsigned v) const { return __reduce_or_sync(0xffffffffu, v); }
#else
    CB_MEM int sum(int v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    CB_MEM int max(int v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = ::max(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    CB_MEM unsigned bor(unsigned v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
#endif
    CB_MEM int bcast(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    CB_MEM int shfl_xor(int v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    // Phase boundary of a frame: with CB_PHASE_SYNC the warps of a block (each on its own stream) wait for each other here,
    // so that co-resident warps execute the same region of a kernel much larger than the instruction cache.  EVERY warp of the
    // block must pass the same number of phase() calls per frame (see kEncPhases).
#if defined(CB_PHASE_SYNC) && defined(CB_PHASE_GROUP)
    // sub-block groups of CB_PHASE_GROUP warps on named barriers 1..15 (A/B knob: less waiting, more code positions in flight)
    CB_MEM void phase() const {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> 5) / CB_PHASE_GROUP), "r"(CB_PHASE_GROUP * 32) : "memory");
    }
#elif defined(CB_PHASE_SYNC) && defined(CB_PHASE_PROF)
    // dev build: thread 0 of block 0 accumulates the cycles between consecutive phase boundaries per boundary ordinal
    // (tools/enc_phase_prof.py); arrival time is taken before the barrier, so a slot is this stream's own latency
    CB_MEM void phase() const {
        const bool rec = blockIdx.x == 0 && threadIdx.x == 0;
        const long long now = clock64();
        const int k = g_phase_idx % 30;   // kEncPhases
        if (rec) g_phase_cycles[k] += now - g_phase_last;
        __syncthreads();
        if (rec) {
            g_phase_last = clock64();
            g_phase_wait[k] += g_phase_last - now;
            g_phase_idx++;
        }
    }
#elif defined(CB_PHASE_SYNC)
    CB_MEM void phase() const { __syncthreads(); }
#else
    CB_MEM void phase() const {}
#endif
    CB_MEM int exscan(int v) const {
        unsigned inc = (unsigned)v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane_ >= o) inc += t;
        }
        return (int)(inc - (unsigned)v);
    }
};
// A warp team whose phase() is a no-op: kernels of the frame-synchronous encoder pipeline (celt_enc_pipe.cuh) are small enough
// for the instruction cache, and their warps must not meet at block barriers.
struct FreeWarpTeam : WarpTeam {
    CB_MEM void phase() const {}
};
// A warp team whose phase() is a block barrier whatever the build flags say (the pipeline's band-walk kernel: the warps of a block
// step through the bands together so that they share the instruction cache lines of the band they are in).
struct SyncWarpTeam : WarpTeam {
    CB_MEM void phase() const { __syncthreads(); }
};
#endif

}  // namespace cb

// for (i over [0,n)) distributed over the team
#define CB_TEAM_FOR(i, n, tm) CB_NOUNROLL for (int i = (tm).lane(); i < (n); i += (tm).W)
