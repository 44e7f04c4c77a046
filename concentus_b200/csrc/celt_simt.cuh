// celt_simt.cuh — execution-model shim shared by every codec header.
//
// The codec is written once, in "team" style: a team of CB_LANES threads owns one Opus stream.  On the
// GPU the team is a warp (CB_LANES = 32): strictly serial bitstream work runs on lane 0, every
// data-parallel loop is strided over the lanes and separated by __syncwarp().  The very same source
// is also compiled by g++ with CB_LANES = 1 ("host simulation", tests/hostsim/) so that the integer
// semantics can be checked against the oracle on a machine without a GPU.  The host simulation is a
// TEST TOOL: it is never linked into libconcentus_b200.so and the product library has no CPU path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_LANES 32
#define CB_DEV __device__ __forceinline__
#define CB_MEM __device__ __forceinline__
#define CB_DEV_NOINLINE __device__ __noinline__
#define CB_TABLE static __device__ const
#define CB_SYNC() __syncwarp()
#define CB_CLZ(x) __clz((int)(x))
#else
#define CB_LANES 1
#define CB_DEV static inline
#define CB_MEM inline
#define CB_DEV_NOINLINE static
#define CB_TABLE static const
#define CB_SYNC() ((void)0)
#define CB_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#endif

namespace cb {

// A "team" handle: lane index inside the team.  All team-wide helpers take it explicitly so that no
// code depends on threadIdx directly.
struct Team {
    int lane;
};

#if defined(__CUDACC__)
CB_DEV int team_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CB_DEV int team_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
CB_DEV int team_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
CB_DEV unsigned team_or(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CB_DEV int team_bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
CB_DEV unsigned team_bcast(unsigned v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#else
CB_DEV int team_sum(int v) { return v; }
CB_DEV int team_max(int v) { return v; }
CB_DEV int team_min(int v) { return v; }
CB_DEV unsigned team_or(unsigned v) { return v; }
CB_DEV int team_bcast(int v, int) { return v; }
CB_DEV unsigned team_bcast(unsigned v, int) { return v; }
#endif

}  // namespace cb

// for (i over [0,n)) distributed over the team
#define CB_TEAM_FOR(i, n, tm) for (int i = (tm).lane; i < (n); i += CB_LANES)
