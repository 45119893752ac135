// Launch helpers shared by the .cu files of libbetazero_b200.so
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/betazero_b200.h"

namespace bz {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? BZ_OK : -(1000 + (int)e); }

inline int launch_rc() { return cuda_rc(cudaGetLastError()); }

inline cudaStream_t as_stream(bz_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a grid-stride kernel over `items` work items: whole multiples of the SM count,
// enough CTAs to fill the machine (`ctas_per_sm` resident), never more than the work needs
inline int persistent_grid(int64_t items, int threads, int ctas_per_sm) {
    int64_t need = (items + threads - 1) / threads;
    int64_t full = (int64_t)kNumSMs * ctas_per_sm;
    if (need >= full) return (int)full;
    if (need <= 0) return 1;
    // round up to a multiple of the SM count when that does not overshoot the work by much
    int64_t r = ((need + kNumSMs - 1) / kNumSMs) * kNumSMs;
    return (int)(r <= need + kNumSMs / 2 ? r : need);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// streaming (read-once / write-once) 128-bit global accesses that do not allocate in L1
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const uint64_t *p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u64x2(uint64_t *p, ulonglong2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

}  // namespace bz
