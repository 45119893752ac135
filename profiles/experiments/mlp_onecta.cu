// Fused policy/value MLP inference for sm_100a: the whole 128 -> 256 -> 256 -> 256 -> (65 + 1) network of
// betazero_b200.net.PolicyValueMLP (the reference's TicTacToeNet family, SL/neural_networks.py:4-30, widened)
// in ONE launch instead of four library GEMMs.  This is the only dense contraction on the self-play path, so
// it is the only code here that uses the tensor cores: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by
// one thread, accumulators in TMEM, operands staged in shared memory in the canonical K-major SWIZZLE_128B
// layout, tcgen05.ld epilogue (bias + ReLU + bf16 round) that writes the next layer's A operand straight back
// into shared memory.  One CTA owns 128 leaves (rows); the activations never leave the SM between layers.
//
// Numerics = the PyTorch bf16 module: bf16 inputs/weights, fp32 accumulate, fp32 bias add, ReLU, round to bf16
// after every layer.  Output: bf16 [B, 72] = 65 policy logits, the pre-tanh value, zero padding -- exactly the
// BZ_PRIOR_LOGITS_BF16 input of the tree kernel.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace bz {
namespace {

constexpr int kRows = 128;       // leaves per CTA == UMMA M
constexpr int kIn = 128;         // 2 x 8 x 8 canonical planes
constexpr int kHidden = 256;
constexpr int kHeadRows = 80;    // 65 logits + value, padded to a legal UMMA N (multiple of 16)
constexpr int kOutStride = 72;   // row stride of the output (multiple of 8 elements: 16-byte rows)
constexpr int kThreads = 512;       // 16 warps: 4 per TMEM lane quadrant, each owns 64 accumulator columns in the epilogue
constexpr int kSlabA = kRows * 128;      // one 64-element K slab of A: 128 rows x 128 B
constexpr int kSlabB = kHidden * 128;    // one K slab of B: 256 rows x 128 B
constexpr int kSmemA = 4 * kSlabA;       // 64 KB
constexpr int kSmemB = 4 * kSlabB;       // 128 KB
constexpr int kSmemBias = kHidden * 4;   // 1 KB
constexpr int kSmemTotal = kSmemA + kSmemB + kSmemBias + 128 + 1024;  // + barriers/tmem slot + alignment slack
constexpr int kTmemCols = 256;

// rows x K bf16, row-major in global (row stride `ld` elements) -> K-major SWIZZLE_128B slabs in shared memory:
// 16-byte chunk j of row r in slab s lands at s*slab_bytes + r*128 + ((j ^ (r & 7)) * 16)
__device__ __forceinline__ void load_operand(uint32_t sbase, int slab_bytes, const __nv_bfloat16 *g, int rows, int K, int ld,
                                             int valid_rows) {
    const int chunks_per_row = K / 8;
    const int total = rows * chunks_per_row;
    for (int i = threadIdx.x; i < total; i += kThreads) {
        const int r = i / chunks_per_row, c = i - r * chunks_per_row;
        const int s = c >> 3, j = c & 7;
        const uint32_t dst = sbase + s * slab_bytes + r * 128 + ((j ^ (r & 7)) << 4);
        const bool ok = r < valid_rows;
        cp_async16(dst, g + (size_t)(ok ? r : 0) * ld + c * 8, ok);
    }
}

struct MlpParams {
    const __nv_bfloat16 *x;                 // [B, 128]
    const __nv_bfloat16 *w[4], *b[4];       // [256,128] [256,256] [256,256] [80,256]; biases 256/256/256/80
    __nv_bfloat16 *out;                     // [B, 72]
    int B;
    const uint8_t *wimg;                    // optional: 14 units of 32 KB (layer, K slab) already in the shared-memory
                                            // layout -> weights arrive by cp.async.bulk (TMA) instead of 16-byte LDGSTS
};

__device__ __forceinline__ int image_unit_of_layer(int layer) { return layer == 0 ? 0 : 2 + (layer - 1) * 4; }

__global__ void __launch_bounds__(kThreads, 1) mlp_kernel(const MlpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B operands need 1024-byte alignment
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t sA = base, sB = base + kSmemA;
    float *sBias = reinterpret_cast<float *>(smem + kSmemA + kSmemB);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kSmemA + kSmemB + kSmemBias);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const uint32_t bar_addr = smem_u32(bar), wbar_addr = smem_u32(bar + 1);
    const bool bulk = p.wimg != nullptr;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kRows;
    const int valid_rows = min(kRows, p.B - row0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        mbar_init(bar_addr, 1);
        mbar_init(wbar_addr, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (bulk) {  // W1: two 32 KB K slabs, one bulk copy
            mbar_expect_tx(wbar_addr, 2 * kSlabB);
            bulk_load(sB, p.wimg, 2 * kSlabB, wbar_addr);
        }
    }
    // layer-0 operands: W1 does not depend on the previous kernel; the leaf planes do (PDL: wait for the tree
    // kernel that wrote them, then let the next tree kernel start its own prologue)
    if (!bulk) load_operand(sB, kSlabB, p.w[0], kHidden, kIn, kIn, kHidden);
    pdl_wait();
    pdl_launch_dependents();
    load_operand(sA, kSlabA, p.x + (size_t)row0 * kIn, kRows, kIn, kIn, valid_rows);
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

#pragma unroll 1
    for (int layer = 0; layer < 4; ++layer) {
        const int K = layer == 0 ? kIn : kHidden;
        const int N = layer == 3 ? kHeadRows : kHidden;
        if (threadIdx.x == 0) {
            if (bulk) {  // this layer's weights were sent by cp.async.bulk: wait for them to land
                mbar_wait(wbar_addr, (uint32_t)(layer & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t idesc = umma_idesc(kRows, N);
            for (int k = 0; k < K / 16; ++k) {  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle atom
                const uint32_t off = (uint32_t)(k >> 2), kk = (uint32_t)(k & 3) * 32u;
                umma_bf16(tmem, umma_desc(sA + off * kSlabA + kk), umma_desc(sB + off * kSlabB + kk), idesc, k > 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
        }
        // bias of this layer -> shared (overlaps the MMAs)
        for (int i = threadIdx.x; i < N; i += kThreads) sBias[i] = __bfloat162float(p.b[layer][i]);
        mbar_wait(bar_addr, (uint32_t)(layer & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();  // bias visible; A and B are free again (the MMAs have read them)
        if (layer < 3) {  // next layer's weights stream in while the epilogue runs
            const int nextN = layer == 2 ? kHeadRows : kHidden;
            if (!bulk) {
                load_operand(sB, kSlabB, p.w[layer + 1], nextN, kHidden, kHidden, nextN);
            } else if (threadIdx.x == 32) {
                const uint8_t *src = p.wimg + (size_t)image_unit_of_layer(layer + 1) * kSlabB;
                if (layer + 1 < 3) {  // four 32 KB K slabs, contiguous in the image and in shared memory
                    mbar_expect_tx(wbar_addr, 4 * kSlabB);
                    bulk_load(sB, src, 4 * kSlabB, wbar_addr);
                } else {  // head: 80 rows per slab
                    mbar_expect_tx(wbar_addr, 4 * kHeadRows * 128);
                    for (int sl = 0; sl < 4; ++sl) bulk_load(sB + sl * kSlabB, src + (size_t)sl * kSlabB, kHeadRows * 128, wbar_addr);
                }
            }
        }
        // epilogue: thread <-> accumulator row; warp w reads TMEM lanes 32*(w%4).., column quarter w/4
        const int r = (warp & 3) * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        if (layer < 3) {
            const int c_begin = (warp >> 2) * (kHidden / 4);
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_begin + kHidden / 4; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(trow + (uint32_t)c0, acc);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float a = fmaxf(__uint_as_float(acc[2 * j]) + sBias[c0 + 2 * j], 0.f);
                    const float b = fmaxf(__uint_as_float(acc[2 * j + 1]) + sBias[c0 + 2 * j + 1], 0.f);
                    packed[j] = pack_bf16(a, b);
                }
                // 32 columns = 4 chunks of 16 B in slab c0/64, chunk index (c0%64)/8 + q, swizzled by the row
                const uint32_t rowbase = sA + (uint32_t)(c0 >> 6) * kSlabA + (uint32_t)r * 128u;
                const int j0 = (c0 & 63) >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t dst = rowbase + (uint32_t)(((j0 + q) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * q]), "r"(packed[4 * q + 1]),
                                 "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                                 : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        } else if (warp < 12) {
            // head: 72 output columns (65 logits, value, padding) -> global, bf16, no activation;
            // column group w/4 takes one 32-column chunk
            __nv_bfloat16 *orow = p.out + (size_t)(row0 + r) * kOutStride;
            {
                const int c0 = (warp >> 2) * 32;
                uint32_t acc[32];
                tmem_ld32(trow + (uint32_t)c0, acc);  // columns >= 80 of the last chunk are stale accumulators: never stored
                if (r < valid_rows) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int c = c0 + q * 8;
                        if (c < kOutStride) {
                            uint4 v;
                            v.x = pack_bf16(__uint_as_float(acc[q * 8 + 0]) + sBias[c + 0], __uint_as_float(acc[q * 8 + 1]) + sBias[c + 1]);
                            v.y = pack_bf16(__uint_as_float(acc[q * 8 + 2]) + sBias[c + 2], __uint_as_float(acc[q * 8 + 3]) + sBias[c + 3]);
                            v.z = pack_bf16(__uint_as_float(acc[q * 8 + 4]) + sBias[c + 4], __uint_as_float(acc[q * 8 + 5]) + sBias[c + 5]);
                            v.w = pack_bf16(__uint_as_float(acc[q * 8 + 6]) + sBias[c + 6], __uint_as_float(acc[q * 8 + 7]) + sBias[c + 7]);
                            *reinterpret_cast<uint4 *>(orow + c) = v;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}


// =====================================================================================================
// v2: the same network as a warp-specialised pipeline.
//   warp 0      MMA issuer (one thread): per layer, per N-half, per 64-wide K slab: 4 x tcgen05.mma
//   warp 1      producer (one thread): streams the weights as 16 KB units (one N-half x one K slab) from a
//               host-prepared image that already has the SWIZZLE_128B shared-memory layout, one
//               cp.async.bulk per unit into a 4-stage ring (mbarrier complete_tx)
//   warps 2-17  epilogue: 4 warps per TMEM lane quadrant, each owns 64 accumulator columns = one K slab of the
//               NEXT layer's A operand: tcgen05.ld -> +bias, ReLU, bf16 -> swizzled st.shared -> arrive
// Accumulators are double buffered in TMEM (2 x 256 columns) and activations in shared memory (2 x 64 KB), so
// layer L+1's MMAs start on K slab s as soon as layer L's epilogue has produced it, and the epilogue of N-half 0
// overlaps the MMAs of N-half 1.
// =====================================================================================================
constexpr int kEpiWarps = 16;
constexpr int kThreads2 = (2 + kEpiWarps) * 32;
constexpr int kUnitBytes = 128 * 128;           // 128 weight rows x one 64-element K slab
constexpr int kHeadUnitBytes = kHeadRows * 128;  // the head has 80 rows
#ifndef BZ_MLP_STAGES
#define BZ_MLP_STAGES 6
#endif
constexpr int kStages = BZ_MLP_STAGES;  // 16 KB each; 6 = all the shared memory left beside the two activation buffers
constexpr int kNumUnits = 4 + 8 + 8 + 4;
constexpr int kSmem2 = 2 * kSmemA + kStages * kUnitBytes + 512 + 1024;

#ifdef BZ_MLP_TRACE
// debug timeline of CTA 0 (clock64 at key events); built only for profiling experiments
__device__ long long g_mlp_trace[64];
#define MLP_TRACE(i) do { if (blockIdx.x == 0) g_mlp_trace[i] = clock64(); } while (0)
#else
#define MLP_TRACE(i) do { } while (0)
#endif

struct Mlp2Params {
    const __nv_bfloat16 *x;     // [B, 128]
    const uint8_t *wimg;        // kNumUnits x 16 KB, units in consumption order, SWIZZLE_128B image
    const float *bias;          // [256 + 256 + 256 + 80] fp32
    __nv_bfloat16 *out;         // [B, 72]
    int B;
};

__global__ void __launch_bounds__(kThreads2, 1) mlp_pipe_kernel(const Mlp2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t sA0 = base, sRing = base + 2 * kSmemA;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * kSmemA + kStages * kUnitBytes);
    // barrier map: full[kStages] empty[kStages] a_ready[2][4] d_ready[2][2]
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto EMPTY = [&](int i) { return bar0 + 8u * (uint32_t)(kStages + i); };
    auto AREADY = [&](int b, int sl) { return bar0 + 8u * (uint32_t)(2 * kStages + b * 4 + sl); };
    auto DREADY = [&](int b, int h) { return bar0 + 8u * (uint32_t)(2 * kStages + 8 + b * 2 + h); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kRows;
    const int valid_rows = min(kRows, p.B - row0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int i = 0; i < kStages; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
        for (int i = 0; i < 8; ++i) mbar_init(AREADY(i >> 2, i & 3), 4);  // one elected lane of each of the 4 quadrant warps
        for (int i = 0; i < 4; ++i) mbar_init(DREADY(i >> 1, i & 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) MLP_TRACE(0);

    if (warp == 1) {
        // ---------------- producer: weights never depend on the previous kernel, start at once ----------------
        if (lane == 0) {
            for (int u = 0; u < kNumUnits; ++u) {
                const int st = u % kStages;
                if (u >= kStages) mbar_wait(EMPTY(st), (uint32_t)(((u / kStages) - 1) & 1));
                const uint32_t bytes = u >= 20 ? kHeadUnitBytes : kUnitBytes;
                mbar_expect_tx(FULL(st), bytes);
                bulk_load(sRing + (uint32_t)st * kUnitBytes, p.wimg + (size_t)u * kUnitBytes, bytes, FULL(st));
            }
        }
    } else if (warp == 0) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            uint32_t a_par = 0;  // bit (b*4+s): phase parity of a_ready[b][s]
            int u = 0;
            for (int L = 0; L < 4; ++L) {
                const int b = L & 1, nh = L == 3 ? 1 : 2, ns = L == 0 ? 2 : 4, N = L == 3 ? kHeadRows : 128;
                const uint32_t idesc = umma_idesc(kRows, N);
                const uint32_t sA = sA0 + (uint32_t)b * kSmemA;
                for (int h = 0; h < nh; ++h) {
                    for (int sl = 0; sl < ns; ++sl, ++u) {
                        const int st = u % kStages;
                        mbar_wait(FULL(st), (uint32_t)((u / kStages) & 1));
                        if (h == 0) {
                            mbar_wait(AREADY(b, sl), (a_par >> (b * 4 + sl)) & 1u);
                            a_par ^= 1u << (b * 4 + sl);
                        }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t d = tmem + (uint32_t)(b * 256 + h * 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(d, umma_desc(sA + (uint32_t)sl * kSlabA + kk * 32u),
                                      umma_desc(sRing + (uint32_t)st * kUnitBytes + kk * 32u), idesc, (sl | kk) != 0);
                        umma_commit(EMPTY(st));  // the ring stage is free once these MMAs have read it
                    }
                    umma_commit(DREADY(b, h));  // accumulator half complete
                    MLP_TRACE(1 + L * 2 + h);  // MMA issue of (L, h) finished
                }
            }
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int e = warp - 2;
        const int q = warp & 3;   // TMEM lane quadrant this warp may access
        const int cg = e >> 2;    // 64-column group == K slab of the next layer's A operand
        const int r = q * 32 + lane;
        // layer-0 input: the leaf planes (written by the previous kernel: PDL wait first)
        pdl_wait();
        pdl_launch_dependents();
        if (cg < 2) {  // slab cg of A[0]: this warp loads its 32 rows x 8 chunks
            const bool ok = r < valid_rows;
            const __nv_bfloat16 *src = p.x + (size_t)(row0 + (ok ? r : 0)) * kIn + cg * 64;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                cp_async16(sA0 + (uint32_t)cg * kSlabA + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4), src + j * 8, ok);
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(AREADY(0, cg));
            if (e == 0 && lane == 0) MLP_TRACE(10);  // x loaded
        }
        uint32_t d_par = 0;  // bit b: parity of this warp's d_ready[b][h]
        const int h = cg >> 1;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        for (int L = 0; L < 3; ++L) {
            const int b = L & 1;
            const float *bias = p.bias + L * kHidden;
            mbar_wait(DREADY(b, h), (d_par >> b) & 1u);
            d_par ^= 1u << b;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0 && (e == 0 || e == 8)) MLP_TRACE(11 + L * 4 + (e ? 2 : 0));  // D ready seen by epilogue (h0 / h1)
            const uint32_t dstA = sA0 + (uint32_t)(b ^ 1) * kSmemA + (uint32_t)cg * kSlabA + (uint32_t)r * 128u;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int c0 = cg * 64 + half * 32;
                uint32_t acc[32];
                tmem_ld32(trow + (uint32_t)(b * 256 + c0), acc);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bv = *reinterpret_cast<const float4 *>(bias + c0 + 4 * j);
                    const float v0 = fmaxf(__uint_as_float(acc[4 * j + 0]) + bv.x, 0.f), v1 = fmaxf(__uint_as_float(acc[4 * j + 1]) + bv.y, 0.f);
                    const float v2 = fmaxf(__uint_as_float(acc[4 * j + 2]) + bv.z, 0.f), v3 = fmaxf(__uint_as_float(acc[4 * j + 3]) + bv.w, 0.f);
                    packed[2 * j] = pack_bf16(v0, v1);
                    packed[2 * j + 1] = pack_bf16(v2, v3);
                }
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const uint32_t dst = dstA + (uint32_t)(((half * 4 + qq) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * qq]), "r"(packed[4 * qq + 1]),
                                 "r"(packed[4 * qq + 2]), "r"(packed[4 * qq + 3])
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(AREADY(b ^ 1, cg));
            if (lane == 0 && (e == 0 || e == 8)) MLP_TRACE(12 + L * 4 + (e ? 2 : 0));  // epilogue slab done
        }
        // head (layer 3, accumulator buffer 1, one N block of 80 columns): 72 output columns -> global
        if (cg < 2) {
            mbar_wait(DREADY(1, 0), (d_par >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float *bias = p.bias + 3 * kHidden;
            __nv_bfloat16 *orow = p.out + (size_t)(row0 + r) * kOutStride;
            for (int half = 0; half < 2; ++half) {
                const int c0 = cg * 64 + half * 32;
                if (c0 >= 96) break;  // warp-uniform
                uint32_t acc[32];
                tmem_ld32(trow + (uint32_t)(256 + c0), acc);  // columns >= 80 are stale: never stored
                if (r < valid_rows) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int c = c0 + qq * 8;
                        if (c < kOutStride) {
                            const float4 b0 = *reinterpret_cast<const float4 *>(bias + c), b1 = *reinterpret_cast<const float4 *>(bias + c + 4);
                            uint4 v;
                            v.x = pack_bf16(__uint_as_float(acc[qq * 8 + 0]) + b0.x, __uint_as_float(acc[qq * 8 + 1]) + b0.y);
                            v.y = pack_bf16(__uint_as_float(acc[qq * 8 + 2]) + b0.z, __uint_as_float(acc[qq * 8 + 3]) + b0.w);
                            v.z = pack_bf16(__uint_as_float(acc[qq * 8 + 4]) + b1.x, __uint_as_float(acc[qq * 8 + 5]) + b1.y);
                            v.w = pack_bf16(__uint_as_float(acc[qq * 8 + 6]) + b1.z, __uint_as_float(acc[qq * 8 + 7]) + b1.w);
                            *reinterpret_cast<uint4 *>(orow + c) = v;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) MLP_TRACE(30);
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace
}  // namespace bz

using namespace bz;

extern "C" int bz_mlp_forward(const void *x_bf16, const void *w1, const void *b1, const void *w2, const void *b2, const void *w3,
                              const void *b3, const void *w_head, const void *b_head, void *out_bf16, int64_t n, int in_features,
                              int hidden, int head_rows, int out_stride, bz_stream_t stream) {
    if (n < 0 || in_features != kIn || hidden != kHidden || head_rows != kHeadRows || out_stride != kOutStride) return BZ_ERR_ARG;
    if (n && (!x_bf16 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w_head || !b_head || !out_bf16)) return BZ_ERR_ARG;
    if (!aligned16(x_bf16) || !aligned16(w1) || !aligned16(w2) || !aligned16(w3) || !aligned16(w_head) || !aligned16(out_bf16))
        return BZ_ERR_UNALIGNED;
    if (n == 0) return BZ_OK;
    static bool configured[64] = {};
    {
        cudaError_t e = allow_dynamic_smem(mlp_kernel, kSmemTotal, configured);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    MlpParams p;
    p.x = (const __nv_bfloat16 *)x_bf16;
    p.w[0] = (const __nv_bfloat16 *)w1; p.b[0] = (const __nv_bfloat16 *)b1;
    p.w[1] = (const __nv_bfloat16 *)w2; p.b[1] = (const __nv_bfloat16 *)b2;
    p.w[2] = (const __nv_bfloat16 *)w3; p.b[2] = (const __nv_bfloat16 *)b3;
    p.w[3] = (const __nv_bfloat16 *)w_head; p.b[3] = (const __nv_bfloat16 *)b_head;
    p.out = (__nv_bfloat16 *)out_bf16;
    p.B = (int)n;
    p.wimg = nullptr;
    cudaError_t e = launch_kernel(mlp_kernel, dim3((unsigned)((n + kRows - 1) / kRows)), dim3(kThreads), (size_t)kSmemTotal,
                                  as_stream(stream), pdl_enabled(), p);
    if (e != cudaSuccess) return cuda_rc(e);
    return launch_rc();
}

extern "C" int bz_mlp_forward_packed(const void *x_bf16, const void *weight_image, const void *bias_f32, void *out_bf16,
                                     int64_t n, bz_stream_t stream) {
    if (n < 0 || (n && (!x_bf16 || !weight_image || !bias_f32 || !out_bf16))) return BZ_ERR_ARG;
    if (!aligned16(x_bf16) || !aligned16(weight_image) || !aligned16(bias_f32) || !aligned16(out_bf16)) return BZ_ERR_UNALIGNED;
    if (n == 0) return BZ_OK;
    static bool configured[64] = {};
    {
        cudaError_t e = allow_dynamic_smem(mlp_pipe_kernel, kSmem2, configured);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    Mlp2Params p;
    p.x = (const __nv_bfloat16 *)x_bf16;
    p.wimg = (const uint8_t *)weight_image;
    p.bias = (const float *)bias_f32;
    p.out = (__nv_bfloat16 *)out_bf16;
    p.B = (int)n;
    cudaError_t e = launch_kernel(mlp_pipe_kernel, dim3((unsigned)((n + kRows - 1) / kRows)), dim3(kThreads2), (size_t)kSmem2,
                                  as_stream(stream), pdl_enabled(), p);
    if (e != cudaSuccess) return cuda_rc(e);
    return launch_rc();
}

extern "C" int64_t bz_mlp_weight_image_bytes(void) { return (int64_t)kNumUnits * kUnitBytes; }

#ifdef BZ_MLP_TRACE
extern "C" int bz_mlp_debug_trace(long long *host_out) {
    return cuda_rc(cudaMemcpyFromSymbol(host_out, g_mlp_trace, sizeof(long long) * 64));
}
#endif

// bz_mlp_forward with TMA weight streaming: biases as in bz_mlp_forward (bf16), weights as an image of 14 units of
// 32 KB (layer 0: 2 K slabs, layers 1 and 2: 4, head: 4 with 80 rows each), each unit in the shared-memory layout
extern "C" int bz_mlp_forward_image(const void *x_bf16, const void *weight_image32, const void *b1, const void *b2, const void *b3,
                                    const void *b_head, void *out_bf16, int64_t n, bz_stream_t stream) {
    if (n < 0 || (n && (!x_bf16 || !weight_image32 || !b1 || !b2 || !b3 || !b_head || !out_bf16))) return BZ_ERR_ARG;
    if (!aligned16(x_bf16) || !aligned16(weight_image32) || !aligned16(out_bf16)) return BZ_ERR_UNALIGNED;
    if (n == 0) return BZ_OK;
    static bool configured[64] = {};
    {
        cudaError_t e = allow_dynamic_smem(mlp_kernel, kSmemTotal, configured);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    MlpParams p = {};
    p.x = (const __nv_bfloat16 *)x_bf16;
    p.b[0] = (const __nv_bfloat16 *)b1; p.b[1] = (const __nv_bfloat16 *)b2;
    p.b[2] = (const __nv_bfloat16 *)b3; p.b[3] = (const __nv_bfloat16 *)b_head;
    p.out = (__nv_bfloat16 *)out_bf16;
    p.B = (int)n;
    p.wimg = (const uint8_t *)weight_image32;
    cudaError_t e = launch_kernel(mlp_kernel, dim3((unsigned)((n + kRows - 1) / kRows)), dim3(kThreads), (size_t)kSmemTotal,
                                  as_stream(stream), pdl_enabled(), p);
    if (e != cudaSuccess) return cuda_rc(e);
    return launch_rc();
}
