// mlp_tile.cuh -- shared pieces of the fp32 MLP kernels (actor.cu: forward + sampling; ppo_update.cu: forward/backward):
// the packed weight image, TMA bulk-copy / mbarrier wrappers, the fast tanh and the 8 x CT FFMA2 register tile.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/satb200.h"

namespace mlp {

constexpr int IN = 18, HID = 256, ACTP = 4;     // ACTP: padded head count (actor 3, critic 1)
#ifndef SAT_TILE_CT
#define SAT_TILE_CT 8
#endif
// register tile per thread: RT rows x CT columns. CT = 16: 16 x 8 threads = 128 per CTA; CT = 8: 32 x 8 = 256 per CTA
// (measured: actor 190 us vs 201 us, fused PPO step 1.136 ms vs 1.200 ms - twice the warps hide the non-GEMM phases).
// A thread's columns are NC chunks of 4: j = c * CSTR + tx * 4 + q, so one float4 shared-memory load per chunk.
constexpr int RT = 8, CT = SAT_TILE_CT, NC = CT / 4, NP = CT / 2;
constexpr int TXN = HID / CT, CSTR = TXN * 4;
constexpr int M = 64, THREADS = TXN * (M / RT), NWARPS = THREADS / 32, KT = 16, NSTAGE = 2;
static_assert(CT == 16 || CT == 8, "column tile");
constexpr int H1_LD = M + 4;

// packed weight buffer (floats)
constexpr int OFF_W1T = 0;                      // [IN][HID]
constexpr int OFF_B1 = OFF_W1T + IN * HID;      // [HID]
constexpr int OFF_W2T = OFF_B1 + HID;           // [HID][HID]  (k-major: W2T[k][j] = fc2.weight[j][k])
constexpr int OFF_B2 = OFF_W2T + HID * HID;     // [HID]
constexpr int OFF_W3 = OFF_B2 + HID;            // [ACTP][HID]
constexpr int OFF_B3 = OFF_W3 + ACTP * HID;     // [ACTP]
constexpr int OFF_LS = OFF_B3 + ACTP;           // [ACTP]
constexpr int PACKED_FLOATS = OFF_LS + ACTP;
static_assert(PACKED_FLOATS == SAT_ACTOR_PACKED_FLOATS, "header constant out of sync");
static_assert((OFF_W2T * 4) % 16 == 0 && (OFF_W3 * 4) % 16 == 0, "TMA bulk sources must be 16-byte aligned");

struct __align__(128) ActorSmem {
    float h1T[HID * H1_LD];          // hidden-1 activations, transposed [k][m]
    float wt[NSTAGE][KT * HID];      // W2^T k-tiles; W1^T (IN x HID) aliases the start of this region
    float xT[IN * M];                // observations, transposed [k][m]
    float w3[ACTP * HID];
    float pre[M * ACTP];             // head pre-activations
    uint64_t full[NSTAGE];
    uint64_t bar_misc;
};
static_assert(IN * HID <= NSTAGE * KT * HID, "W1^T must fit in the stage region");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// tanh(x) = 1 - 2/(exp(2x)+1): MUFU.EX2 + MUFU.RCP, abs. error ~1e-7 (reference: torch.tanh fp32)
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float activate(float x, int use_tanh) { return use_tanh ? fast_tanh(x) : fmaxf(x, 0.0f); }

// acc[RT][CT] += A[k][m0..m0+RT) * B[k][cols], k in [0, K); A row stride lda, B row stride HID.
// The tile is held as RT x CT/2 float2 and updated with Blackwell's packed FFMA2 (fma.rn.f32x2): the same
// IEEE fp32 FMAs, two per issue slot, which leaves issue bandwidth for the LDS/address instructions.
template <int K>
__device__ __forceinline__ void tile_fma(float2 (&acc)[RT][NP], const float* __restrict__ A, int lda,
                                         const float* __restrict__ B, int ty, int tx) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        float a[RT];
#pragma unroll
        for (int i4 = 0; i4 < RT / 4; ++i4) {
            const float4 av = *reinterpret_cast<const float4*>(A + k * lda + ty * RT + i4 * 4);
            a[i4 * 4] = av.x; a[i4 * 4 + 1] = av.y; a[i4 * 4 + 2] = av.z; a[i4 * 4 + 3] = av.w;
        }
        float2 b[NP];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float4 bv = *reinterpret_cast<const float4*>(B + k * HID + c * CSTR + tx * 4);
            b[c * 2] = make_float2(bv.x, bv.y); b[c * 2 + 1] = make_float2(bv.z, bv.w);
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const float2 ai = make_float2(a[i], a[i]);
#pragma unroll
            for (int j = 0; j < NP; ++j) acc[i][j] = __ffma2_rn(ai, b[j], acc[i][j]);
        }
    }
}

// element (row i, local column c*4+q) of the packed accumulator tile
__device__ __forceinline__ float acc_at(const float2 (&acc)[RT][NP], int i, int c, int q) {
    const float2 v = acc[i][c * 2 + (q >> 1)];
    return (q & 1) ? v.y : v.x;
}


}  // namespace mlp
