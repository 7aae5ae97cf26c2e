// tc_mlp.cuh -- shared pieces of the tensor-core MLP kernels (actor_tc.cu: forward + sampling; ppo_fb_tc.cu: forward /
// backward of the fused PPO step): tile geometry, tcgen05 descriptors and wrappers, the exact bf16x3 operand split, TMEM
// loads, the packed-fp32 activations. sm_100a only. See actor_tc.cu for the scheme (six bf16 word products per fp32
// product, the exact 16-bit A_h B_h products in their own TMEM accumulator).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "mlp_tile.cuh"

namespace tcm {
using namespace mlp;

constexpr int TM = 128;                         // rows per CTA
constexpr int KC = 32;                          // K per chunk = one 64-byte swizzle row of bf16
constexpr int NCH = HID / KC;                   // 8 hidden-layer chunks
constexpr int NCHUNK = NCH + 1;                 // + layer 1 (K = 18 padded to 32) as chunk 0
constexpr int NPART = 4;                        // threads per row
constexpr int UPT = KC / NPART;                 // K values per thread per chunk: 8 = one 16-byte swizzle chunk
constexpr int TC_COMPUTE = TM * NPART;          // 512
constexpr int TC_THREADS = TC_COMPUTE + 64;     // + the MMA-issue warp and the weight-stream warp
constexpr int A_WORD = TM * 64;                 // one bf16 word (h, m or l) of an A chunk (32 k, SWIZZLE_64B rows): 8 KB
constexpr int A_STAGE = 3 * A_WORD;             // 24 KB
constexpr int NSUB = KC / 16;                   // weight sub-chunks (one UMMA k-step of 16) per chunk
constexpr int B_WORD = HID * 32;                // one word of a weight sub-chunk (16 k, SWIZZLE_32B rows): 8 KB
constexpr int B_STAGE = 3 * B_WORD;             // 24 KB
constexpr int NSA = 4, NSB = 4;                 // ring depths
constexpr uint32_t kTmemCols = 512;             // [0, 256): A_h B_h; [256, 512): the five small products
// tcgen05 instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D fp32 (1 << 4), A and B bf16 (1 << 7, 1 << 10),
// both K-major (bits 15, 16 = 0), N = 256 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// shared-memory matrix descriptors (SmemDescriptor: start >> 4, LBO = 1, SBO >> 4 at bit 32, version 1 at bit 46, layout type at
// bit 61). A: K-major SWIZZLE_64B (64-byte rows, 8-row groups 512 bytes apart, type 4); B: K-major SWIZZLE_32B (32-byte rows =
// exactly one UMMA k-step, 8-row groups 256 bytes apart, type 6)
__device__ __forceinline__ uint64_t umma_desc_a(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint64_t umma_desc_b(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// byte offset of (row, 16-byte chunk c16) inside one SWIZZLE_64B word buffer: Swizzle<2,4,3>, address bits [4,6) ^= bits [7,9)
__host__ __device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t c16) {
    return row * 64u + (((c16 ^ (row >> 1)) & 3u) << 4);
}
// the same for SWIZZLE_32B (c16 in {0, 1}): Swizzle<1,4,3>, address bit 4 ^= bit 7
__host__ __device__ __forceinline__ uint32_t sw32(uint32_t row, uint32_t c16) {
    return row * 32u + (((c16 ^ (row >> 2)) & 1u) << 4);
}
// exact split of two fp32 values into bf16 words by truncation: v = h + m + l, every word has <= 8 significant bits
// (h: the top 8 bits; v - h has <= 16, m its top 8; the rest has <= 8 and is a bf16 number). Packed as (v0 | v1 << 16).
__device__ __forceinline__ void split3(float v0, float v1, uint32_t& H, uint32_t& M, uint32_t& L) {
    const uint32_t b0 = __float_as_uint(v0), b1 = __float_as_uint(v1);
    H = __byte_perm(b0, b1, 0x7632);
    const float r0 = v0 - __uint_as_float(b0 & 0xffff0000u), r1 = v1 - __uint_as_float(b1 & 0xffff0000u);
    const uint32_t c0 = __float_as_uint(r0), c1 = __float_as_uint(r1);
    M = __byte_perm(c0, c1, 0x7632);
    const float q0 = r0 - __uint_as_float(c0 & 0xffff0000u), q1 = r1 - __uint_as_float(c1 & 0xffff0000u);
    L = __byte_perm(__float_as_uint(q0), __float_as_uint(q1), 0x7632);
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4& H, uint4& M, uint4& L) {
    split3(v[0], v[1], H.x, M.x, L.x); split3(v[2], v[3], H.y, M.y, L.y);
    split3(v[4], v[5], H.z, M.z, L.z); split3(v[6], v[7], H.w, M.w, L.w);
}
// 16 consecutive accumulator columns, no wait (pair with tmem_wait())
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// sum of the main and the small accumulator over the thread's 64 columns [c0, c0 + 64), 16 at a time (at most 32 transient registers)
__device__ __forceinline__ void tmem_sum64(uint32_t lane_base, int c0, float (&sacc)[64]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16], u[16];
        tmem_ld16_nowait(lane_base + (uint32_t)(c0 + q * 16), v);
        tmem_ld16_nowait(lane_base + (uint32_t)(HID + c0 + q * 16), u);
        tmem_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[q * 16 + j] = v[j] + u[j];
    }
}
constexpr float kTwoLog2e = 2.8853900817779268f;
// act() of two pre-activations on the packed fp32 pipe. TANH: tanh(x) = 1 - 2 / (2^(2 log2(e) x) + 1), the argument arrives
// already multiplied by 2 log2(e)
template <bool TANH>
__device__ __forceinline__ float2 act2_scaled(float2 t) {
    if (!TANH) return make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f));
    float2 e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    e = __fadd2_rn(e, make_float2(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(e.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(e.y));
    return __ffma2_rn(r, make_float2(-2.0f, -2.0f), make_float2(1.0f, 1.0f));
}

// The same with the reciprocal on the packed FMA pipe (bit-trick seed + three Newton steps, ~1 ulp like rcp.approx): the epilogue
// runs while the tensor core idles and two MUFU operations per value (16 per clock per SM) were its limit; this halves them.
// The argument is clamped so that 2^t + 1 stays a normal number for the seed.
template <bool TANH>
__device__ __forceinline__ float2 act2_scaled_fma(float2 t) {
    if (!TANH) return make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f));
    float2 e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(fminf(t.x, 64.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(fminf(t.y, 64.0f)));
    e = __fadd2_rn(e, make_float2(1.0f, 1.0f));
    r.x = __uint_as_float(0x7EF311C7u - __float_as_uint(e.x));
    r.y = __uint_as_float(0x7EF311C7u - __float_as_uint(e.y));
    const float2 ne = make_float2(-e.x, -e.y), one = make_float2(1.0f, 1.0f);
#pragma unroll
    for (int it = 0; it < 3; ++it) r = __ffma2_rn(r, __ffma2_rn(ne, r, one), r);
    return __ffma2_rn(r, make_float2(-2.0f, -2.0f), one);
}

// hidden unit held by layer-1 accumulator column c: thread quarter p = c >> 6 reads columns [64 p, 64 p + 64), and its value
// kc * 8 + j there must be unit kc * 32 + p * 8 + j (the 8 units it writes into chunk kc of the layer-2 A operand)
__host__ __device__ __forceinline__ int l1_unit(int c) { return ((c >> 3) & 7) * KC + (c >> 6) * UPT + (c & 7); }

// one arrival per warp once all its lanes are past their stores / TMEM loads
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

}  // namespace tcm
