// actor_tc.cu -- K3 on the 5th-generation tensor cores: both dense layers of Actor_Gaussian (18 -> 256 -> 256) as
// error-free-split BF16x3 products on tcgen05.mma with the accumulators in TMEM (sm_100a only).
//
// Reference behaviour replaced: the same as actor.cu (PPO_continuous.choose_action -> Actor_Gaussian.forward,
// ppo_continuous.py:83-95, 176-189), which the reference evaluates in fp32.
//
// Why: config 3 "with fused actor sampling" spends 67 % of its step in the two FFMA2 actor launches (2 x 193 us at
// 65 536 rows, 65 % of the nominal fp32 rate: the CUDA-core ceiling). The hidden layer is a 65 536 x 256 x 256
// contraction. A plain bf16 / TF32 product would be narrower than the reference's fp32, so every fp32 operand is split
// EXACTLY into three bf16 words (x = h + m + l, 8 significant bits each, by truncation) and six of the nine word
// products are accumulated in fp32 in TMEM:
//     A B ~= A_h B_h  +  (A_h B_m + A_m B_h)  +  (A_m B_m + A_h B_l + A_l B_h)        (dropped: ~2^-24 relative).
// Round 2's first version used two TF32 words (3xTF32): its error against an fp64 ground truth was 1.6x the FFMA2
// kernel's, because the tensor core's fp32 accumulation TRUNCATES and every 22-bit TF32 x TF32 product loses bits when
// it is added to a larger accumulator (32 such additions per output). A bf16 x bf16 product has 16 significant bits,
// so the large A_h B_h term is added (nearly always) exactly, in 16 instead of 32 accumulations (UMMA K = 16), and all
// the rounding happens in the SECOND accumulator, whose content is 2^-8 of the result. Same number of MMA instructions
// (6 products x 2 k-steps per 32-unit chunk instead of 3 x 4) and 25 % fewer operand bytes (6 instead of 8 per value).
// tests/test_gpu_actor_tc.py measures the error of this path and of the FFMA2 path against an fp64 ground truth.
//
// Persistent kernel, one CTA per SM; a row tile = 128 observations (= the 128 TMEM lanes), 544 threads:
//   warps 0-15 (thread = row r = tid & 127, quarter p = tid >> 7):
//             chunk 0: rebuild / load 8 of the row's observation values, split, write the layer-1 A operand (K = 18
//             padded to 32); when the layer-1 MMAs are complete, tcgen05.ld 64 pre-activations (the W1 image permutes
//             the output units so that a thread's 64 units - 8 per later chunk - are 64 CONSECUTIVE accumulator columns);
//             chunks 1..8: bias, tanh, split, the K-major 64-byte-swizzled A operand of 32 hidden units;
//             epilogue: tcgen05.ld of 64 of the row's accumulator columns (a warp may only touch TMEM lanes
//             32 (w % 4) .. +31, which is exactly its rows), bias, tanh, partial head dot products, reduced over the four
//             quarters through shared memory; thread (r, p < 3) then finishes action p: Philox sample, clamp, log-prob;
//   warp 16, one elected lane: streams the pre-split, pre-swizzled weight image (48 KB per chunk: W1 then the eight
//             K-chunks of W2, h | m | l) through a three-stage shared-memory ring with cp.async.bulk + mbarrier and issues
//             the 12 tcgen05.mma (kind::f16, M 128, N 256, K 16) of each chunk; tcgen05.commit hands the stages back and
//             signals "layer 1 complete" / "layer 2 complete".
// TMEM: columns [0, 256) A_h B_h, [256, 512) the five small products; layer 1 uses the same columns before layer 2.
// Across tiles: the weight ring never drains, the next tile's observation operand is written while the current tile's last
// chunks are on the tensor core, and its layer-1 MMAs start as soon as the epilogue has the accumulators in registers.
// Shared memory: A 2 x 24 KB, X 24 KB, B 3 x 48 KB, head table 4 KB, bias 1 KB = 222 KB -> one CTA per SM.
#include <cstdint>
#include <cuda_runtime.h>
#include "sat_math.cuh"
#include "mlp_tile.cuh"

namespace {
using namespace mlp;

constexpr int TM = 128;                         // rows per CTA
constexpr int KC = 32;                          // K per chunk = one 64-byte swizzle row of bf16
constexpr int NCH = HID / KC;                   // 8 hidden-layer chunks
constexpr int NCHUNK = NCH + 1;                 // + layer 1 (K = 18 padded to 32) as chunk 0
constexpr int NPART = 4;                        // threads per row
constexpr int UPT = KC / NPART;                 // K values per thread per chunk: 8 = one 16-byte swizzle chunk
constexpr int TC_COMPUTE = TM * NPART;          // 512
constexpr int TC_THREADS = TC_COMPUTE + 32;
constexpr int A_WORD = TM * 64;                 // one bf16 word (h, m or l) of an A chunk: 8 KB
constexpr int A_STAGE = 3 * A_WORD;
constexpr int B_WORD = HID * 64;                // 16 KB
constexpr int B_STAGE = 3 * B_WORD;             // 48 KB
constexpr int NSA = 2, NSB = 3;                 // ring depths
constexpr int OFF_A = 0;
constexpr int OFF_X = OFF_A + NSA * A_STAGE;    // the observation operand of the NEXT tile's layer 1 (written during this tile)
constexpr int OFF_B = OFF_X + A_STAGE;
constexpr int OFF_W3P = OFF_B + NSB * B_STAGE;  // float4 [HID]: (W3[0][c], W3[1][c], W3[2][c], b2[c])
constexpr int OFF_B1P = OFF_W3P + HID * 16;     // b1 in layer-1 accumulator column order
constexpr int OFF_BAR = OFF_B1P + HID * 4;      // mbarriers
constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
constexpr int TC_SMEM = OFF_TMEM + 16 + 1024;   // + slack for the 1024-byte alignment of the swizzled tiles
static_assert(TC_SMEM <= 227 * 1024, "shared memory budget");
static_assert(NPART * TM * 16 <= NSA * A_STAGE, "the head reduction scratch aliases the A ring");
static_assert(NCHUNK * B_STAGE <= SAT_ACTOR_TC_IMAGE_FLOATS * 4, "weight image larger than the caller's scratch");
constexpr uint32_t kTmemCols = 512;             // [0, 256): A_h B_h; [256, 512): the five small products
// tcgen05 instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D fp32 (1 << 4), A and B bf16 (1 << 7, 1 << 10),
// both K-major (bits 15, 16 = 0), N = 256 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// shared-memory matrix descriptor of a K-major SWIZZLE_64B operand (64-byte rows) whose 8-row groups are 512 bytes apart
// (SmemDescriptor: start >> 4, LBO = 1, SBO = 512 >> 4 at bit 32, version 1 at bit 46, layout type 4 at bit 61)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
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
// tanh(x) = 1 - 2 / (exp(2x) + 1) on the two MUFU units, 5 instructions, abs. error ~1e-7 (inf / 0 saturate to +-1 by themselves)
__device__ __forceinline__ float tanh5(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}
template <bool TANH>
__device__ __forceinline__ float act_fn(float x) { return TANH ? tanh5(x) : fmaxf(x, 0.0f); }

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// hidden unit held by layer-1 accumulator column c: thread quarter p = c >> 6 reads columns [64 p, 64 p + 64), and its value
// kc * 8 + j there must be unit kc * 32 + p * 8 + j (the 8 units it writes into chunk kc of the layer-2 A operand)
__host__ __device__ __forceinline__ int l1_unit(int c) { return ((c >> 3) & 7) * KC + (c >> 6) * UPT + (c & 7); }

// The weight image: chunk 0 = W1 (n = layer-1 accumulator column, k = observation dimension, zero beyond 18), chunks 1..8 =
// the K-chunks of W2 (n = output unit); per chunk [h | m | l] x 256 rows x 64 bytes (32 bf16 along k), 8-row groups of 512
// bytes with the 16-byte column chunks XOR-swizzled: exactly the bytes the SWIZZLE_64B descriptor expects, so one linear 48 KB
// bulk copy per chunk brings it in. packed: the FFMA kernel's image (W1T[k][j], W2T[k][n] = fc2.weight[n][k]).
// One thread = 8 consecutive k of one n (coalesced over n).
__global__ void actor_tc_pack_kernel(const float* __restrict__ packed, unsigned char* __restrict__ image) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NCHUNK * 4 * HID) return;
    const int n = idx % HID, c16 = (idx / HID) & 3, chunk = idx / (4 * HID);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kk = c16 * 8 + j;
        if (chunk == 0) v[j] = (kk < IN) ? packed[OFF_W1T + kk * HID + l1_unit(n)] : 0.0f;
        else v[j] = packed[OFF_W2T + ((chunk - 1) * KC + kk) * HID + n];
    }
    uint4 H, M, L;
    split8(v, H, M, L);
    unsigned char* base = image + (size_t)chunk * B_STAGE + sw64((uint32_t)n, (uint32_t)c16);
    *reinterpret_cast<uint4*>(base) = H;
    *reinterpret_cast<uint4*>(base + B_WORD) = M;
    *reinterpret_cast<uint4*>(base + 2 * B_WORD) = L;
}

// the row's observation dimensions [8 part, 8 part + 8), split into the layer-1 A operand (one 16-byte chunk of each word buffer)
__device__ __forceinline__ void produce_x(unsigned char* xs, uint32_t a_off, int part, int64_t g, bool live,
                                          const float* __restrict__ obs_f32, const SatEnvState& st,
                                          const double* __restrict__ obs_stats, float* __restrict__ obs_out) {
    float xv[UPT];
#pragma unroll
    for (int j = 0; j < UPT; ++j) {
        const int k = part * UPT + j;
        float val = 0.0f;
        if (k < IN) {
            if (obs_f32) val = obs_f32[g * IN + k];
            else {
                // rebuild the observation from the fp64 SoA env state (environment.py:76-77: P - E, Pv - Ev, P, Pv, E, Ev),
                // normalise in fp64 (normalization.py:41)
                const int64_t ld = st.ld;
                double y = (k < 6) ? st.state[k * ld + g] - st.state[(k + 6) * ld + g] : st.state[(k - 6) * ld + g];
                if (obs_stats) y = (y - obs_stats[1 + k]) / (obs_stats[1 + 2 * IN + k] + 1e-8);
                val = (float)y;
            }
            if (obs_out && live) obs_out[g * IN + k] = val;
        }
        xv[j] = val;
    }
    uint4 H, M, L;
    split8(xv, H, M, L);
    unsigned char* a0 = xs + a_off;
    *reinterpret_cast<uint4*>(a0) = H;
    *reinterpret_cast<uint4*>(a0 + A_WORD) = M;
    *reinterpret_cast<uint4*>(a0 + 2 * A_WORD) = L;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core
}
// one arrival per warp once all its lanes are past their stores / TMEM loads
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// Persistent: one CTA per SM walks over row tiles blockIdx.x, blockIdx.x + gridDim.x, ... The weight chunks stream through the
// B ring continuously across tiles; the next tile's observation operand is produced while the last hidden-layer MMAs of the
// current tile run, and its layer-1 MMAs start as soon as the epilogue has READ the accumulators (the epilogue arithmetic and
// the sampling overlap them).
template <bool TANH>
__global__ void __launch_bounds__(TC_THREADS, 1)
actor_tc_kernel(const float* __restrict__ packed, const unsigned char* __restrict__ image, const float* __restrict__ obs_f32,
                const SatEnvState st, const double* __restrict__ obs_stats, int64_t n, int64_t row_offset, uint64_t seed,
                uint64_t step, float max_action, const float* __restrict__ eps_in, float* __restrict__ act,
                float* __restrict__ logp, float* __restrict__ mean_out, float* __restrict__ eps_out, float* __restrict__ obs_out) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the swizzled tiles by pointer arithmetic on the shared array (an integer round trip would turn
    // every access into a generic-space LD/ST)
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    float4* w3p = reinterpret_cast<float4*>(sm + OFF_W3P);
    float* b1p = reinterpret_cast<float*>(sm + OFF_B1P);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* b_full = bars;          // [NSB] bulk copy of a weight chunk landed
    uint64_t* b_empty = bars + 3;     // [NSB] the chunk's MMAs have read the B stage
    uint64_t* a_full = bars + 6;      // [NSA] 128 rows of the A chunk written
    uint64_t* a_empty = bars + 8;     // [NSA] the chunk's MMAs have read the A stage
    uint64_t* x_full = bars + 10;     // the tile's observation operand written
    uint64_t* l1_full = bars + 11;    // layer-1 accumulators complete
    uint64_t* l1_read = bars + 12;    // every row has its layer-1 pre-activations in registers: the accumulators may be overwritten
    uint64_t* l2_full = bars + 13;    // layer-2 accumulators complete
    uint64_t* acc_free = bars + 14;   // the epilogue has read the accumulators: the next tile's layer 1 may start
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ntiles = (int)((n + TM - 1) / TM);
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // >= 1: gridDim.x <= ntiles

    if (tid == 0) {
        for (int s = 0; s < NSB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(&a_full[s], TC_COMPUTE / 32); mbar_init(&a_empty[s], 1); }
        mbar_init(x_full, TC_COMPUTE / 32); mbar_init(l1_full, 1); mbar_init(l1_read, TC_COMPUTE / 32);
        mbar_init(l2_full, 1); mbar_init(acc_free, TC_COMPUTE / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_COMPUTE / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < HID; i += TC_THREADS) {
        b1p[i] = packed[OFF_B1 + l1_unit(i)];
        w3p[i] = make_float4(packed[OFF_W3 + i], packed[OFF_W3 + HID + i], packed[OFF_W3 + 2 * HID + i], packed[OFF_B2 + i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == TC_COMPUTE / 32) {
        // ------------------------------------------------------------------ control lane: weight stream + MMA issue
        if (tid == TC_COMPUTE) {
            const int total_q = my_tiles * NCHUNK;                       // weight chunks this CTA consumes
            for (int s = 0; s < NSB; ++s) {
                mbar_expect_tx(&b_full[s], B_STAGE);
                bulk_g2s(sm + OFF_B + s * B_STAGE, image + (size_t)s * B_STAGE, B_STAGE, &b_full[s]);
            }
            int q = 0, sb = 0, bphase = 0;                               // sb = q % NSB, bphase = (q / NSB) & 1
#pragma unroll 1
            for (int t = 0; t < my_tiles; ++t) {
#pragma unroll 1
                for (int c = 0; c < NCHUNK; ++c) {
                    mbar_wait(&b_full[sb], bphase);
                    uint32_t a_h;
                    int sa = 0;
                    if (c == 0) {
                        mbar_wait(x_full, t & 1);
                        if (t > 0) mbar_wait(acc_free, (t - 1) & 1);     // the previous tile's epilogue has read its accumulators
                        a_h = smem_u32(sm + OFF_X);
                    } else {
                        const int qa = t * NCH + c - 1;
                        sa = qa % NSA;
                        mbar_wait(&a_full[sa], (qa / NSA) & 1);
                        if (c == 1) mbar_wait(l1_read, t & 1);           // layer 2 starts over in the same accumulators
                        a_h = smem_u32(sm + OFF_A + sa * A_STAGE);
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_m = a_h + A_WORD, a_l = a_m + A_WORD;
                    const uint32_t b_h = smem_u32(sm + OFF_B + sb * B_STAGE), b_m = b_h + B_WORD, b_l = b_m + B_WORD;
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks) {               // UMMA K = 16 bf16 = 32 bytes along the swizzled row
                        const uint32_t o = ks * 32;
                        const uint32_t acc = (c <= 1 && ks == 0) ? 0u : 1u;
                        // smallest products first into the small accumulator, the exact 16-bit products alone into the main one
                        umma_bf16(tmem_d + HID, umma_desc(a_m + o), umma_desc(b_m + o), acc);
                        umma_bf16(tmem_d + HID, umma_desc(a_h + o), umma_desc(b_l + o), 1u);
                        umma_bf16(tmem_d + HID, umma_desc(a_l + o), umma_desc(b_h + o), 1u);
                        umma_bf16(tmem_d + HID, umma_desc(a_h + o), umma_desc(b_m + o), 1u);
                        umma_bf16(tmem_d + HID, umma_desc(a_m + o), umma_desc(b_h + o), 1u);
                        umma_bf16(tmem_d, umma_desc(a_h + o), umma_desc(b_h + o), acc);
                    }
                    if (c > 0) umma_commit(&a_empty[sa]);                // arrive when these MMAs have read their operands
                    umma_commit(&b_empty[sb]);
                    if (c == 0) umma_commit(l1_full);
                    if (c == NCHUNK - 1) umma_commit(l2_full);
                    // refill the B stage last used by chunk q - 1 with chunk q - 1 + NSB while chunk q computes
                    if (q >= 1 && q - 1 + NSB < total_q) {
                        const int so = (sb + NSB - 1) % NSB;
                        mbar_wait(&b_empty[so], ((q - 1) / NSB) & 1);
                        mbar_expect_tx(&b_full[so], B_STAGE);
                        bulk_g2s(sm + OFF_B + so * B_STAGE, image + (size_t)((q - 1 + NSB) % NCHUNK) * B_STAGE, B_STAGE, &b_full[so]);
                    }
                    ++q;
                    if (++sb == NSB) { sb = 0; bphase ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ rows
        const int r = tid & (TM - 1), part = tid >> 7;
        const uint32_t a_off = sw64((uint32_t)r, (uint32_t)part);       // this thread's 16 bytes of every A word buffer
        const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int CPT = HID / NPART;                                // accumulator columns per thread: 64
        {
            int64_t g = (int64_t)blockIdx.x * TM + r;
            const bool live = g < n;
            produce_x(sm + OFF_X, a_off, part, live ? g : n - 1, live, obs_f32, st, obs_stats, obs_out);
            warp_arrive(x_full);
        }
#pragma unroll 1
        for (int t = 0; t < my_tiles; ++t) {
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TM;
            int64_t g = row0 + r;
            const bool live = g < n;
            if (!live) g = n - 1;
            // ---- layer-1 pre-activations of the thread's 64 hidden units (columns [64 part, +64) of both accumulators)
            float pre1[CPT];
            mbar_wait(l1_full, t & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int cb = 0; cb < CPT / 32; ++cb) {
                float v[32], u[32];
                tmem_ld32(lane_base + (uint32_t)(part * CPT + cb * 32), v);
                tmem_ld32(lane_base + (uint32_t)(HID + part * CPT + cb * 32), u);
#pragma unroll
                for (int j = 0; j < 32; ++j) pre1[cb * 32 + j] = (v[j] + u[j]) + b1p[part * CPT + cb * 32 + j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            warp_arrive(l1_read);
            // ---- chunks 1..8: h1 = act(pre1), split, A operand of 32 hidden units
#pragma unroll
            for (int kc = 0; kc < NCH; ++kc) {
                const int qa = t * NCH + kc, sa = kc % NSA;             // NCH is a multiple of NSA
                float h[UPT];
#pragma unroll
                for (int j = 0; j < UPT; ++j) h[j] = act_fn<TANH>(pre1[kc * UPT + j]);
                uint4 H, M, L;
                split8(h, H, M, L);
                if (qa >= NSA) mbar_wait(&a_empty[sa], ((qa / NSA) - 1) & 1);   // the MMAs of A chunk qa - NSA have consumed this stage
                unsigned char* a0 = sm + OFF_A + sa * A_STAGE + a_off;
                *reinterpret_cast<uint4*>(a0) = H;
                *reinterpret_cast<uint4*>(a0 + A_WORD) = M;
                *reinterpret_cast<uint4*>(a0 + 2 * A_WORD) = L;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                warp_arrive(&a_full[sa]);
            }
            // ---- the next tile's observation operand, while the last chunks of this tile are on the tensor core (the layer-1
            // MMAs that read the X stage completed before l1_full)
            if (t + 1 < my_tiles) {
                int64_t gn = row0 + (int64_t)gridDim.x * TM + r;
                const bool live_n = gn < n;
                produce_x(sm + OFF_X, a_off, part, live_n ? gn : n - 1, live_n, obs_f32, st, obs_stats, obs_out);
                warp_arrive(x_full);
            }

            // ------------------------------------------------------------------ epilogue: h2 = act(D + b2), heads, sample
            mbar_wait(l2_full, t & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float pre[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int cb = 0; cb < CPT / 32; ++cb) {
                const int c0 = part * CPT + cb * 32;
                float v[32], u[32];
                tmem_ld32(lane_base + (uint32_t)c0, v);
                tmem_ld32(lane_base + (uint32_t)(HID + c0), u);
                if (cb == CPT / 32 - 1) {                                 // accumulators are in registers: release them
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    warp_arrive(acc_free);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float4 wb = w3p[c0 + j];                        // (W3[0][col], W3[1][col], W3[2][col], b2[col]): one broadcast load
                    const float h2 = act_fn<TANH>((v[j] + u[j]) + wb.w);
                    pre[0] = fmaf(h2, wb.x, pre[0]); pre[1] = fmaf(h2, wb.y, pre[1]); pre[2] = fmaf(h2, wb.z, pre[2]);
                }
            }
            // reduce the four quarters of every row (all of this tile's MMAs are complete and the next tile's chunks are
            // written by these same threads after the second barrier: the A ring is free to be used as scratch)
            float* red = reinterpret_cast<float*>(sm + OFF_A);          // [NPART][TM][4]
            *reinterpret_cast<float4*>(red + (part * TM + r) * 4) = make_float4(pre[0], pre[1], pre[2], 0.0f);
            asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE) : "memory");   // the 16 compute warps only
            float pre_a = 0.0f;
            if (part < 3)
                pre_a = ((red[(0 * TM + r) * 4 + part] + red[(1 * TM + r) * 4 + part]) + red[(2 * TM + r) * 4 + part]) + red[(3 * TM + r) * 4 + part];
            asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE) : "memory");
            if (live && part < 3) {
                const int a = part;
                float eps;
                if (eps_in) eps = eps_in[g * 3 + a];
                else {
                    const uint64_t gid = (uint64_t)(row_offset + g);
                    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
                    sat::philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
                    // Box-Muller on (0,1) uniforms: (c0, c1) -> eps 0, 1; (c2, c3) -> eps 2 (same draws as actor.cu)
                    const uint32_t ca = (a == 2) ? c[2] : c[0], cb2 = (a == 2) ? c[3] : c[1];
                    const float ua = ((float)ca + 0.5f) * 2.3283064365386963e-10f, ub = ((float)cb2 + 0.5f) * 2.3283064365386963e-10f;
                    const float rr = sqrtf(-2.0f * logf(fminf(ua, 0.99999994f)));
                    float sn, cs;
                    sincosf(6.283185307179586f * ub, &sn, &cs);
                    eps = rr * ((a == 1) ? sn : cs);
                }
                const float mean = max_action * tanhf(pre_a + __ldg(packed + OFF_B3 + a));                        // :87
                const float sd = expf(__ldg(packed + OFF_LS + a));                                               // :93
                float xs = fmaf(sd, eps, mean);                                                                  // :186
                xs = fminf(fmaxf(xs, -max_action), max_action);                                                  // :187
                const float diff = xs - mean;
                const float lp = -(diff * diff) / (2.0f * sd * sd) - logf(sd) - 0.9189385332046727f;             // :188
                act[g * 3 + a] = xs; logp[g * 3 + a] = lp;
                if (mean_out) mean_out[g * 3 + a] = mean;
                if (eps_out) eps_out[g * 3 + a] = eps;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_COMPUTE / 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}

}  // namespace

extern "C" {

int sat_actor_sample_tc(const SatActorWeights* w, float* tc_image, const float* obs_f32, const SatEnvState* st,
                        const double* obs_stats, int64_t n, int64_t row_offset, uint64_t seed, uint64_t step,
                        const float* eps_in, float* act, float* logp, float* mean_out, float* eps_out, float* obs_out,
                        void* stream) {
    if (!w || !w->packed || !tc_image || !act || !logp || (!obs_f32 && !st)) return SAT_ERR_NULL;
    if (w->in_dim != IN || w->hidden != HID || w->act_dim != 3) return SAT_ERR_SIZE;
    if (n <= 0 || ((uintptr_t)w->packed & 15) || ((uintptr_t)tc_image & 15)) return SAT_ERR_SIZE;
    if (!obs_f32 && (st->n < n || st->ld < st->n)) return SAT_ERR_SIZE;
    static unsigned char done[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(actor_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(actor_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) done[dev] = 1;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // the image is rebuilt from the live weights on every call (9216 threads, a few microseconds): it can never be stale
    unsigned char* image = reinterpret_cast<unsigned char*>(tc_image);
    actor_tc_pack_kernel<<<(NCHUNK * 4 * HID + 255) / 256, 256, 0, s>>>(w->packed, image);
    int rc = launch_status();
    if (rc) return rc;
    SatEnvState s0 = {};
    if (!obs_f32) s0 = *st;
    static int sm_count[64] = {0};
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (sms <= 0) {
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) sm_count[dev] = sms;
    }
    const int64_t ntiles = (n + TM - 1) / TM;
    const unsigned blocks = (unsigned)(ntiles < sms ? ntiles : sms);      // persistent: one CTA per SM walks over the row tiles
    if (w->use_tanh)
        actor_tc_kernel<true><<<blocks, TC_THREADS, TC_SMEM, s>>>(w->packed, image, obs_f32, s0, obs_stats, n, row_offset, seed, step,
                                                                  w->max_action, eps_in, act, logp, mean_out, eps_out, obs_out);
    else
        actor_tc_kernel<false><<<blocks, TC_THREADS, TC_SMEM, s>>>(w->packed, image, obs_f32, s0, obs_stats, n, row_offset, seed, step,
                                                                   w->max_action, eps_in, act, logp, mean_out, eps_out, obs_out);
    return launch_status();
}

}  // extern "C"
