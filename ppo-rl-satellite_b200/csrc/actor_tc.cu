// actor_tc.cu -- K3 on the 5th-generation tensor cores: the 256 x 256 hidden layer of Actor_Gaussian as an
// error-compensated 3xTF32 product on tcgen05.mma with the accumulator in TMEM (sm_100a only).
//
// Reference behaviour replaced: the same as actor.cu (PPO_continuous.choose_action -> Actor_Gaussian.forward,
// ppo_continuous.py:83-95, 176-189), which the reference evaluates in fp32.
//
// Why: config 3 "with fused actor sampling" spends 67 % of its step in the two FFMA2 actor launches (2 x 191 us at
// 65 536 rows, 65 % of the nominal fp32 rate: the CUDA-core ceiling). The hidden layer is a 65 536 x 256 x 256
// contraction; plain TF32 would be narrower than the reference's fp32, so every operand is split into two TF32 words
// (x = hi + lo exactly to 22 significant bits) and three products are accumulated in fp32 in TMEM:
//     A B ~= A_lo B_hi + A_hi B_lo + A_hi B_hi            (the dropped A_lo B_lo term is ~2^-22 relative).
// tests/test_gpu_actor_tc.py measures the error of this path and of the FFMA2 path against an fp64 ground truth.
//
// One CTA = 128 observations (= the 128 TMEM lanes), 544 threads:
//   warps 0-15 (thread = row r = tid & 127, quarter p = tid >> 7): rebuild / load the observation, layer 1 (18 -> 256,
//             CUDA-core FFMA) in chunks of 32 hidden units (8 per thread), tanh, TF32 hi/lo split, written as the K-major,
//             128-byte-swizzled A operand of the chunk; later the epilogue: tcgen05.ld of 64 of the row's accumulator
//             columns (a warp may only touch TMEM lanes 32 (w % 4) .. +31, which is exactly its rows), bias, tanh, partial
//             head dot products, reduced over the four quarters through shared memory; thread (r, p < 3) then finishes
//             action p: Philox Gaussian sample, clamp, log-prob;
//   warp 16, one elected lane: streams the pre-split, pre-swizzled W2 image (64 KB per K-chunk: hi | lo) into a two-stage
//             shared-memory ring with cp.async.bulk + mbarrier and issues the 12 tcgen05.mma (M 128, N 256, K 8) of each
//             chunk; tcgen05.commit hands the stage back and finally signals the epilogue.
// The large product A_hi B_hi and the two small cross terms go to SEPARATE TMEM accumulators: the tensor core's fp32
// accumulation truncates, and with all 96 accumulations in one place that bias was the dominant error (measured).
// Shared memory: A 2 x 32 KB, B 2 x 64 KB, W1^T 18 KB, W3 / biases 6 KB = 216 KB -> one CTA per SM; TMEM: 512 columns.
#include <cstdint>
#include <cuda_runtime.h>
#include "sat_math.cuh"
#include "mlp_tile.cuh"

namespace {
using namespace mlp;

constexpr int TM = 128;                         // rows per CTA
constexpr int KC = 32;                          // hidden units (K) per chunk = one 128-byte swizzle row of TF32
constexpr int NCH = HID / KC;                   // 8 chunks
constexpr int NPART = 4;                        // threads per row
constexpr int TC_COMPUTE = TM * NPART;          // 512
constexpr int TC_THREADS = TC_COMPUTE + 32;
constexpr int A_BYTES = TM * 128;               // one chunk of A, hi or lo
constexpr int B_BYTES = HID * 128;              // one chunk of B, hi or lo
constexpr int OFF_A = 0;                        // [stage][hi/lo][A_BYTES]
constexpr int OFF_B = OFF_A + 2 * 2 * A_BYTES;  // [stage][hi/lo][B_BYTES]
constexpr int OFF_W1 = OFF_B + 2 * 2 * B_BYTES; // W1^T [IN][HID] fp32
constexpr int OFF_W3S = OFF_W1 + IN * HID * 4;  // W3 [ACTP][HID]
constexpr int OFF_B1S = OFF_W3S + ACTP * HID * 4;
constexpr int OFF_B2S = OFF_B1S + HID * 4;
constexpr int OFF_BAR = OFF_B2S + HID * 4;      // mbarriers
constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
constexpr int TC_SMEM = OFF_TMEM + 16 + 1024;   // + slack for the 1024-byte alignment of the swizzled tiles
static_assert(TC_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t kTmemCols = 512;             // [0, 256): A_hi B_hi; [256, 512): A_lo B_hi + A_hi B_lo
// tcgen05 instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D fp32 (bit 4), A and B TF32 (2 << 7, 2 << 10),
// both K-major (bits 15, 16 = 0), N = 256 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// shared-memory matrix descriptor of a K-major SWIZZLE_128B operand whose 8-row groups are 1024 bytes apart
// (SmemDescriptor: start >> 4, LBO = 1, SBO = 1024 >> 4 at bit 32, version 1 at bit 46, layout type 2 at bit 61)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
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

// W2 as the B operand: for K-chunk kc, [hi | lo] x 256 rows (n = output unit) x 128 bytes (32 TF32 along k), rows grouped by 8
// (1024 bytes per group) with the 16-byte column chunks XOR-swizzled by the row within the group: exactly the bytes the
// SWIZZLE_128B descriptor expects, so one linear 64 KB bulk copy per chunk brings it in. packed: the FFMA kernel's image
// (W2T[k][n] = fc2.weight[n][k]).
__global__ void actor_tc_pack_kernel(const float* __restrict__ packed, uint32_t* __restrict__ image) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one (n, k)
    if (idx >= HID * HID) return;
    const int k = idx / HID, n = idx % HID;                     // coalesced read of W2T[k][n]
    const float w = packed[OFF_W2T + k * HID + n];
    const uint32_t hi = to_tf32(w);
    const uint32_t lo = to_tf32(w - __uint_as_float(hi));
    const int kc = k / KC, kk = k % KC;
    const int off = (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 2) ^ (n & 7)) << 4)) + (kk & 3) * 4;   // bytes
    uint32_t* chunk = image + (size_t)kc * (2 * B_BYTES / 4);
    chunk[off >> 2] = hi;
    chunk[(B_BYTES + off) >> 2] = lo;
}

template <bool TANH>
__global__ void __launch_bounds__(TC_THREADS, 1)
actor_tc_kernel(const float* __restrict__ packed, const uint32_t* __restrict__ image, const float* __restrict__ obs_f32,
                const SatEnvState st, const double* __restrict__ obs_stats, int64_t n, int64_t row_offset, uint64_t seed,
                uint64_t step, float max_action, const float* __restrict__ eps_in, float* __restrict__ act,
                float* __restrict__ logp, float* __restrict__ mean_out, float* __restrict__ eps_out, float* __restrict__ obs_out) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment for the swizzled tiles by pointer arithmetic on the shared array (an integer round trip would turn
    // every access into a generic-space LD/ST: measured 14 % of the kernel's instructions)
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    float* w1s = reinterpret_cast<float*>(sm + OFF_W1);
    float* w3s = reinterpret_cast<float*>(sm + OFF_W3S);
    float* b1s = reinterpret_cast<float*>(sm + OFF_B1S);
    float* b2s = reinterpret_cast<float*>(sm + OFF_B2S);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* b_full = bars;          // [2] bulk copy of a W2 chunk landed
    uint64_t* a_full = bars + 2;      // [2] 128 rows of the A chunk written
    uint64_t* ab_empty = bars + 4;    // [2] the chunk's MMAs are complete: both operand stages reusable
    uint64_t* d_full = bars + 6;      // accumulator complete
    uint64_t* misc = bars + 7;        // W1^T / W3 landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * TM;

    if (tid == 0) {
        mbar_init(&b_full[0], 1); mbar_init(&b_full[1], 1);
        mbar_init(&a_full[0], TC_COMPUTE); mbar_init(&a_full[1], TC_COMPUTE);
        mbar_init(&ab_empty[0], 1); mbar_init(&ab_empty[1], 1);
        mbar_init(d_full, 1); mbar_init(misc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_COMPUTE / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < HID; i += TC_THREADS) { b1s[i] = packed[OFF_B1 + i]; b2s[i] = packed[OFF_B2 + i]; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == TC_COMPUTE / 32) {
        // ------------------------------------------------------------------ control lane: W2 stream + MMA issue
        if (tid == TC_COMPUTE) {
            mbar_expect_tx(misc, IN * HID * 4 + ACTP * HID * 4);
            bulk_g2s(w1s, packed + OFF_W1T, IN * HID * 4, misc);
            bulk_g2s(w3s, packed + OFF_W3, ACTP * HID * 4, misc);
            for (int s = 0; s < 2; ++s) {
                mbar_expect_tx(&b_full[s], 2 * B_BYTES);
                bulk_g2s(sm + OFF_B + s * 2 * B_BYTES, image + (size_t)s * (2 * B_BYTES / 4), 2 * B_BYTES, &b_full[s]);
            }
            for (int kc = 0; kc < NCH; ++kc) {
                const int s = kc & 1, ph = (kc >> 1) & 1;
                mbar_wait(&b_full[s], ph);
                mbar_wait(&a_full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(sm + OFF_A + s * 2 * A_BYTES), a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = smem_u32(sm + OFF_B + s * 2 * B_BYTES), b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int kk = 0; kk < KC / 8; ++kk) {                    // UMMA K = 8 TF32 = 32 bytes along the swizzled row
                    const uint32_t o = kk * 32;
                    umma_tf32(tmem_d + HID, umma_desc(a_lo + o), umma_desc(b_hi + o), (kc | kk) ? 1u : 0u);   // cross terms
                    umma_tf32(tmem_d + HID, umma_desc(a_hi + o), umma_desc(b_lo + o), 1u);
                    umma_tf32(tmem_d, umma_desc(a_hi + o), umma_desc(b_hi + o), (kc | kk) ? 1u : 0u);         // main term
                }
                umma_commit(&ab_empty[s]);                               // arrives when these MMAs have read their operands
                if (kc == NCH - 1) umma_commit(d_full);
                // refill the OTHER stage (last used by chunk kc - 1) with chunk kc + 1 while chunk kc computes
                if (kc >= 1 && kc + 1 < NCH) {
                    const int so = (kc + 1) & 1;
                    mbar_wait(&ab_empty[so], ((kc - 1) >> 1) & 1);
                    mbar_expect_tx(&b_full[so], 2 * B_BYTES);
                    bulk_g2s(sm + OFF_B + so * 2 * B_BYTES, image + (size_t)(kc + 1) * (2 * B_BYTES / 4), 2 * B_BYTES, &b_full[so]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ rows: observation, layer 1, A operand
        const int r = tid & (TM - 1), part = tid >> 7;
        int64_t g = row0 + r;
        const bool live = g < n;
        if (!live) g = n - 1;
        float x[IN];
        if (obs_f32) {
#pragma unroll
            for (int d = 0; d < IN; ++d) x[d] = obs_f32[g * IN + d];
        } else {
            // rebuild the observation from the fp64 SoA env state (environment.py:76-77), normalise in fp64 (normalization.py:41)
            const int64_t ld = st.ld;
            double o[IN];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double P = st.state[(SAT_COL_P + k) * ld + g], Pv = st.state[(SAT_COL_PV + k) * ld + g];
                const double E = st.state[(SAT_COL_E + k) * ld + g], Ev = st.state[(SAT_COL_EV + k) * ld + g];
                o[k] = P - E; o[3 + k] = Pv - Ev; o[6 + k] = P; o[9 + k] = Pv; o[12 + k] = E; o[15 + k] = Ev;
            }
#pragma unroll
            for (int d = 0; d < IN; ++d) {
                double y = o[d];
                if (obs_stats) y = (y - obs_stats[1 + d]) / (obs_stats[1 + 2 * IN + d] + 1e-8);
                x[d] = (float)y;
            }
        }
        if (obs_out && live && part == 0) {
#pragma unroll
            for (int d = 0; d < IN; ++d) obs_out[g * IN + d] = x[d];
        }
        mbar_wait(misc, 0);
        const int r8 = r & 7;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)r8 * 128u;
        constexpr int UPT = KC / NPART;                 // hidden units per thread per chunk: 8 = two 16-byte swizzle chunks
#pragma unroll 1
        for (int kc = 0; kc < NCH; ++kc) {
            const int s = kc & 1;
            const int j0 = kc * KC + part * UPT;
            float h[UPT];
#pragma unroll
            for (int j = 0; j < UPT; ++j) h[j] = b1s[j0 + j];
#pragma unroll
            for (int k = 0; k < IN; ++k) {
                const float xv = x[k];
                const float4* wrow = reinterpret_cast<const float4*>(w1s + k * HID + j0);
#pragma unroll
                for (int q = 0; q < UPT / 4; ++q) {
                    const float4 w = wrow[q];
                    h[4 * q] = fmaf(xv, w.x, h[4 * q]); h[4 * q + 1] = fmaf(xv, w.y, h[4 * q + 1]);
                    h[4 * q + 2] = fmaf(xv, w.z, h[4 * q + 2]); h[4 * q + 3] = fmaf(xv, w.w, h[4 * q + 3]);
                }
            }
            if (kc >= 2) mbar_wait(&ab_empty[s], ((kc >> 1) - 1) & 1);   // the MMAs of chunk kc - 2 have consumed this stage
            unsigned char* a_hi = sm + OFF_A + s * 2 * A_BYTES + row_off;
#pragma unroll
            for (int q = 0; q < UPT / 4; ++q) {
                const int c = part * (UPT / 4) + q;                      // 16-byte chunk of the row's 128 bytes
                uint4 hi, lo;
                float v;
                v = act_fn<TANH>(h[4 * q]);     hi.x = to_tf32(v); lo.x = to_tf32(v - __uint_as_float(hi.x));
                v = act_fn<TANH>(h[4 * q + 1]); hi.y = to_tf32(v); lo.y = to_tf32(v - __uint_as_float(hi.y));
                v = act_fn<TANH>(h[4 * q + 2]); hi.z = to_tf32(v); lo.z = to_tf32(v - __uint_as_float(hi.z));
                v = act_fn<TANH>(h[4 * q + 3]); hi.w = to_tf32(v); lo.w = to_tf32(v - __uint_as_float(hi.w));
                const uint32_t sw = (uint32_t)((c ^ r8) << 4);
                *reinterpret_cast<uint4*>(a_hi + sw) = hi;
                *reinterpret_cast<uint4*>(a_hi + A_BYTES + sw) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
            mbar_arrive(&a_full[s]);
        }

        // ------------------------------------------------------------------ epilogue: h2 = act(D + b2), heads, sample
        // head weights and the second bias interleaved per hidden unit, so the epilogue needs one broadcast load per column.
        // The table reuses the W1^T region: every compute thread is past layer 1 at the barrier (the operand rings cannot be
        // reused yet - the last chunk's MMAs may still be reading them)
        asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE) : "memory");
        float4* w3p = reinterpret_cast<float4*>(w1s);
        if (tid < HID) w3p[tid] = make_float4(w3s[tid], w3s[HID + tid], w3s[2 * HID + tid], b2s[tid]);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE) : "memory");
        mbar_wait(d_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float pre[3] = {0.0f, 0.0f, 0.0f};
        const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int CPT = HID / NPART;                // accumulator columns per thread: 64
#pragma unroll 1
        for (int cb = 0; cb < CPT / 32; ++cb) {
            const int c0 = part * CPT + cb * 32;
            float v[32], u[32];
            tmem_ld32(lane_base + (uint32_t)c0, v);
            tmem_ld32(lane_base + (uint32_t)(HID + c0), u);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c0 + j;
                const float4 wb = w3p[col];                               // (W3[0][col], W3[1][col], W3[2][col], b2[col]): one broadcast load
                const float h2 = act_fn<TANH>((v[j] + u[j]) + wb.w);
                pre[0] = fmaf(h2, wb.x, pre[0]); pre[1] = fmaf(h2, wb.y, pre[1]); pre[2] = fmaf(h2, wb.z, pre[2]);
            }
        }
        // reduce the four quarters of every row (all MMAs are complete: the A ring is free to be reused as scratch)
        float* red = reinterpret_cast<float*>(sm + OFF_A);              // [NPART][TM][4]
        *reinterpret_cast<float4*>(red + (part * TM + r) * 4) = make_float4(pre[0], pre[1], pre[2], 0.0f);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_COMPUTE) : "memory");   // the 16 compute warps only
        if (live && part < 3) {
            const int a = part;
            const float pre_a = ((red[(0 * TM + r) * 4 + a] + red[(1 * TM + r) * 4 + a]) + red[(2 * TM + r) * 4 + a]) + red[(3 * TM + r) * 4 + a];
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
    // the image is rebuilt from the live weights on every call (65 536 elements, a few microseconds): it can never be stale
    actor_tc_pack_kernel<<<HID * HID / 256, 256, 0, s>>>(w->packed, reinterpret_cast<uint32_t*>(tc_image));
    int rc = launch_status();
    if (rc) return rc;
    SatEnvState s0 = {};
    if (!obs_f32) s0 = *st;
    const unsigned blocks = (unsigned)((n + TM - 1) / TM);
    if (w->use_tanh)
        actor_tc_kernel<true><<<blocks, TC_THREADS, TC_SMEM, s>>>(w->packed, reinterpret_cast<const uint32_t*>(tc_image), obs_f32, s0,
                                                                  obs_stats, n, row_offset, seed, step, w->max_action, eps_in, act,
                                                                  logp, mean_out, eps_out, obs_out);
    else
        actor_tc_kernel<false><<<blocks, TC_THREADS, TC_SMEM, s>>>(w->packed, reinterpret_cast<const uint32_t*>(tc_image), obs_f32, s0,
                                                                   obs_stats, n, row_offset, seed, step, w->max_action, eps_in, act,
                                                                   logp, mean_out, eps_out, obs_out);
    return launch_status();
}

}  // extern "C"
