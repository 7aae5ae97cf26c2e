// ppo_wgrad2_tc.cu -- dW2 = dz2^T h1 of the fused PPO minibatch step on the tensor cores (sm_100a).
//
// Reference behaviour replaced: the fc2.weight gradient of loss.backward() in PPO_continuous.update (ppo_continuous.py:228-235),
// which the reference accumulates in fp32. Same inputs / output as ppo_wgrad2_kernel (ppo_update.cu): dz2 in column blocks
// [4][mp][64], h1 [mp][256], per-slab partial sums part_w2 [slab][256][256] (fc2.weight layout) that colsum_kernel adds up.
//
// A 256 x 256 output with the minibatch rows as the K dimension. One CTA = (row slab, 128-row half of the output); 74 slabs x
// 2 halves = 148 CTAs. Both operands live in HBM with K (the rows) as the SLOW dimension, so the row warps transpose while they
// stage: thread = (output row / column c, group of 8 minibatch rows) reads 8 floats of column c (coalesced over the warp's 32
// columns), splits them exactly into three bf16 words (tc_mlp.cuh) and writes one 16-byte piece of the K-major, 64-byte-swizzled
// operand row c. Six word products per fp32 product accumulate in TMEM (exact A_h B_h products in their own accumulator) over the
// slab's rows: <= 56 accumulations deep, then the slabs are summed in fp32 as before.
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_mlp.cuh"
#include "ppo_fb_tc.cuh"

namespace {
using namespace mlp;
using namespace tcm;

constexpr int WG_NS = 3;                          // ring depth
constexpr int WA_STAGE = 3 * TM * 64;             // dz2^T chunk: [h | m | l] x 128 output rows x 64 bytes (32 minibatch rows): 24 KB
constexpr int WB_STAGE = 3 * HID * 64;            // h1^T chunk: [h | m | l] x 256 output columns x 64 bytes: 48 KB
constexpr int OFF_WA = 0;
constexpr int OFF_WB = OFF_WA + WG_NS * WA_STAGE;
constexpr int OFF_WBAR = OFF_WB + WG_NS * WB_STAGE;
constexpr int OFF_WTMEM = OFF_WBAR + 16 * 8;
constexpr int WG_SMEM = OFF_WTMEM + 16 + 1024;
static_assert(WG_SMEM <= 227 * 1024, "shared memory budget");
constexpr int WST_LD = 20;                        // epilogue staging row (16 floats + padding), aliases the operand ring
static_assert(16 * 32 * WST_LD * 4 <= WG_NS * WA_STAGE, "epilogue staging must fit in the A ring");

__global__ void __launch_bounds__(TC_THREADS, 1)
ppo_wgrad2_tc_kernel(const float* __restrict__ dz2b, const float* __restrict__ h1g, int64_t mp, int slabs, float* __restrict__ part_w2) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_WBAR);
    uint64_t* full = bars;                        // [WG_NS] both operand chunks written
    uint64_t* empty = bars + WG_NS;               // [WG_NS] the chunk's MMAs have read them
    uint64_t* done = bars + 2 * WG_NS;            // accumulators complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_WTMEM);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int mt = blockIdx.x & 1, slab = blockIdx.x >> 1;        // the two halves of a slab run side by side: h1 is read from L2 once
    const int64_t nchunks = mp / KC;
    const int64_t c0 = nchunks * slab / slabs, c1 = nchunks * (slab + 1) / slabs;
    const int nc = (int)(c1 - c0);

    if (tid == 0) {
        for (int i = 0; i < WG_NS; ++i) { mbar_init(&full[i], TC_COMPUTE / 32); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_COMPUTE / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == TC_COMPUTE / 32) {
        if (tid == TC_COMPUTE) {
            int s = 0, ph = 0;
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_h = smem_u32(sm + OFF_WA + s * WA_STAGE), a_m = a_h + TM * 64, a_l = a_m + TM * 64;
                const uint32_t b_h = smem_u32(sm + OFF_WB + s * WB_STAGE), b_m = b_h + HID * 64, b_l = b_m + HID * 64;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                    const uint32_t o = ks * 32;
                    const uint32_t acc = (c | ks) ? 1u : 0u;
                    umma_bf16(tmem_d + HID, umma_desc_a(a_m + o), umma_desc_a(b_m + o), acc);
                    umma_bf16(tmem_d + HID, umma_desc_a(a_h + o), umma_desc_a(b_l + o), 1u);
                    umma_bf16(tmem_d + HID, umma_desc_a(a_l + o), umma_desc_a(b_h + o), 1u);
                    umma_bf16(tmem_d + HID, umma_desc_a(a_h + o), umma_desc_a(b_m + o), 1u);
                    umma_bf16(tmem_d + HID, umma_desc_a(a_m + o), umma_desc_a(b_h + o), 1u);
                    umma_bf16(tmem_d, umma_desc_a(a_h + o), umma_desc_a(b_h + o), acc);
                }
                umma_commit(&empty[s]);
                if (c == nc - 1) umma_commit(done);
                if (++s == WG_NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp < TC_COMPUTE / 32) {
        const int col = tid & (TM - 1), kg = tid >> 7, lane = tid & 31;       // operand row (output row / column), group of 8 minibatch rows
        const uint32_t off = sw64((uint32_t)col, (uint32_t)kg);
        const int j = mt * TM + col;                                            // output row = column of dz2
        const float* asrc = dz2b + ((int64_t)(j >> 6) * mp) * 64 + (j & 63);
        int s = 0, ph = 0;
#pragma unroll 1
        for (int c = 0; c < nc; ++c) {
            const int64_t row = (c0 + c) * KC + kg * UPT;
            float va[UPT], vb0[UPT], vb1[UPT];
#pragma unroll
            for (int q = 0; q < UPT; ++q) {
                va[q] = asrc[(row + q) * 64];
                vb0[q] = h1g[(row + q) * HID + col];
                vb1[q] = h1g[(row + q) * HID + TM + col];
            }
            uint4 H, M, L;
            if (c >= WG_NS) mbar_wait(&empty[s], ph ^ 1);
            unsigned char* a0 = sm + OFF_WA + s * WA_STAGE + off;
            split8(va, H, M, L);
            *reinterpret_cast<uint4*>(a0) = H; *reinterpret_cast<uint4*>(a0 + TM * 64) = M; *reinterpret_cast<uint4*>(a0 + 2 * TM * 64) = L;
            unsigned char* b0 = sm + OFF_WB + s * WB_STAGE + off;                // rows col and col + 128: the swizzle of row r only uses (r >> 1) & 3
            split8(vb0, H, M, L);
            *reinterpret_cast<uint4*>(b0) = H; *reinterpret_cast<uint4*>(b0 + HID * 64) = M; *reinterpret_cast<uint4*>(b0 + 2 * HID * 64) = L;
            split8(vb1, H, M, L);
            b0 += TM * 64;
            *reinterpret_cast<uint4*>(b0) = H; *reinterpret_cast<uint4*>(b0 + HID * 64) = M; *reinterpret_cast<uint4*>(b0 + 2 * HID * 64) = L;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            warp_arrive(&full[s]);
            if (++s == WG_NS) { s = 0; ph ^= 1; }
        }
        // epilogue: this thread = output row 128 mt + (tid & 127), columns [64 kg, 64 kg + 64) -> part_w2[slab] (fc2.weight layout)
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[64];
        const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        tmem_sum64(lane_base, kg * 64, v);
        float* st = reinterpret_cast<float*>(sm + OFF_WA) + warp * (32 * WST_LD);
        float* gwarp = part_w2 + (int64_t)slab * HID * HID + (int64_t)(mt * TM + col - lane) * HID + kg * 64;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(st + lane * WST_LD + q * 4) = make_float4(v[b * 16 + q * 4], v[b * 16 + q * 4 + 1], v[b * 16 + q * 4 + 2], v[b * 16 + q * 4 + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = i * 32 + lane, rr = idx >> 2, pc = idx & 3;
                *reinterpret_cast<float4*>(gwarp + (int64_t)rr * HID + b * 16 + pc * 4) = *reinterpret_cast<const float4*>(st + rr * WST_LD + pc * 4);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_COMPUTE / 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

int ppo_wgrad2_tc_launch(const float* dz2b, const float* h1g, int64_t mp, int slabs, float* part_w2, cudaStream_t stream) {
    if (mp % TM || slabs < 1 || slabs > mp / KC) return SAT_ERR_SIZE;
    cudaError_t e = cudaFuncSetAttribute(ppo_wgrad2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) return (int)e;
    ppo_wgrad2_tc_kernel<<<2 * slabs, TC_THREADS, WG_SMEM, stream>>>(dz2b, h1g, mp, slabs, part_w2);
    e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}
