// gae.cu -- K4: GAE reverse scan + advantage normalisation for sm_100a.
//
// Reference behaviour replaced: ppo_continuous.py:198-210.
//   deltas = r + gamma*(1-dw)*vs_ - vs                    (torch fp32, separate ops)
//   gae    = delta + gamma*lamda*gae*(1-d)                (numpy float32 scalars under numpy 2)
//   v_target = adv + vs;  adv = (adv - mean)/(std_unbiased + 1e-5)
// The recursion is reproduced with explicit round-to-nearest fp32 ops in the reference's order,
// so advantages are bit-identical to the reference (bar: 1e-6).
//
// Layouts: time-major [T][N] (one thread per env walks t = T-1..0, every access coalesced, v[t+1]
// carried in a register), and the reference's flat (B,1) buffer of a single env, which is staged
// through shared memory in chunks and scanned there.
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/satb200.h"

namespace {

__device__ __forceinline__ float gae_delta(float r, float gamma, float dw, float vnext, float v) {
    // (r + ((gamma * (1 - dw)) * vs_)) - vs
    return __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, __fsub_rn(1.0f, dw)), vnext)), v);
}
__device__ __forceinline__ float gae_step(float delta, float gl, float gae, float d) {
    // delta + ((gl * gae) * (1 - d))
    return __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, gae), __fsub_rn(1.0f, d)));
}

__global__ void __launch_bounds__(128)
gae_time_major_kernel(const float* __restrict__ r, const float* __restrict__ v, const uint8_t* __restrict__ done,
                      const float* __restrict__ r_scale, int64_t T, int64_t N, float gamma, float gl,
                      float* __restrict__ adv, float* __restrict__ vt) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float gae = 0.0f;
    float vnext = v[T * N + n];
    for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = t * N + n;
        const float d = done[i] ? 1.0f : 0.0f;
        float rr = r[i];
        if (r_scale) rr = __fmul_rn(rr, r_scale[t]);
        const float vs = v[i];
        const float delta = gae_delta(rr, gamma, d, vnext, vs);
        gae = gae_step(delta, gl, gae, d);
        adv[i] = gae;
        vt[i] = __fadd_rn(gae, vs);                          // :208
        vnext = vs;
    }
}

constexpr int kFlatChunk = 4096;

__global__ void __launch_bounds__(256)
gae_flat_kernel(const float* __restrict__ r, const float* __restrict__ vs, const float* __restrict__ vs_next,
                const float* __restrict__ dw, const float* __restrict__ done, int64_t B, float gamma, float gl,
                float* __restrict__ adv, float* __restrict__ vt) {
    // single CTA: chunks of the buffer are staged into shared memory (delta, 1-d), scanned in reverse by
    // one thread (the recurrence is strictly sequential and must round like the reference), then the
    // chunk's advantages are written back cooperatively.
    __shared__ float s_delta[kFlatChunk];
    __shared__ float s_keep[kFlatChunk];
    __shared__ float s_carry;
    if (threadIdx.x == 0) s_carry = 0.0f;
    __syncthreads();
    for (int64_t hi = B; hi > 0; hi -= kFlatChunk) {
        const int64_t lo = hi > kFlatChunk ? hi - kFlatChunk : 0;
        const int len = (int)(hi - lo);
        for (int j = threadIdx.x; j < len; j += blockDim.x) {
            const int64_t i = lo + j;
            s_delta[j] = gae_delta(r[i], gamma, dw[i], vs_next[i], vs[i]);
            s_keep[j] = done[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float gae = s_carry;
            for (int j = len - 1; j >= 0; --j) {
                gae = gae_step(s_delta[j], gl, gae, s_keep[j]);
                s_delta[j] = gae;
            }
            s_carry = gae;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < len; j += blockDim.x) {
            const int64_t i = lo + j;
            const float g = s_delta[j];
            adv[i] = g;
            vt[i] = __fadd_rn(g, vs[i]);
        }
        __syncthreads();
    }
}

constexpr int kMomThreads = 256;

__global__ void __launch_bounds__(kMomThreads)
moments_partial_kernel(const float* __restrict__ x, int64_t count, double* __restrict__ partial) {
    __shared__ double s_sum[kMomThreads / 32], s_sq[kMomThreads / 32];
    double sum = 0.0, sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kMomThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kMomThreads) {
        const double v = (double)x[i];
        sum += v; sq += v * v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        sq += __shfl_xor_sync(0xffffffffu, sq, off);
    }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_sq[threadIdx.x >> 5] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kMomThreads / 32; ++w) { a += s_sum[w]; b += s_sq[w]; }
        partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
    }
}

__global__ void moments_final_kernel(const double* __restrict__ partial, int nblocks, int64_t count, double* __restrict__ sums) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < nblocks; ++i) { a += partial[2 * i]; b += partial[2 * i + 1]; }   // fixed order
        sums[0] = a; sums[1] = b; sums[2] = (double)count;
    }
}

__global__ void adv_normalize_kernel(float* __restrict__ adv, int64_t count, const double* __restrict__ sums) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double n = sums[2];
    const double mean = sums[0] / n;
    double var = (sums[1] - n * mean * mean) / (n - 1.0);      // torch.std(): unbiased
    if (var < 0.0) var = 0.0;
    const float m = (float)mean, sd = (float)sqrt(var);
    adv[i] = __fdiv_rn(__fsub_rn(adv[i], m), __fadd_rn(sd, 1e-5f));   // ppo_continuous.py:210
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}
constexpr int kMomBlocks = 296;   // 2 per SM

}  // namespace

extern "C" {

int sat_gae(const float* r, const float* v, const uint8_t* done, const float* r_scale, int64_t T, int64_t N,
            float gamma, float lamda, float* adv, float* v_target, void* stream) {
    if (!r || !v || !done || !adv || !v_target) return SAT_ERR_NULL;
    if (T <= 0 || N <= 0) return SAT_ERR_SIZE;
    const float gl = (float)((double)gamma * (double)lamda);   // python float product, weak-cast to fp32
    const unsigned blocks = (unsigned)((N + 127) / 128);
    gae_time_major_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(r, v, done, r_scale, T, N, gamma, gl, adv, v_target);
    return launch_status();
}

int sat_gae_flat(const float* r, const float* vs, const float* vs_next, const float* dw, const float* done,
                 int64_t B, float gamma, float lamda, float* adv, float* v_target, void* stream) {
    if (!r || !vs || !vs_next || !dw || !done || !adv || !v_target) return SAT_ERR_NULL;
    if (B <= 0) return SAT_ERR_SIZE;
    const float gl = (float)((double)gamma * (double)lamda);
    gae_flat_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(r, vs, vs_next, dw, done, B, gamma, gl, adv, v_target);
    return launch_status();
}

int sat_adv_moments(const float* adv, int64_t count, double* sums, void* workspace, void* stream) {
    if (!adv || !sums || !workspace) return SAT_ERR_NULL;
    if (count <= 0) return SAT_ERR_SIZE;
    cudaStream_t s = (cudaStream_t)stream;
    int nblocks = (int)((count + kMomThreads - 1) / kMomThreads);
    if (nblocks > kMomBlocks) nblocks = kMomBlocks;
    moments_partial_kernel<<<nblocks, kMomThreads, 0, s>>>(adv, count, (double*)workspace);
    int rc = launch_status();
    if (rc) return rc;
    moments_final_kernel<<<1, 32, 0, s>>>((const double*)workspace, nblocks, count, sums);
    return launch_status();
}

int sat_adv_normalize(float* adv, int64_t count, const double* sums, void* stream) {
    if (!adv || !sums) return SAT_ERR_NULL;
    if (count <= 0) return SAT_ERR_SIZE;
    adv_normalize_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(adv, count, sums);
    return launch_status();
}

}  // extern "C"
