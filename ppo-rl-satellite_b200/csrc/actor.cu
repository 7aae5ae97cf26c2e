// actor.cu -- K3: fused Gaussian-actor forward + Philox sampling (and the critic forward) for sm_100a.
//
// Reference behaviour replaced: PPO_continuous.choose_action -> Actor_Gaussian.get_dist/forward
// (ppo_continuous.py:83-95, 176-189) and Critic.forward (:123-128), evaluated for n observations:
//     h1 = tanh(W1 s + b1); h2 = tanh(W2 h1 + b2); mean = max_action * tanh(W3 h2 + b3)
//     a  = clamp(mean + exp(log_std) * eps, -max_action, max_action); logp = Normal(mean, std).log_prob(a)
//
// fp32 CUDA-core FFMA by design (the reference computes in fp32 and the north star rules tensor cores
// out for this path). One CTA = 64 observations, each thread an 8 x CT register tile (mlp_tile.cuh).
// h1 stays in shared memory (transposed, [256][64+4]); W2^T (256 KB, larger than smem) is streamed
// through two 16 KB stages with TMA bulk copies (cp.async.bulk + mbarrier) from the pre-packed,
// L2-resident weight buffer. 2 CTAs per SM (112 KB smem each, 256 threads with the 8 x 8 tile); 65 536 observations
// -> 1024 CTAs = 6.9 per SM.
#include <cstdint>
#include <cuda_runtime.h>
#include "sat_math.cuh"
#include "mlp_tile.cuh"

namespace {
using namespace mlp;
using Smem = mlp::ActorSmem;

// one network's side of a launch: grid.y selects the job, so the pursuer's and the evader's actor (same observations,
// different weights and Philox step) run as ONE launch - at config-5 shard sizes (8192 envs = 128 CTAs per network) that
// puts both networks on the GPU at the same time instead of one half-empty launch after the other
struct ActorJob {
    const float* packed; const float* eps_in;
    float *act, *logp, *mean_out, *eps_out, *obs_out, *v_out;
    uint64_t step;
    float max_action; int use_tanh;
};
struct ActorJobs { ActorJob j[2]; };

template <bool CRITIC, bool PAIR>
__global__ void __launch_bounds__(THREADS, 2)
actor_kernel(const __grid_constant__ ActorJobs jobs, const float* __restrict__ obs_f32, const SatEnvState st,
             const double* __restrict__ obs_stats, int64_t n, int64_t row_offset, uint64_t seed) {
    // field-by-field select (constant-bank reads): indexing the parameter array with blockIdx.y would copy it to local memory
    const bool second = PAIR && blockIdx.y != 0;
#define SAT_JOB(f) (PAIR ? (second ? jobs.j[1].f : jobs.j[0].f) : jobs.j[0].f)
    const float* __restrict__ packed = SAT_JOB(packed);
    const int use_tanh = SAT_JOB(use_tanh);
    const float max_action = SAT_JOB(max_action);
    const uint64_t step = SAT_JOB(step);
    const float* __restrict__ eps_in = SAT_JOB(eps_in);
    float* __restrict__ act = SAT_JOB(act); float* __restrict__ logp = SAT_JOB(logp); float* __restrict__ mean_out = SAT_JOB(mean_out);
    float* __restrict__ eps_out = SAT_JOB(eps_out); float* __restrict__ obs_out = SAT_JOB(obs_out); float* __restrict__ v_out = SAT_JOB(v_out);
#undef SAT_JOB
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
    const int64_t row0 = (int64_t)blockIdx.x * M;
    const int heads = CRITIC ? 1 : 3;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1); mbar_init(&sm.bar_misc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        // W1^T (18 KB) into the stage region, W3 (4 KB) into its own buffer
        mbar_expect_tx(&sm.bar_misc, IN * HID * 4 + ACTP * HID * 4);
        bulk_g2s(&sm.wt[0][0], packed + OFF_W1T, IN * HID * 4, &sm.bar_misc);
        bulk_g2s(&sm.w3[0], packed + OFF_W3, ACTP * HID * 4, &sm.bar_misc);
    }

    // ---------------- observations -> xT (fp32, transposed)
    if (obs_f32) {
        for (int idx = tid; idx < M * IN; idx += THREADS) {
            const int row = idx / IN, d = idx - row * IN;
            int64_t g = row0 + row; if (g >= n) g = n - 1;
            sm.xT[d * M + row] = obs_f32[g * IN + d];
        }
    } else if (tid < M) {
        // fused path: rebuild the observation from the fp64 SoA env state (environment.py:76-77) and
        // normalise in fp64 (normalization.py:41) before the cast to fp32
        int64_t g = row0 + tid; if (g >= n) g = n - 1;
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
            sm.xT[d * M + tid] = (float)y;
        }
    }
    __syncthreads();
    if (obs_out) {
        for (int idx = tid; idx < M * IN; idx += THREADS) {
            const int row = idx / IN, d = idx - row * IN;
            if (row0 + row < n) obs_out[(row0 + row) * IN + d] = sm.xT[d * M + row];
        }
    }

    // ---------------- layer 1: h1 = act(W1 x + b1) -> h1T
    float2 acc[RT][NP];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[i][j] = make_float2(0.0f, 0.0f);
    mbar_wait(&sm.bar_misc, 0);
    tile_fma<IN>(acc, sm.xT, M, &sm.wt[0][0], ty, tx);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = c * CSTR + tx * 4 + q;
            const float bias = __ldg(packed + OFF_B1 + j);
#pragma unroll
            for (int i4 = 0; i4 < RT / 4; ++i4) {
                float4 o;
                o.x = activate(acc_at(acc, i4 * 4 + 0, c, q) + bias, use_tanh); o.y = activate(acc_at(acc, i4 * 4 + 1, c, q) + bias, use_tanh);
                o.z = activate(acc_at(acc, i4 * 4 + 2, c, q) + bias, use_tanh); o.w = activate(acc_at(acc, i4 * 4 + 3, c, q) + bias, use_tanh);
                *reinterpret_cast<float4*>(&sm.h1T[j * H1_LD + ty * RT + i4 * 4]) = o;
            }
        }
    __syncthreads();      // h1T complete; W1^T region free for the W2^T stages

    // ---------------- layer 2: stream W2^T k-tiles through the two stages
    constexpr int NT = HID / KT;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_expect_tx(&sm.full[s], KT * HID * 4);
            bulk_g2s(&sm.wt[s][0], packed + OFF_W2T + s * KT * HID, KT * HID * 4, &sm.full[s]);
        }
    }
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[i][j] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int t = 0; t < NT; ++t) {
        const int s = t & 1;
        mbar_wait(&sm.full[s], (t >> 1) & 1);
        tile_fma<KT>(acc, sm.h1T + t * KT * H1_LD, H1_LD, &sm.wt[s][0], ty, tx);
        __syncthreads();                                   // every thread is done reading stage s
        if (tid == 0 && t + NSTAGE < NT) {
            mbar_expect_tx(&sm.full[s], KT * HID * 4);
            bulk_g2s(&sm.wt[s][0], packed + OFF_W2T + (t + NSTAGE) * KT * HID, KT * HID * 4, &sm.full[s]);
        }
    }

    // ---------------- layer 3 (heads): partial dot products over this thread's 16 hidden units
    float part[RT][3];
#pragma unroll
    for (int i = 0; i < RT; ++i) { part[i][0] = 0.0f; part[i][1] = 0.0f; part[i][2] = 0.0f; }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = c * CSTR + tx * 4 + q;
            const float bias = __ldg(packed + OFF_B2 + j);
            float w[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) w[a] = (a < heads) ? sm.w3[a * HID + j] : 0.0f;
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const float h = activate(acc_at(acc, i, c, q) + bias, use_tanh);
#pragma unroll
                for (int a = 0; a < 3; ++a) if (a < heads) part[i][a] = fmaf(h, w[a], part[i][a]);
            }
        }
#pragma unroll
    for (int off = 1; off < TXN; off <<= 1)
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int a = 0; a < 3; ++a) if (a < heads) part[i][a] += __shfl_xor_sync(0xffffffffu, part[i][a], off);
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int a = 0; a < 3; ++a) if (a < heads) sm.pre[(ty * RT + i) * ACTP + a] = part[i][a];
    }
    __syncthreads();

    // ---------------- heads: value, or mean + Gaussian sample + log-prob
    if (tid < M) {
        const int64_t g = row0 + tid;
        if (g < n) {
            if (CRITIC) {
                v_out[g] = sm.pre[tid * ACTP] + __ldg(packed + OFF_B3);                       // :127
            } else {
                float eps[4];
                if (eps_in) { eps[0] = eps_in[g * 3]; eps[1] = eps_in[g * 3 + 1]; eps[2] = eps_in[g * 3 + 2]; }
                else {
                    const uint64_t gid = (uint64_t)(row_offset + g);
                    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
                    sat::philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
                    // Box-Muller on (0,1) uniforms
                    const float u0 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
                    const float u2 = ((float)c[2] + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)c[3] + 0.5f) * 2.3283064365386963e-10f;
                    const float r0 = sqrtf(-2.0f * logf(fminf(u0, 0.99999994f))), r1 = sqrtf(-2.0f * logf(fminf(u2, 0.99999994f)));
                    float s0, c0, s1, c1;
                    sincosf(6.283185307179586f * u1, &s0, &c0);
                    sincosf(6.283185307179586f * u3, &s1, &c1);
                    eps[0] = r0 * c0; eps[1] = r0 * s0; eps[2] = r1 * c1; eps[3] = r1 * s1;
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float mean = max_action * tanhf(sm.pre[tid * ACTP + a] + __ldg(packed + OFF_B3 + a));   // :87
                    const float ls = __ldg(packed + OFF_LS + a);
                    const float sd = expf(ls);                                                                   // :93
                    float x = fmaf(sd, eps[a], mean);                                                            // :186
                    x = fminf(fmaxf(x, -max_action), max_action);                                                // :187
                    const float diff = x - mean;
                    const float lp = -(diff * diff) / (2.0f * sd * sd) - logf(sd) - 0.9189385332046727f;         // :188
                    act[g * 3 + a] = x; logp[g * 3 + a] = lp;
                    if (mean_out) mean_out[g * 3 + a] = mean;
                    if (eps_out) eps_out[g * 3 + a] = eps[a];
                }
            }
        }
    }
}

// torch-layout weights -> packed buffer (one launch; run once per weight update)
__global__ void pack_kernel(const SatActorWeights w, float* __restrict__ packed, int heads) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= PACKED_FLOATS) return;
    float val = 0.0f;
    if (idx < OFF_B1) { const int k = idx / HID, j = idx % HID; val = w.w1[j * IN + k]; }
    else if (idx < OFF_W2T) val = w.b1[idx - OFF_B1];
    else if (idx < OFF_B2) { const int r = idx - OFF_W2T; const int k = r / HID, j = r % HID; val = w.w2[j * HID + k]; }
    else if (idx < OFF_W3) val = w.b2[idx - OFF_B2];
    else if (idx < OFF_B3) { const int r = idx - OFF_W3; const int a = r / HID; val = a < heads ? w.w3[r] : 0.0f; }
    else if (idx < OFF_LS) { const int a = idx - OFF_B3; val = a < heads ? w.b3[a] : 0.0f; }
    else { const int a = idx - OFF_LS; val = (w.log_std && a < heads) ? w.log_std[a] : 0.0f; }
    packed[idx] = val;
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}

template <bool CRITIC, bool PAIR = false>
int configure() {
    // the attribute is per device (per context): a process that samples on cuda:0 and then on cuda:1 needs it on both.
    // One flag per device ordinal; benign race (idempotent attribute set).
    static unsigned char done[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(actor_kernel<CRITIC, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) done[dev] = 1;
    }
    return SAT_OK;
}

}  // namespace

extern "C" {

int sat_actor_pack(const SatActorWeights* w, float* packed, void* stream) {
    if (!w || !packed || !w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->w3 || !w->b3) return SAT_ERR_NULL;
    if (w->in_dim != IN || w->hidden != HID || w->act_dim < 1 || w->act_dim > 3) return SAT_ERR_SIZE;
    if ((uintptr_t)packed & 15) return SAT_ERR_SIZE;
    pack_kernel<<<(PACKED_FLOATS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*w, packed, w->act_dim);
    return launch_status();
}

int sat_actor_sample(const SatActorWeights* w, const float* obs_f32, const SatEnvState* st,
                     const double* obs_stats, int64_t n, int64_t row_offset, uint64_t seed, uint64_t step,
                     const float* eps_in, float* act, float* logp, float* mean_out, float* eps_out,
                     float* obs_out, void* stream) {
    if (!w || !w->packed || !act || !logp) return SAT_ERR_NULL;
    if (!obs_f32 && !(st && st->state)) return SAT_ERR_NULL;
    if (w->in_dim != IN || w->hidden != HID || w->act_dim != 3) return SAT_ERR_SIZE;
    if (n <= 0 || ((uintptr_t)w->packed & 15)) return SAT_ERR_SIZE;
    if (!obs_f32 && (st->n < n || st->ld < st->n)) return SAT_ERR_SIZE;
    int rc = configure<false>();
    if (rc) return rc;
    SatEnvState s0 = {};
    if (!obs_f32) s0 = *st;
    const unsigned blocks = (unsigned)((n + M - 1) / M);
    ActorJobs jobs = {};
    jobs.j[0] = ActorJob{w->packed, eps_in, act, logp, mean_out, eps_out, obs_out, nullptr, step, w->max_action, w->use_tanh};
    actor_kernel<false, false><<<dim3(blocks, 1), THREADS, sizeof(Smem), (cudaStream_t)stream>>>(jobs, obs_f32, s0, obs_stats, n, row_offset, seed);
    return launch_status();
}

int sat_critic_forward(const SatActorWeights* w, const float* obs_f32, int64_t n, float* v, void* stream) {
    if (!w || !w->packed || !obs_f32 || !v) return SAT_ERR_NULL;
    if (w->in_dim != IN || w->hidden != HID || w->act_dim != 1) return SAT_ERR_SIZE;
    if (n <= 0 || ((uintptr_t)w->packed & 15)) return SAT_ERR_SIZE;
    int rc = configure<true>();
    if (rc) return rc;
    SatEnvState s0 = {};
    const unsigned blocks = (unsigned)((n + M - 1) / M);
    ActorJobs jobs = {};
    jobs.j[0] = ActorJob{w->packed, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, v, 0, 0.0f, w->use_tanh};
    actor_kernel<true, false><<<dim3(blocks, 1), THREADS, sizeof(Smem), (cudaStream_t)stream>>>(jobs, obs_f32, s0, nullptr, n, 0, 0);
    return launch_status();
}

int sat_actor_sample_pair(const SatActorWeights* wa, const SatActorWeights* wb, const float* obs_f32, const SatEnvState* st,
                          const double* obs_stats, int64_t n, int64_t row_offset, uint64_t seed, uint64_t step_a,
                          uint64_t step_b, float* act_a, float* logp_a, float* obs_out, float* act_b, float* logp_b,
                          void* stream) {
    if (!wa || !wb || !wa->packed || !wb->packed || !act_a || !logp_a || !act_b || !logp_b) return SAT_ERR_NULL;
    if (!obs_f32 && !(st && st->state)) return SAT_ERR_NULL;
    if (wa->in_dim != IN || wa->hidden != HID || wa->act_dim != 3 || wb->in_dim != IN || wb->hidden != HID || wb->act_dim != 3)
        return SAT_ERR_SIZE;
    if (n <= 0 || (((uintptr_t)wa->packed | (uintptr_t)wb->packed) & 15)) return SAT_ERR_SIZE;
    if (!obs_f32 && (st->n < n || st->ld < st->n)) return SAT_ERR_SIZE;
    int rc = configure<false, true>();
    if (rc) return rc;
    SatEnvState s0 = {};
    if (!obs_f32) s0 = *st;
    const unsigned blocks = (unsigned)((n + M - 1) / M);
    ActorJobs jobs = {};
    jobs.j[0] = ActorJob{wa->packed, nullptr, act_a, logp_a, nullptr, nullptr, obs_out, nullptr, step_a, wa->max_action, wa->use_tanh};
    jobs.j[1] = ActorJob{wb->packed, nullptr, act_b, logp_b, nullptr, nullptr, nullptr, nullptr, step_b, wb->max_action, wb->use_tanh};
    actor_kernel<false, true><<<dim3(blocks, 2), THREADS, sizeof(Smem), (cudaStream_t)stream>>>(jobs, obs_f32, s0, obs_stats, n, row_offset, seed);
    return launch_status();
}

}  // extern "C"
