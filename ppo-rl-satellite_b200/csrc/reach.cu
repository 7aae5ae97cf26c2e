// reach.cu -- batched reachable-domain sweep, reusing the device MINPACK hybrd.
//
// Reference behaviour replaced: single_pluse_model/RD_single_pulse.py:40-148 (Reachable_Domain with N1 = 1): for every
// direction (gama_i = 2 pi i / N2, alpha_j = -pi/2 + pi j / N3), i <= N2, j <= N3, the reachability test (:82), beta /
// dVm / theta (:83-91) and the two fsolve extremes from alpha_guess = +-pi/2 (:93-121); the point clouds
// RF_max = max(|rf|) P, RF_min = min(|rf|) P (:123-124). The reference runs this sweep offline (201 x 201 directions x
// 2 fsolve per state, minutes per state in Python); the ellipse fit that follows (curve_fitting.py, sklearn) is not
// part of this library.
//
// One thread per direction, grid.y = state. Only a thin band of directions around the orbital plane passes the
// reachability test, so the CTA compacts the root problems of its directions through the same shared-memory queue /
// persistent-lane solver the env step uses.
#include "sat_math.cuh"
#include "../../include/satb200.h"

namespace {
using namespace sat;
constexpr int kThreads = 128;

struct Queue {
    double A[2 * kThreads], sth[2 * kThreads], dvm[2 * kThreads], alpha[2 * kThreads];
    int guess[2 * kThreads];
    int count, next;
};

__global__ void __launch_bounds__(kThreads)
reach_kernel(const double* __restrict__ elements, const double* __restrict__ delta_max, int N2, int N3, double u,
             double* __restrict__ rf_max, double* __restrict__ rf_min, uint8_t* __restrict__ valid) {
    __shared__ Queue q;
    if (threadIdx.x == 0) { q.count = 0; q.next = 0; }
    __syncthreads();
    const int64_t state = blockIdx.y;
    const int64_t per_state = (int64_t)(N2 + 1) * (N3 + 1);
    const int64_t d = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool in_range = d < per_state;
    const double* el = elements + state * 6;
    const double a = el[0], e0 = el[1], f = el[5];
    const double dV = -delta_max[state] + 2.0 * delta_max[state] * 1.0 / 1.0;      // :65 with N1 = jj = 1
    double sf, cf;
    sincos(f, &sf, &cf);
    const double k = 1.0 + e0 * cf;
    const double one_m_e2 = 1.0 - e0 * e0;
    const double r0 = a * one_m_e2 / k;                                            // :47
    const double p0 = a * one_m_e2;                                                // :48
    double P[3] = {0, 0, 0}, theta = 0.0, dvm = 0.0, sq_e_sin = 0.0, sq_k = 0.0, sth = 0.0, cth = 1.0, A0 = 0.0, A1 = 0.0;
    bool reach = false;
    if (in_range) {
        const int i = (int)(d / (N3 + 1)), j = (int)(d - (int64_t)i * (N3 + 1));
        const double gama = kTwoPi * i / N2;                                       // :67
        const double alpha = -kPi / 2 + kPi * j / N3;                              // :69
        double sg, cg, sa, ca;
        sincos(gama, &sg, &cg); sincos(alpha, &sa, &ca);
        P[0] = sg * ca; P[1] = cg * ca; P[2] = sa;                                 // :72
        const double df = gama - f;
        double sdf, cdf;
        sincos(df, &sdf, &cdf);
        const double temp1 = (sdf * sdf) / (u * (k * k) / (p0 * (dV * dV)) - 1.0); // :80
        const double ta = tan(alpha), t2 = ta * ta;
        if (0.0 <= t2 && t2 <= temp1) {                                            // :82
            reach = true;
            const double beta = atan(ta / sdf);                                    // :83
            double sb, cb;
            sincos(beta, &sb, &cb);
            dvm = sqrt(dV * dV - u * (k * k) * (sb * sb) / p0);                    // :85
            if ((-kTwoPi <= df && df < -kPi) || (0.0 <= df && df < kPi)) theta = acos(cdf * ca);             // :88-89
            else if ((-kPi <= df && df < 0.0) || (kPi <= df && df < kTwoPi)) theta = kTwoPi - acos(cdf * ca);  // :90-91
            sincos(theta, &sth, &cth);
            const double sq = sqrt(u / p0);
            sq_e_sin = sq * e0 * sf;                                               // :96 first term
            sq_k = sq * k * cb;                                                    // :97 first term
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const double ag = g == 0 ? kPi / 2 : -kPi / 2;                     // :94 / :109
                double s_, c_;
                sincos(ag, &s_, &c_);
                const double v1x = sq_e_sin + dvm * c_, v1y = sq_k + dvm * s_;
                const double h = r0 * v1y;                                         // :99
                const double A = (2.0 * u * (1.0 - cth)) / (h * v1y) - v1x * sth / v1y;   // :151
                if (g == 0) A0 = A; else A1 = A;
            }
        }
    }
    // ---- queue the non-degenerate root problems of this CTA's directions, solve them densely
    int slot0 = -1, slot1 = -1;
    double al0 = kPi / 2, al1 = -kPi / 2;
    if (reach) {
        const bool g0 = dz_degenerate(A0, sth, dvm), g1 = dz_degenerate(A1, sth, dvm);
        const int cnt = (g0 ? 0 : 1) + (g1 ? 0 : 1);
        if (cnt) {
            int s = atomicAdd(&q.count, cnt);
            if (!g0) { slot0 = s; q.A[s] = A0; q.sth[s] = sth; q.dvm[s] = dvm; q.guess[s] = 0; ++s; }
            if (!g1) { slot1 = s; q.A[s] = A1; q.sth[s] = sth; q.dvm[s] = dvm; q.guess[s] = 1; }
        }
    }
    __syncthreads();
    {
        const int total = q.count;
        int task = -1;
        Hybrd1<PFai> hs;
        for (;;) {
            if (task < 0) {
                const int t = atomicAdd(&q.next, 1);
                if (t < total) {
                    task = t;
                    PFai fn; fn.A = q.A[t]; fn.sth = q.sth[t]; fn.dvm = q.dvm[t];
                    hs.init(fn, dz_guess(q.guess[t]));
                } else task = total;
            }
            const bool active = task < total;
            if (!__any_sync(0xffffffffu, active)) break;
            if (active && hs.step()) { q.alpha[task] = hs.x; task = -1; }
        }
    }
    __syncthreads();
    if (!in_range) return;
    const int64_t o = state * per_state + d;
    double hi = 0.0, lo = 0.0;
    if (reach) {
        if (slot0 >= 0) al0 = q.alpha[slot0];
        if (slot1 >= 0) al1 = q.alpha[slot1];
        double rf[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            double s_, c_;
            sincos(g == 0 ? al0 : al1, &s_, &c_);
            const double vx = sq_e_sin + dvm * c_, vy = sq_k + dvm * s_;           // :103-104 / :117-118
            const double hm = r0 * vy;
            rf[g] = fabs((hm * hm) / (u * (1.0 - cth) + hm * vy * cth - hm * vx * sth));   // :107 / :121
        }
        hi = fmax(rf[0], rf[1]); lo = fmin(rf[0], rf[1]);                          // :123-124
    }
    valid[o] = reach ? 1 : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { rf_max[o * 3 + c] = hi * P[c]; rf_min[o * 3 + c] = lo * P[c]; }
}
}  // namespace

extern "C" int sat_reachable_domain(const double* elements, const double* delta_max, int64_t n, int N2, int N3, double u,
                                    double* rf_max, double* rf_min, uint8_t* valid, void* stream) {
    if (!elements || !delta_max || !rf_max || !rf_min || !valid) return SAT_ERR_NULL;
    if (n <= 0 || n > 65535 || N2 < 1 || N3 < 1) return SAT_ERR_SIZE;
    const int64_t per_state = (int64_t)(N2 + 1) * (N3 + 1);
    dim3 grid((unsigned)((per_state + kThreads - 1) / kThreads), (unsigned)n);
    reach_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(elements, delta_max, N2, N3, u, rf_max, rf_min, valid);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}
