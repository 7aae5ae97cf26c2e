// env_step.cu -- K2: fused batched satellites.step()/reset() for sm_100a.
//
// Reference behaviour replaced (paths relative to the upstream repo):
//   environment.py:66-79 (reset), :81-179 (step, Flag 0), :181-255 (step, Flag 1),
//   :317-343 (danger-zone count, relative->inertial), :346-396 (reward cosines),
//   satellite_function.py:744-781 (CW STM), :18-99,161-255,317-373,462-565 (danger zone + fsolve),
//   normalization.py:19-29,56-60 (running statistics of observations / discounted returns).
//
// Mapping: TWO lanes per environment (even lane = pursuer, odd lane = escaper). Each lane owns one
// craft through impulse + propagation + orbital elements, the pair exchanges results with
// __shfl_xor(…, 1), and each lane then evaluates one of the two relative-node reachability
// solves. 64-thread CTAs (32 envs) keep the CTA count a near-multiple of 148 SMs x resident CTAs
// at the headline batch (65 536 envs -> 2048 CTAs = 13.8 per SM).
//
// Data layout in HBM: SoA fp64 columns [16][ld] + int32 columns [4][ld] (include/satb200.h).
// Compiled with -fmad=false: the env arithmetic must round exactly like numpy's.
#include "sat_math.cuh"
#include "../../include/satb200.h"
#include <cmath>

namespace {

using namespace sat;

constexpr int kBlock = 64;              // threads per CTA
constexpr int kEnvsPerBlock = kBlock / 2;
constexpr int kObs = 18;
constexpr int kStatDims = 19;           // 18 observation dims + discounted return

SAT_DEV double shfl1(double v) { return __shfl_xor_sync(0xffffffffu, v, 1); }
SAT_DEV int shfl1(int v) { return __shfl_xor_sync(0xffffffffu, v, 1); }

SAT_DEV double clip16(double a) { return a < -1.6 ? -1.6 : (a > 1.6 ? 1.6 : a); }   // np.clip, environment.py:86-87

template <int MODE, typename ActT>
__global__ void __launch_bounds__(kBlock)
env_step_kernel(const SatEnvState st, const ActT* __restrict__ pa, const ActT* __restrict__ ea,
                const int32_t* __restrict__ count_override, float* __restrict__ obs_f32,
                double* __restrict__ obs_f64, double* __restrict__ term_obs_f64,
                double* __restrict__ reward_out, uint8_t* __restrict__ done_out,
                double* __restrict__ partials, const __grid_constant__ SatEnvParams p) {
    __shared__ double tile[kEnvsPerBlock][kStatDims];        // next observation (+ return) of the CTA's envs
    __shared__ double tile_term[kEnvsPerBlock][kObs];        // pre-reset observation

    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t env_raw = tid >> 1;
    const int craft = (int)(tid & 1);
    const bool valid = env_raw < st.n;
    const int64_t e = valid ? env_raw : st.n - 1;
    const int64_t ld = st.ld;
    double* __restrict__ S = st.state;
    int32_t* __restrict__ I = st.istate;

    // ---------------- load own craft + env scalars
    double r[3], v[3];
    const int base = craft * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) { r[k] = S[(base + k) * ld + e]; v[k] = S[(base + 3 + k) * ld + e]; }
    double fuel_own = S[(SAT_COL_FUEL_C + craft) * ld + e];
    const double dis_stale = S[SAT_COL_DIS * ld + e];
    double ret = S[SAT_COL_RET * ld + e];
    const int dz_stale = I[SAT_ICOL_DZ * ld + e];
    const int count = I[SAT_ICOL_COUNT * ld + e];
    const int int_state = I[SAT_ICOL_INTSTATE * ld + e];
    int err = I[SAT_ICOL_ERR * ld + e];

    // ---------------- actions: clip, gate (environment.py:86-104 / :185-198)
    double a[3];
    {
        const ActT* act = craft ? ea : pa;
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] = clip16((double)act[e * 3 + k]);
    }
    bool frozen;
    if (p.flag == 0) frozen = (craft == 0) && (dis_stale < p.d_range) && (dz_stale != 0);
    else frozen = (craft == 1) && (dz_stale == 0);
    if (frozen) { a[0] = 0.0; a[1] = 0.0; a[2] = 0.0; }

    // previous distance (:89) needs the other craft's position
    double o_r[3], o_v[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) o_r[k] = shfl1(r[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = craft == 0 ? __dsub_rn(r[k], o_r[k]) : __dsub_rn(o_r[k], r[k]);
    const double dis_prev = norm3(d);

    // impulse; right after reset() the arrays are int64 and `+=` truncates toward zero (Q1)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double s = __dadd_rn(v[k], a[k]);
        v[k] = int_state ? trunc(s) : s;
    }
    fuel_own = __dsub_rn(fuel_own, __dadd_rn(__dadd_rn(fabs(a[0]), fabs(a[1])), fabs(a[2])));   // :106-107

    // ---------------- propagate own craft
    if (MODE == SAT_MODE_CW) {
        double x6[6] = {r[0], r[1], r[2], v[0], v[1], v[2]}, y6[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) y6[i] = gemv6_row(p.stm + 6 * i, x6);     // satellite_function.py:778-779
#pragma unroll
        for (int k = 0; k < 3; ++k) { r[k] = y6[k]; v[k] = y6[3 + k]; }
    } else {
        double X[3], V[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { X[k] = p.r_cw[k] + r[k]; V[k] = p.v_cw[k] + v[k]; }
        const Rk4Consts c = make_rk4_consts(p.h, p.mu, p.re, p.j2);
        if (p.j2 != 0.0) {
#pragma unroll 1
            for (int s = 0; s < p.substeps; ++s) rk4_step<true>(X, V, c);
        } else {
#pragma unroll 1
            for (int s = 0; s < p.substeps; ++s) rk4_step<false>(X, V, c);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { r[k] = X[k] - p.r_cw[k]; v[k] = V[k] - p.v_cw[k]; }
    }

    // ---------------- exchange, distance, terminal checks (:130-147)
#pragma unroll
    for (int k = 0; k < 3; ++k) { o_r[k] = shfl1(r[k]); o_v[k] = shfl1(v[k]); }
    double P[3], Pv[3], E[3], Ev[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        P[k] = craft == 0 ? r[k] : o_r[k];  Pv[k] = craft == 0 ? v[k] : o_v[k];
        E[k] = craft == 0 ? o_r[k] : r[k];  Ev[k] = craft == 0 ? o_v[k] : v[k];
        d[k] = __dsub_rn(P[k], E[k]);
    }
    const double dis = norm3(d);                                              // :132
    const int count_new = count_override ? count_override[e] : count + 1;
    const bool captured = dis <= p.d_capture;                                 // :139
    const bool timeout = count_new >= p.max_episode_steps;                    // :144
    const bool done = captured || timeout;
    const double fuel_oth = shfl1(fuel_own);
    const double fuel_c = craft == 0 ? fuel_own : fuel_oth;                   // Delta_V_c = fuel_c (:328)
    double pa_gated[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { double o = shfl1(a[k]); pa_gated[k] = craft == 0 ? a[k] : o; }

    // ---------------- danger-zone count (:150 -> :317-332); each lane: own craft's elements
    const bool need_dz = !done && !p.skip_danger_zone;
    double Ri[3], Vi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Ri[k] = __dadd_rn(p.r_cw[k], r[k]); Vi[k] = __dadd_rn(p.v_cw[k], v[k]); }   // :338-341
    const int dz_eval = danger_zone_pair(craft, need_dz, Ri, Vi, fuel_c, p.u_grav, nullptr);
    int dz_new = dz_stale;
    if (need_dz) {
        if (dz_eval >= 0) dz_new = dz_eval;
        else { dz_new = 0; err = 1; }        // the reference raises here (circular / parabolic element set)
    }

    // ---------------- reward (:139-147, :161-175, Flag 1 :221-251)
    double reward;
    if (captured) reward = (p.flag == 0) ? 100.0 : -150.0;
    else if (timeout) reward = (p.flag == 0) ? 0.0 : 100.0;
    else {
        const double ra = (dis < dis_prev) ? 1.0 : -1.0;                                      // :161
        const double rb = (p.d_capture <= dis && dis <= 4.0 * p.d_capture) ? -1.0 : -2.0;     // :162
        const double rc = (dz_new == 0) ? -1.0 : dz_new * 0.5;                                // :164
        const double pv1 = cosine3(P, E);                                                     // :166
        const double pv2 = cosine3(Pv, Ev);                                                   // :167
        const double pv3 = cosine3(d, Pv);                                                    // :168
        double pv4 = 0.0;                                                                     // :169
        if (pa_gated[0] != 0.0 && pa_gated[1] != 0.0 && pa_gated[2] != 0.0) pv4 = -cosine3(d, pa_gated);
        double rr = __dadd_rn(__dadd_rn(ra, rb), rc);
        rr = __dadd_rn(rr, pv1);                                                              // :172-175
        rr = __dadd_rn(rr, __dmul_rn(0.6, pv2));
        rr = __dadd_rn(rr, __dmul_rn(0.2, pv3));
        rr = __dadd_rn(rr, __dmul_rn(2.0, pv4));
        reward = (p.flag == 0) ? rr : -rr;                                                    // :251
    }
    // discounted return of RewardScaling (normalization.py:57), reset on done (:62-63)
    const double ret_new = __dadd_rn(__dmul_rn(p.gamma, ret), reward);

    // ---------------- observation before reset (what the reference returns as s_)
    const int le = threadIdx.x >> 1;       // env slot within the CTA
    if (term_obs_f64) {
        if (craft == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { tile_term[le][k] = d[k]; tile_term[le][3 + k] = __dsub_rn(Pv[k], Ev[k]); tile_term[le][6 + k] = P[k]; }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) { tile_term[le][9 + k] = Pv[k]; tile_term[le][12 + k] = E[k]; tile_term[le][15 + k] = Ev[k]; }
        }
    }

    // ---------------- auto reset (environment.py:66-79; fuel/dis/dangerous_zone persist, Q2)
    int int_state_new = 0, count_store = count_new;
    if (done && p.auto_reset) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            P[k] = p.reset_p[k]; E[k] = p.reset_e[k]; Pv[k] = 0.0; Ev[k] = 0.0;
            r[k] = craft == 0 ? P[k] : E[k]; v[k] = 0.0;
        }
        int_state_new = 1; count_store = 0;
    }

    // ---------------- store state
    if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { S[(base + k) * ld + e] = r[k]; S[(base + 3 + k) * ld + e] = v[k]; }
        S[(SAT_COL_FUEL_C + craft) * ld + e] = fuel_own;
        if (craft == 0) {
            S[SAT_COL_DIS * ld + e] = dis;
            S[SAT_COL_RET * ld + e] = done ? 0.0 : ret_new;
            I[SAT_ICOL_DZ * ld + e] = dz_new;
            I[SAT_ICOL_COUNT * ld + e] = count_store;
            reward_out[e] = reward;
            done_out[e] = done ? 1 : 0;
        } else {
            I[SAT_ICOL_INTSTATE * ld + e] = int_state_new;
            I[SAT_ICOL_ERR * ld + e] = err;
        }
    }

    // ---------------- next observation tile -> coalesced stores + per-CTA statistics partial
    if (craft == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { tile[le][k] = __dsub_rn(P[k], E[k]); tile[le][3 + k] = __dsub_rn(Pv[k], Ev[k]); tile[le][6 + k] = P[k]; }
        tile[le][18] = ret_new;
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { tile[le][9 + k] = Pv[k]; tile[le][12 + k] = E[k]; tile[le][15 + k] = Ev[k]; }
    }
    __syncthreads();
    const int64_t env0 = (int64_t)blockIdx.x * kEnvsPerBlock;
    const int64_t rem = st.n - env0;
    const int rows = (int)(rem < kEnvsPerBlock ? rem : kEnvsPerBlock);
    const int total = rows * kObs;
    for (int idx = threadIdx.x; idx < total; idx += kBlock) {
        const int row = idx / kObs, col = idx - row * kObs;
        const double val = tile[row][col];
        if (obs_f32) obs_f32[env0 * kObs + idx] = (float)val;
        if (obs_f64) obs_f64[env0 * kObs + idx] = val;
        if (term_obs_f64) term_obs_f64[env0 * kObs + idx] = tile_term[row][col];
    }
    if (partials && threadIdx.x < kStatDims) {
        // two-pass mean / M2 over the CTA's rows in a fixed order (deterministic)
        const int dim = threadIdx.x;
        double sum = 0.0;
        for (int row = 0; row < rows; ++row) sum += tile[row][dim];
        const double mean = sum / rows;
        double m2 = 0.0;
        for (int row = 0; row < rows; ++row) { double t = tile[row][dim] - mean; m2 += t * t; }
        partials[((int64_t)blockIdx.x * kStatDims + dim) * 2 + 0] = mean;
        partials[((int64_t)blockIdx.x * kStatDims + dim) * 2 + 1] = m2;
    }
}


// ---------------------------------------------------------------------------------------------
// batched danger-zone count on explicit inertial states (Time_window_of_danger_zone(...)
// .calculate_number_of_hanger_area(), satellite_function.py:18-99, 341-373)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
danger_zone_kernel(const double* __restrict__ rv, const double* __restrict__ dv, int64_t n, double u_grav,
                   int32_t* __restrict__ count_out, double* __restrict__ debug_out) {
    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t env_raw = tid >> 1;
    const int craft = (int)(tid & 1);
    const bool valid = env_raw < n;
    const int64_t e = valid ? env_raw : n - 1;
    double Ri[3], Vi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Ri[k] = rv[e * 12 + craft * 6 + k]; Vi[k] = rv[e * 12 + craft * 6 + 3 + k]; }
    DzDebug dbg = {0, 0, 0, 0, 0, 0, 0, 0};
    const int dz = danger_zone_pair(craft, true, Ri, Vi, dv[e], u_grav, debug_out ? &dbg : nullptr);
    if (valid) {
        if (craft == 0) count_out[e] = dz;
        if (debug_out) {
            double* o = debug_out + (e * 2 + craft) * 8;
            o[0] = dbg.rf_max; o[1] = dbg.rf_min; o[2] = dbg.r_ft; o[3] = dbg.alpha0; o[4] = dbg.alpha1;
            o[5] = dbg.theta; o[6] = dbg.dvm; o[7] = dbg.f_cx;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// merge of per-CTA partials into the running statistics (RunningMeanStd, normalization.py:19-29).
// One warp per statistic dimension, fixed merge order -> bitwise deterministic.
// ---------------------------------------------------------------------------------------------
struct Moments { double n, mean, m2; };
SAT_DEV Moments chan(Moments A, Moments B) {
    if (B.n == 0.0) return A;
    if (A.n == 0.0) return B;
    Moments C;
    C.n = A.n + B.n;
    const double delta = B.mean - A.mean;
    C.mean = A.mean + delta * (B.n / C.n);
    C.m2 = A.m2 + B.m2 + delta * delta * (A.n * B.n / C.n);
    return C;
}

SAT_DEV void fold_into_running(double* stats, int dim, int d, Moments B, double* std_out) {
    // stats: [0]=n, mean[dim], S[dim], std[dim]
    double n_old = stats[0];
    double mean_old = stats[1 + d], S_old = stats[1 + dim + d];
    double n_new, mean_new, S_new, std_new;
    if (B.n == 1.0) {
        // literal Welford step of the reference incl. its first-sample rule (normalization.py:21-29)
        const double x = B.mean;
        n_new = n_old + 1.0;
        if (n_new == 1.0) { mean_new = x; S_new = S_old; std_new = x; }
        else {
            mean_new = mean_old + (x - mean_old) / n_new;
            S_new = S_old + (x - mean_old) * (x - mean_new);
            std_new = sqrt(S_new / n_new);
        }
    } else {
        Moments A = {n_old, mean_old, S_old};
        Moments C = chan(A, B);
        n_new = C.n; mean_new = C.mean; S_new = C.m2;
        std_new = sqrt(S_new / n_new);
    }
    stats[1 + d] = mean_new; stats[1 + dim + d] = S_new; stats[1 + 2 * dim + d] = std_new;
    if (std_out) *std_out = std_new;
    (void)n_new;
}

__global__ void stats_merge_kernel(const double* __restrict__ partials, int64_t nblocks, int64_t n_rows,
                                   int rows_per_block, int ndims_in_partials,
                                   double* obs_stats, int obs_dim, double* ret_stats, double* ret_std_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active = warp < ndims_in_partials;
    Moments acc = {0.0, 0.0, 0.0};
    for (int64_t b = lane; active && b < nblocks; b += 32) {
        int64_t rem = n_rows - b * rows_per_block;
        Moments B;
        B.n = (double)(rem < rows_per_block ? rem : rows_per_block);
        B.mean = partials[(b * ndims_in_partials + warp) * 2 + 0];
        B.m2 = partials[(b * ndims_in_partials + warp) * 2 + 1];
        acc = chan(acc, B);
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        Moments O;
        O.n = __shfl_xor_sync(0xffffffffu, acc.n, off);
        O.mean = __shfl_xor_sync(0xffffffffu, acc.mean, off);
        O.m2 = __shfl_xor_sync(0xffffffffu, acc.m2, off);
        // keep a fixed (lower lane first) operand order so both partners compute the same value
        acc = (lane & off) ? chan(O, acc) : chan(acc, O);
    }
    if (active && lane == 0) {
        if (warp < obs_dim) { if (obs_stats) fold_into_running(obs_stats, obs_dim, warp, acc, nullptr); }
        else if (ret_stats) fold_into_running(ret_stats, 1, 0, acc, ret_std_out);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (obs_stats) obs_stats[0] += (double)n_rows;
        if (ret_stats) ret_stats[0] += (double)n_rows;
    }
}

// ---------------------------------------------------------------------------------------------
// init / reset / observe
// ---------------------------------------------------------------------------------------------
__global__ void env_init_kernel(const SatEnvState st, double fuel_c, double fuel_t,
                                const __grid_constant__ SatEnvParams p, const uint8_t* __restrict__ mask, int full) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.n) return;
    if (mask && !mask[e]) return;
    const int64_t ld = st.ld;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        st.state[(SAT_COL_P + k) * ld + e] = p.reset_p[k];     // environment.py:67
        st.state[(SAT_COL_PV + k) * ld + e] = 0.0;             // :68
        st.state[(SAT_COL_E + k) * ld + e] = p.reset_e[k];     // :70
        st.state[(SAT_COL_EV + k) * ld + e] = 0.0;             // :71
    }
    st.istate[SAT_ICOL_INTSTATE * ld + e] = 1;                 // int64 arrays until the first step (Q1)
    st.istate[SAT_ICOL_COUNT * ld + e] = 0;
    if (full) {                                                // constructor, environment.py:41-44
        st.state[SAT_COL_FUEL_C * ld + e] = fuel_c;
        st.state[SAT_COL_FUEL_T * ld + e] = fuel_t;
        st.state[SAT_COL_DIS * ld + e] = INFINITY;
        st.state[SAT_COL_RET * ld + e] = 0.0;
        st.istate[SAT_ICOL_DZ * ld + e] = 0;
        st.istate[SAT_ICOL_ERR * ld + e] = 0;
    }
}

__global__ void env_observe_kernel(const SatEnvState st, float* __restrict__ obs_f32, double* __restrict__ obs_f64) {
    // one thread per (env, obs element): reads are column-coalesced per element class, writes fully coalesced
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= st.n * kObs) return;
    const int64_t e = idx / kObs;
    const int j = (int)(idx - e * kObs);
    const int64_t ld = st.ld;
    const int k = j % 3, g = j / 3;
    double val;
    if (g == 0) val = __dsub_rn(st.state[(SAT_COL_P + k) * ld + e], st.state[(SAT_COL_E + k) * ld + e]);
    else if (g == 1) val = __dsub_rn(st.state[(SAT_COL_PV + k) * ld + e], st.state[(SAT_COL_EV + k) * ld + e]);
    else val = st.state[((g - 2) * 3 + k) * ld + e];
    if (obs_f32) obs_f32[idx] = (float)val;
    if (obs_f64) obs_f64[idx] = val;
}


// ---------------------------------------------------------------------------------------------
// stand-alone batch normalisation (Normalization.__call__, normalization.py:37-43)
// ---------------------------------------------------------------------------------------------
constexpr int kNormRows = 128;

__global__ void __launch_bounds__(kNormRows)
norm_partial_kernel(const double* __restrict__ x, int64_t n, int dim, double* __restrict__ partials) {
    __shared__ double tile[kNormRows * 32];
    const int64_t row0 = (int64_t)blockIdx.x * kNormRows;
    const int64_t rem = n - row0;
    const int rows = (int)(rem < kNormRows ? rem : kNormRows);
    const int total = rows * dim;
    for (int idx = threadIdx.x; idx < total; idx += kNormRows) tile[idx] = x[row0 * dim + idx];
    __syncthreads();
    if (threadIdx.x < dim) {
        const int d = threadIdx.x;
        double sum = 0.0;
        for (int r = 0; r < rows; ++r) sum += tile[r * dim + d];
        const double mean = sum / rows;
        double m2 = 0.0;
        for (int r = 0; r < rows; ++r) { double t = tile[r * dim + d] - mean; m2 += t * t; }
        partials[((int64_t)blockIdx.x * dim + d) * 2 + 0] = mean;
        partials[((int64_t)blockIdx.x * dim + d) * 2 + 1] = m2;
    }
}

__global__ void norm_apply_kernel(const double* __restrict__ stats, const double* __restrict__ x, int64_t total,
                                  int dim, double* __restrict__ out64, float* __restrict__ out32) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int d = (int)(idx % dim);
    const double mean = stats[1 + d], sd = stats[1 + 2 * dim + d];
    const double y = (x[idx] - mean) / (sd + 1e-8);             // normalization.py:41
    if (out64) out64[idx] = y;
    if (out32) out32[idx] = (float)y;
}

int check_state(const SatEnvState* st) {
    if (!st || !st->state || !st->istate) return SAT_ERR_NULL;
    if (st->n <= 0 || st->ld < st->n || (st->ld & 1)) return SAT_ERR_SIZE;
    if (((uintptr_t)st->state & 15) || ((uintptr_t)st->istate & 15)) return SAT_ERR_SIZE;
    return SAT_OK;
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}

}  // namespace

extern "C" {

int64_t sat_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    int64_t nb_env = (n + kEnvsPerBlock - 1) / kEnvsPerBlock;
    int64_t nb_norm = (n + kNormRows - 1) / kNormRows;
    int64_t a = nb_env * kStatDims * 2 * (int64_t)sizeof(double);
    int64_t b = nb_norm * 32 * 2 * (int64_t)sizeof(double);
    return (a > b ? a : b) + 256;
}

void sat_env_default_params(SatEnvParams* p) {
    if (!p) return;
    SatEnvParams q = {};
    q.mode = SAT_MODE_CW; q.flag = 0; q.max_episode_steps = 1000; q.auto_reset = 1;
    q.action_dtype = SAT_ACT_F32; q.substeps = 100; q.skip_danger_zone = 0;
    q.d_capture = 100000.0; q.d_range = 100000.0;                 // environment.py:28
    q.gamma = 0.99;                                               // CPPO_main.py:24
    q.h = 1.0; q.mu = 3.986e14; q.re = 6378137.0; q.j2 = 0.00108263;
    q.r_cw[0] = 27098000.0; q.r_cw[1] = 32306000.0; q.r_cw[2] = 0.0;   // environment.py:338
    q.v_cw[0] = -2350.0; q.v_cw[1] = 1970.0; q.v_cw[2] = 0.0;          // :339
    q.u_grav = 3.986e14;                                          // satellite_function.py:28
    q.reset_p[0] = 200000.0; q.reset_e[0] = 18000.0;              // environment.py:67,70
    *p = q;
}

int sat_env_init(const SatEnvState* st, double fuel_c, double fuel_t, const SatEnvParams* p, void* stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!p) return SAT_ERR_NULL;
    const int threads = 256;
    const unsigned blocks = (unsigned)((st->n + threads - 1) / threads);
    env_init_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*st, fuel_c, fuel_t, *p, nullptr, 1);
    return launch_status();
}

int sat_env_reset(const SatEnvState* st, const uint8_t* mask, const SatEnvParams* p, void* stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!p) return SAT_ERR_NULL;
    const int threads = 256;
    const unsigned blocks = (unsigned)((st->n + threads - 1) / threads);
    env_init_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*st, 0.0, 0.0, *p, mask, 0);
    return launch_status();
}

int sat_env_observe(const SatEnvState* st, float* obs_f32, double* obs_f64, void* stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!obs_f32 && !obs_f64) return SAT_ERR_NULL;
    const int threads = 256;
    const int64_t total = st->n * kObs;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    env_observe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*st, obs_f32, obs_f64);
    return launch_status();
}

int sat_env_step(const SatEnvState* st, const void* pa, const void* ea, const int32_t* count_override,
                 float* obs_f32, double* obs_f64, double* term_obs_f64, double* reward, uint8_t* done,
                 double* obs_stats, double* ret_stats, double* ret_std_out, void* workspace,
                 const SatEnvParams* p, void* stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!pa || !ea || !reward || !done || !p) return SAT_ERR_NULL;
    if ((obs_stats || ret_stats) && !workspace) return SAT_ERR_NULL;
    if (p->mode != SAT_MODE_CW && p->mode != SAT_MODE_RK4) return SAT_ERR_MODE;
    if (p->action_dtype != SAT_ACT_F32 && p->action_dtype != SAT_ACT_F64) return SAT_ERR_MODE;
    if (p->flag != 0 && p->flag != 1) return SAT_ERR_MODE;
    if (p->mode == SAT_MODE_RK4 && p->substeps < 1) return SAT_ERR_SIZE;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nblocks = (st->n + kEnvsPerBlock - 1) / kEnvsPerBlock;
    double* partials = (obs_stats || ret_stats) ? (double*)workspace : nullptr;
#define SAT_LAUNCH(MODE, T)                                                                           \
    env_step_kernel<MODE, T><<<(unsigned)nblocks, kBlock, 0, s>>>(*st, (const T*)pa, (const T*)ea,    \
        count_override, obs_f32, obs_f64, term_obs_f64, reward, done, partials, *p)
    if (p->mode == SAT_MODE_CW) {
        if (p->action_dtype == SAT_ACT_F32) SAT_LAUNCH(SAT_MODE_CW, float); else SAT_LAUNCH(SAT_MODE_CW, double);
    } else {
        if (p->action_dtype == SAT_ACT_F32) SAT_LAUNCH(SAT_MODE_RK4, float); else SAT_LAUNCH(SAT_MODE_RK4, double);
    }
#undef SAT_LAUNCH
    rc = launch_status();
    if (rc) return rc;
    if (partials) {
        stats_merge_kernel<<<1, 32 * kStatDims, 0, s>>>(partials, nblocks, st->n, kEnvsPerBlock, kStatDims,
                                                       obs_stats, kObs, ret_stats, ret_std_out);
        rc = launch_status();
    }
    return rc;
}

int sat_danger_zone_count(const double* rv, const double* dv, int64_t n, double u_grav, int32_t* count_out,
                          double* debug_out, void* stream) {
    if (!rv || !dv || !count_out) return SAT_ERR_NULL;
    if (n <= 0) return SAT_ERR_SIZE;
    const int64_t nblocks = (n + kEnvsPerBlock - 1) / kEnvsPerBlock;
    danger_zone_kernel<<<(unsigned)nblocks, kBlock, 0, (cudaStream_t)stream>>>(rv, dv, n, u_grav, count_out, debug_out);
    return launch_status();
}

int64_t sat_env_step_host_bytes(int64_t n) {
    if (n <= 0) return 0;
    // pa | ea (fp32 [n][3] each) | obs fp32 [n][18] | reward fp64 [n] | done u8 [n], each 256-byte aligned
    auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
    return al(n * 12) * 2 + al(n * 72) + al(n * 8) + al(n);
}

int sat_env_step_host(const SatEnvState* st, const float* pa_host, const float* ea_host,
                      float* obs_host, double* reward_host, uint8_t* done_host, void* d_io,
                      const SatEnvParams* p, void* stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!pa_host || !ea_host || !obs_host || !reward_host || !done_host || !d_io || !p) return SAT_ERR_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = st->n;
    auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
    char* base = (char*)d_io;
    float* d_pa = (float*)base;
    float* d_ea = (float*)(base + al(n * 12));
    float* d_obs = (float*)(base + 2 * al(n * 12));
    double* d_rew = (double*)(base + 2 * al(n * 12) + al(n * 72));
    uint8_t* d_done = (uint8_t*)(base + 2 * al(n * 12) + al(n * 72) + al(n * 8));
    cudaError_t ce;
    if ((ce = cudaMemcpyAsync(d_pa, pa_host, n * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemcpyAsync(d_ea, ea_host, n * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)ce;
    SatEnvParams q = *p;
    q.action_dtype = SAT_ACT_F32;
    rc = sat_env_step(st, d_pa, d_ea, nullptr, d_obs, nullptr, nullptr, d_rew, d_done, nullptr, nullptr, nullptr,
                      nullptr, &q, stream);
    if (rc) return rc;
    if ((ce = cudaMemcpyAsync(obs_host, d_obs, n * 72, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemcpyAsync(reward_host, d_rew, n * 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)ce;
    if ((ce = cudaMemcpyAsync(done_host, d_done, n, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)ce;
    if ((ce = cudaStreamSynchronize(s)) != cudaSuccess) return (int)ce;
    return SAT_OK;
}

int sat_norm_update(double* stats, const double* x, int64_t n, int dim, int update,
                    double* x_out_f64, float* x_out_f32, void* workspace, void* stream) {
    if (!stats || !x) return SAT_ERR_NULL;
    if (n <= 0 || dim < 1 || dim > 32) return SAT_ERR_SIZE;
    if (update && !workspace) return SAT_ERR_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (update) {
        const int64_t nblocks = (n + kNormRows - 1) / kNormRows;
        norm_partial_kernel<<<(unsigned)nblocks, kNormRows, 0, s>>>(x, n, dim, (double*)workspace);
        if ((rc = launch_status())) return rc;
        stats_merge_kernel<<<1, 32 * dim, 0, s>>>((const double*)workspace, nblocks, n, kNormRows, dim,
                                                  stats, dim, nullptr, nullptr);
        if ((rc = launch_status())) return rc;
    }
    if (x_out_f64 || x_out_f32) {
        const int64_t total = n * dim;
        norm_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(stats, x, total, dim, x_out_f64, x_out_f32);
        if ((rc = launch_status())) return rc;
    }
    return SAT_OK;
}

}  // extern "C"
