// cw_ode.cu -- batched Numerical_calculation_method.numerical_calculation(t) (satellite_function.py:783-839).
//
// The reference integrates the Clohessy-Wiltshire ODE (orbit_ode, :793-821; thrust and J2 terms are identically
// zero there) with scipy.integrate.solve_ivp(method="RK45", rtol=1e-3, atol=1e-6) and returns the value at the
// last t_eval point (= t). This kernel restates scipy 1.18.1's adaptive Dormand-Prince 5(4) step control
// (select_initial_step, RungeKutta._step_impl, RkDenseOutput) with one thread per state. The reference env has this
// propagator commented out (environment.py:123-128); it is provided for interface completeness, not for speed:
// every thread follows its own step sequence (3-5 accepted steps for t = 100..1000 s).
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/satb200.h"

namespace {

__device__ __forceinline__ void cw_rhs(const double X[6], double w2, double w3, double wz, double f[6]) {
    f[0] = X[3]; f[1] = X[4]; f[2] = X[5];
    f[3] = w2 * X[4] + w3 * X[0];        // 2 omega ydot + 3 omega^2 x   (:818)
    f[4] = (-w2) * X[3];                 // -2 omega xdot                (:819)
    f[5] = (-wz) * X[2];                 // -omega^2 z                   (:820)
}
__device__ __forceinline__ double rms6(const double x[6]) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += x[k] * x[k];
    return sqrt(s) / sqrt(6.0);
}

__constant__ double cA[6][5] = {{0, 0, 0, 0, 0}, {0.2, 0, 0, 0, 0}, {0.075, 0.225, 0, 0, 0},
    {0.9777777777777777, -3.7333333333333334, 3.5555555555555554, 0, 0},
    {2.9525986892242035, -11.595793324188385, 9.822892851699436, -0.2908093278463649, 0},
    {2.8462752525252526, -10.757575757575758, 8.906422717743473, 0.2784090909090909, -0.2735313036020583}};
__constant__ double cB[6] = {0.09114583333333333, 0.0, 0.44923629829290207, 0.6510416666666666, -0.322376179245283, 0.13095238095238096};
__constant__ double cE[7] = {-0.0012326388888888888, 0.0, 0.0042527702905061394, -0.03697916666666667, 0.05086379716981132, -0.0419047619047619, 0.025};
__constant__ double cP[7][4] = {{1.0, -2.8535800653862835, 3.0717434641059005, -1.1270175653862835}, {0, 0, 0, 0},
    {0.0, 4.023133379230305, -6.249321565289, 2.675424484351598}, {0.0, -3.7324019615885042, 10.068970589843675, -5.685526961588504},
    {0.0, 2.5548038301849423, -6.399112377351017, 3.5219323679207912}, {0.0, -1.3744241142186024, 3.272657752246729, -1.7672812570757455},
    {0.0, 1.3824689317781436, -3.764937863556287, 2.382468931778144}};

__global__ void __launch_bounds__(128)
cw_ode_rk45_kernel(double* __restrict__ x, int64_t n, int64_t ld, double t_bound, double w2, double w3, double wz,
                   double rtol, double atol, int32_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0, err_exp = -0.2;
    double y[6], f[6], K[7][6], y_new[6], f_new[6], y_old[6], tmp[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] = x[k * ld + i];
    cw_rhs(y, w2, w3, wz, f);
    double h_abs, h = 0.0, t = 0.0;
    {   // select_initial_step (scipy/integrate/_ivp/common.py)
        double scale[6], y1[6], f1[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { scale[k] = atol + fabs(y[k]) * rtol; tmp[k] = y[k] / scale[k]; }
        const double d0 = rms6(tmp);
#pragma unroll
        for (int k = 0; k < 6; ++k) tmp[k] = f[k] / scale[k];
        const double d1 = rms6(tmp);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, fabs(t_bound));
#pragma unroll
        for (int k = 0; k < 6; ++k) y1[k] = y[k] + h0 * 1.0 * f[k];
        cw_rhs(y1, w2, w3, wz, f1);
#pragma unroll
        for (int k = 0; k < 6; ++k) tmp[k] = (f1[k] - f[k]) / scale[k];
        const double d2 = rms6(tmp) / h0;
        const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), fabs(t_bound));
    }
    int st = 0;
    while (t != t_bound && st == 0) {                        // OdeSolver.step until t reaches t_bound
        const double t_up = __longlong_as_double(__double_as_longlong(t) + 1);   // np.nextafter(t, +inf) for t >= 0
        const double min_step = 10 * fabs(t_up - t);
        double t_new = t;
        bool accepted = false, rejected = false;
        if (h_abs < min_step) h_abs = min_step;
        while (!accepted) {
            if (h_abs < min_step) { st = -1; break; }        // TOO_SMALL_STEP
            h = h_abs; t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t; h_abs = fabs(h);
#pragma unroll
            for (int k = 0; k < 6; ++k) K[0][k] = f[k];
#pragma unroll
            for (int s = 1; s < 6; ++s) {                    // rk_step
                double ys[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    double dy = 0.0;
#pragma unroll
                    for (int j = 0; j < s; ++j) dy += K[j][k] * cA[s][j];
                    ys[k] = y[k] + dy * h;
                }
                cw_rhs(ys, w2, w3, wz, K[s]);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += K[j][k] * cB[j];
                y_new[k] = y[k] + h * acc;
            }
            cw_rhs(y_new, w2, w3, wz, f_new);
#pragma unroll
            for (int k = 0; k < 6; ++k) K[6][k] = f_new[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double e = 0.0;
#pragma unroll
                for (int j = 0; j < 7; ++j) e += K[j][k] * cE[j];
                const double sc = atol + fmax(fabs(y[k]), fabs(y_new[k])) * rtol;
                tmp[k] = (e * h) / sc;
            }
            const double error_norm = rms6(tmp);
            if (error_norm < 1.0) {
                double factor = (error_norm == 0.0) ? MAX_FACTOR : fmin(MAX_FACTOR, SAFETY * pow(error_norm, err_exp));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor; accepted = true;
            } else { h_abs *= fmax(MIN_FACTOR, SAFETY * pow(error_norm, err_exp)); rejected = true; }
        }
        if (st) break;
#pragma unroll
        for (int k = 0; k < 6; ++k) { y_old[k] = y[k]; y[k] = y_new[k]; f[k] = f_new[k]; }
        t = t_new;
    }
    if (st == 0) {
        // solve_ivp takes the value at t_eval[-1] == t_bound from the dense output of the last step at x = 1
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double q = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double c = 0.0;
#pragma unroll
                for (int s = 0; s < 7; ++s) c += K[s][k] * cP[s][j];
                q += c;
            }
            x[k * ld + i] = h * q + y_old[k];
        }
    }
    if (status) status[i] = st;
}

}  // namespace

extern "C" int sat_cw_ode_rk45(double* x, int64_t n, int64_t ld, double t_bound, double w2, double w3, double wz,
                               double rtol, double atol, int32_t* status_out, void* stream) {
    if (!x) return SAT_ERR_NULL;
    if (n <= 0 || ld < n || !(t_bound > 0.0) || !(rtol > 0.0) || !(atol > 0.0)) return SAT_ERR_SIZE;
    cw_ode_rk45_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, n, ld, t_bound, w2, w3, wz, rtol, atol, status_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SAT_OK : (int)e;
}
