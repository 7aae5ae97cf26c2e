// Host instantiation of the product's device math header (csrc/glibm.cuh), compared bit for bit with the system
// libm -- the third-party arithmetic the reference's danger-zone count runs on (satellite_function.py:558-565 etc.).
// Built and run by tests/test_glibm_host.py and tools/soak_glibm.py:
//     g++ -O2 -std=c++17 -ffp-contract=off -mfma -I ppo-rl-satellite_b200/csrc tests/glibm_host.cpp -o glibm_host -lm
//     ./glibm_host <samples per distribution> <seed>
// Prints one line per (function, distribution): name, samples, mismatches; exit status 1 if any mismatch.
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "glibm.cuh"

static uint64_t rng_state;
static inline uint64_t rnd() {   // xorshift64*
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return rng_state * 0x2545F4914F6CDD1Dull;
}
static inline double uni() { return (double)(rnd() >> 11) * 0x1p-53; }                       // [0, 1)
static inline double sgn() { return (rnd() & 1) ? 1.0 : -1.0; }
static inline double logu(double lo_exp, double hi_exp) { return std::exp2(lo_exp + (hi_exp - lo_exp) * uni()) * sgn(); }
static inline double rawbits() { return glibm::dbl_(rnd()); }

typedef double (*fn1)(double);
static long total_bad = 0;

static bool same(double a, double b) { return glibm::bits_(a) == glibm::bits_(b) || (a != a && b != b); }

template <class Gen>
static void run(const char* name, const char* dist, fn1 mine, fn1 ref, long n, Gen gen) {
    long bad = 0;
    for (long i = 0; i < n; ++i) {
        const double x = gen();
        const double a = mine(x), b = ref(x);
        if (!same(a, b)) {
            if (bad < 3) std::printf("  MISMATCH %s(%a) = %a, libm %a\n", name, x, a, b);
            ++bad;
        }
    }
    std::printf("%s %s %ld %ld\n", name, dist, n, bad);
    total_bad += bad;
}

static double ref_pow2(double x) { volatile double two = 2.0; return std::pow(x, two); }
static double my_sin(double x) { return glibm::sin(x); }
static double my_cos(double x) { return glibm::cos(x); }
static double my_sc_s(double x) { double s, c; glibm::sincos(x, &s, &c); return s; }
static double my_sc_c(double x) { double s, c; glibm::sincos(x, &s, &c); return c; }
static double my_acos(double x) { return glibm::acos(x); }
static double my_atan(double x) { return glibm::atan(x); }
static double my_pow2(double x) { return glibm::pow2(x); }
static double ref_sin(double x) { return std::sin(x); }
static double ref_cos(double x) { return std::cos(x); }
static double ref_acos(double x) { return std::acos(x); }
static double ref_atan(double x) { return std::atan(x); }

int main(int argc, char** argv) {
    const long n = argc > 1 ? std::atol(argv[1]) : 1000000;
    rng_state = argc > 2 ? std::strtoull(argv[2], nullptr, 10) * 0x9E3779B97F4A7C15ull + 1 : 88172645463325252ull;
    const double PIO2 = 1.5707963267948966;
    struct { const char* name; fn1 mine; fn1 ref; } trig[] = {
        {"sin", my_sin, ref_sin}, {"cos", my_cos, ref_cos}, {"sincos.s", my_sc_s, ref_sin}, {"sincos.c", my_sc_c, ref_cos}};
    for (auto& t : trig) {
        run(t.name, "uniform[-4,4]", t.mine, t.ref, n, [] { return 8.0 * uni() - 4.0; });
        run(t.name, "uniform[-200,200]", t.mine, t.ref, n, [] { return 400.0 * uni() - 200.0; });
        run(t.name, "log2[-40,26.6]", t.mine, t.ref, n, [] { return logu(-40, 26.6); });
        run(t.name, "near-k*pi/2", t.mine, t.ref, n, [PIO2] { return (double)((long)(rnd() % 2001) - 1000) * PIO2 * (1.0 + (uni() - 0.5) * 0x1p-30); });
        run(t.name, "range-edges", t.mine, t.ref, n / 4, [] {
            static const double e[] = {0x1p-26, 0x1p-27, 0.126, 0.855469, 2.426265, 105414350.0 * 0.999999};
            return e[rnd() % 6] * (1.0 + (uni() - 0.5) * 0x1p-20) * sgn(); });
    }
    run("acos", "uniform[-1,1]", my_acos, ref_acos, 2 * n, [] { return 2.0 * uni() - 1.0; });
    run("acos", "near+-1", my_acos, ref_acos, n, [] { return sgn() * (1.0 - std::exp2(-53.0 * uni())); });
    run("acos", "near0", my_acos, ref_acos, n, [] { return logu(-60, -2); });
    run("acos", "range-edges", my_acos, ref_acos, n / 4, [] {
        static const double e[] = {0.125, 0.25, 0.5, 0.75, 0.921875, 0.953125, 0.96875, 1.0};
        return e[rnd() % 8] * (1.0 + (uni() - 0.5) * 0x1p-30) * sgn(); });
    run("acos", "specials", my_acos, ref_acos, 64, [] {
        static const double e[] = {0.0, -0.0, 1.0, -1.0, 1.5, -2.0, 0x1p-60, -0x1p-55};
        return e[rnd() % 8]; });
    run("atan", "log2[-40,60]", my_atan, ref_atan, 2 * n, [] { return logu(-40, 60); });
    run("atan", "uniform[-20,20]", my_atan, ref_atan, n, [] { return 40.0 * uni() - 20.0; });
    run("atan", "uniform[-1,1]", my_atan, ref_atan, n, [] { return 2.0 * uni() - 1.0; });
    run("atan", "range-edges", my_atan, ref_atan, n / 4, [] {
        static const double e[] = {0x1.bb67ap-27, 0.0625, 1.0, 16.0, 0x1.49ff2p+52};
        return e[rnd() % 5] * (1.0 + (uni() - 0.5) * 0x1p-30) * sgn(); });
    run("atan", "specials", my_atan, ref_atan, 64, [] {
        static const double e[] = {0.0, -0.0, INFINITY, -INFINITY, 1e300, -1e300, 4.9e-324, NAN};
        return e[rnd() % 8]; });
    run("pow2", "log2[-360,360]", my_pow2, ref_pow2, 2 * n, [] { return logu(-360, 360); });
    run("pow2", "log2[-30,30]", my_pow2, ref_pow2, n, [] { return logu(-30, 30); });
    run("pow2", "near1", my_pow2, ref_pow2, n, [] { return 1.0 + (uni() - 0.5) * std::exp2(-52.0 * uni()); });
    run("pow2", "uniform[-2,2]", my_pow2, ref_pow2, n, [] { return 4.0 * uni() - 2.0; });
    run("pow2", "specials", my_pow2, ref_pow2, 64, [] {
        static const double e[] = {0.0, -0.0, 1.0, -1.0, INFINITY, -INFINITY, 2.0, NAN};
        return e[rnd() % 8]; });
    std::printf("total_mismatches %ld\n", total_bad);
    return total_bad ? 1 : 0;
}
