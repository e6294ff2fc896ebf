// Accuracy of csrc/fastmath.cuh against glibc (host build of the same source): max error in ulp.
// g++ -O2 -std=c++17 -I elisa_b200/csrc tools/fastmath_check.cc -o /tmp/fastmath_check && /tmp/fastmath_check
#include <cstdio>
#include <random>
#include "fastmath.cuh"

static double ulp_err(double got, double ref) {
    if (got == ref) return 0.0;
    const double u = std::ldexp(1.0, std::ilogb(ref) - 52);
    return std::fabs(got - ref) / u;
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 20000000;
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> ue(-699.0, 699.0), us(-40.0, 40.0), ul(-300.0, 300.0), um(0.25, 5.0);
    double we = 0, wl = 0, wm = 0, we_at = 0, wl_at = 0;
    for (long i = 0; i < n; ++i) {
        const double x = (i & 1) ? ue(g) : us(g);
        const double ee = ulp_err(elisa::fm::exp(x), (double)std::exp((long double)x));
        if (ee > we) { we = ee; we_at = x; }
        double y = std::exp(ul(g));
        if ((i & 3) == 0) y = um(g);
        const double le = ulp_err(elisa::fm::log(y), (double)std::log((long double)y));
        if (le > wl) { wl = le; wl_at = y; }
        const double z = ((i & 1) ? 1.0 : -1.0) * um(g) * ((i & 2) ? 1.0 : 8.0);
        const double me = ulp_err(elisa::fm::expm1(z), (double)std::expm1((long double)z));
        if (me > wm) wm = me;
    }
    std::printf("exp  max %.3f ulp (at %.17g)\nlog  max %.3f ulp (at %.17g)\nexpm1 max %.3f ulp\n", we, we_at, wl, wl_at, wm);
    // specials
    std::printf("exp(-800)=%g exp(800)=%g exp(nan)=%g log(0)=%g log(-1)=%g log(inf)=%g log(1e-310)=%g log(1)=%g expm1(1e-9)=%.17g\n",
                elisa::fm::exp(-800.0), elisa::fm::exp(800.0), elisa::fm::exp(std::nan("")), elisa::fm::log(0.0),
                elisa::fm::log(-1.0), elisa::fm::log(INFINITY), elisa::fm::log(1e-310), elisa::fm::log(1.0), elisa::fm::expm1(1e-9));
    return (we <= 1.5 && wl <= 1.5 && wm <= 4.5) ? 0 : 1;
}
