// host-side property test of acsfit::node_threshold (csrc/acsfit_math.cuh):
//   thr is exact  <=>  fits(thr) && !fits(nextup(thr))   (given monotonicity of fits in r)
// and the compare-only predicate r <= thr equals the literal predicate for probes around thr.
// Build: g++ -O2 -ffp-contract=off -I<csrc> test_threshold.cpp -o test_threshold
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "acsfit_math.cuh"

using namespace acsfit;

static long failures = 0;

static void check(double cap, double used)
{
    double thr = node_threshold(cap, used);
    if (thr < 0) {
        if (fits_node(cap, used, 0.0)) { ++failures; std::printf("FAIL none cap=%a used=%a\n", cap, used); }
        return;
    }
    if (!fits_node(cap, used, thr)) { ++failures; std::printf("FAIL !fits(thr) cap=%a used=%a thr=%a\n", cap, used, thr); return; }
    if (std::isinf(thr)) return;
    double up = std::nextafter(thr, INFINITY);
    if (fits_node(cap, used, up)) { ++failures; std::printf("FAIL fits(up) cap=%a used=%a thr=%a\n", cap, used, thr); return; }
    // probes
    double probes[] = {0.0, thr, up, std::nextafter(thr, 0.0), thr * 0.5, thr * 2.0, thr + 1.0, 1e-300, 4.9e-324, INFINITY};
    for (double r : probes) {
        if (!(r >= 0)) continue;
        bool lit = fits_node(cap, used, r);
        bool cmp = r <= thr;
        if (lit != cmp) { ++failures; std::printf("FAIL probe cap=%a used=%a r=%a lit=%d cmp=%d\n", cap, used, r, lit, cmp); }
    }
}

int main(int argc, char **argv)
{
    long n = argc > 1 ? std::atol(argv[1]) : 2000000;
    std::mt19937_64 rng(20260921);
    const double caps[] = {1, 2, 4, 6, 8, 16, 110, 7096762368.0, 59087724544.0, 0, 2145336164352.0, 0.3, 1e-3, 1e300, 5e-324};
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    for (long i = 0; i < n; ++i) {
        double cap = caps[rng() % (sizeof caps / sizeof caps[0])];
        if (rng() % 4 == 0) cap = std::ldexp(u01(rng), (int)(rng() % 80) - 20);
        double used;
        switch (rng() % 6) {
        case 0: used = 0; break;
        case 1: used = cap; break;
        case 2: used = std::nextafter(cap, 0.0); break;
        case 3: used = cap * u01(rng); break;
        case 4: { // sum of millicore-like quantities
            used = 0; int k = rng() % 40; for (int j = 0; j < k; ++j) used = used + (double)(rng() % 4000) * 1e-3; break; }
        default: used = cap * 1.5 * u01(rng); break;
        }
        check(cap, used);
    }
    // specials
    const double sp[] = {0.0, 1.0, INFINITY, NAN, 1e308, 5e-324, 2.0, 1.7976931348623157e308};
    for (double c : sp) for (double u : sp) check(c, u);
    std::printf("checked %ld random pairs, failures=%ld\n", n, failures);
    return failures ? 1 : 0;
}
