// acsfit_math.cuh -- the float64 predicates of the path, written once for host and device.
//
// Every expression keeps the reference's operation ORDER:
//   node fit : cap - (used + req) >= 0     autoscaler/kube.py:173-176 (+ :203-213, :247-249)
//   bin fit  : remaining - req    >= 0     autoscaler/scaler.py:134,139-140
// There is no multiply-add pair in them, so no FMA contraction can occur; the one expression
// that could contract (0.3*cap - util, scaler.py:86-87) uses explicit __dmul_rn/__dsub_rn in
// the idle-scan kernel and the library is built with -fmad=false.
//
// Compare-only form.  fits_node(cap, used, r) is monotone non-increasing in r for r >= 0
// (round-to-nearest addition is monotone, subtraction from a constant is antitone), so for a
// node row there is a threshold
//        thr = max { r >= 0 : fits_node(cap, used, r) }      (or -1 when even r = 0 fails)
// with  fits_node(cap, used, r)  <=>  r <= thr  for every r >= 0 that is not NaN.
// node_threshold() finds thr by bisection over the BIT PATTERNS of the non-negative doubles
// (which are ordered like the doubles themselves) calling the literal predicate, so the result
// is exact by construction - no ulp arithmetic to get wrong.  The scan kernels then test
// `req <= thr` (one DSETP per dimension) instead of two DADDs and a DSETP.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define ACSFIT_HD __host__ __device__ __forceinline__
#else
#define ACSFIT_HD inline
#endif

namespace acsfit {

ACSFIT_HD double bits_to_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    std::memcpy(&d, &b, sizeof d);
    return d;
#endif
}

ACSFIT_HD uint64_t double_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    std::memcpy(&b, &d, sizeof b);
    return b;
#endif
}

// KubeNode.can_fit for one resource dimension (kube.py:175): capacity - (used + resources) >= 0
ACSFIT_HD bool fits_node(double cap, double used, double r)
{
#if defined(__CUDA_ARCH__)
    double s = __dadd_rn(used, r);
    double left = __dsub_rn(cap, s);
#else
    volatile double s = used + r;
    volatile double left = cap - s;
#endif
    return left >= 0.0;
}

// (remaining - resources).possible for one dimension (scaler.py:134,139)
ACSFIT_HD bool fits_bin(double remaining, double r)
{
#if defined(__CUDA_ARCH__)
    double left = __dsub_rn(remaining, r);
#else
    volatile double left = remaining - r;
#endif
    return left >= 0.0;
}

constexpr uint64_t kInfBits = 0x7FF0000000000000ull;

// largest r >= 0 (as a double) with fits_node(cap, used, r); -1.0 when none; +inf when all.
ACSFIT_HD double node_threshold(double cap, double used)
{
    if (!fits_node(cap, used, 0.0)) return -1.0;
    if (fits_node(cap, used, bits_to_double(kInfBits))) return bits_to_double(kInfBits);
    // Fast path.  fl(used + r) <= cap  <=>  used + r lies below the midpoint between cap and its
    // upper neighbour (or on it, when cap's last bit is even), so thr sits within a few ulps of
    //     t1 = (cap - used) + (nextup(cap) - cap) / 2.
    // We do NOT trust that arithmetic: the six neighbours of t1 are probed with the literal predicate
    // and the answer is taken only when the boundary fits(x_k) && !fits(x_{k+1}) is actually seen;
    // anything else (specials, huge exponent gaps) falls through to the exact search below.
    {
        const double gap = bits_to_double(double_to_bits(cap) + 1ull) - cap;  // cap >= 0 here (fits(0) held)
        const double t1 = (cap - used) + 0.5 * gap;
        if (t1 > 0.0 && t1 < 1.7e308 && cap > 0.0) {
            const uint64_t b = double_to_bits(t1);
            if (b >= 4ull && b + 3ull < kInfBits) {
                bool f[7];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k = 0; k < 7; ++k) f[k] = fits_node(cap, used, bits_to_double(b - 3ull + (uint64_t)k));
                if (f[0] && !f[6]) {
                    int last = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int k = 1; k < 6; ++k) last = f[k] ? k : last;  // monotone: the last true
                    return bits_to_double(b - 3ull + (uint64_t)last);
                }
            }
        }
    }
    // gallop from the obvious guess cap - used, then bisect.  invariant: fits(lo), !fits(hi).
    double guess = cap - used;
    uint64_t g = (guess >= 0.0) ? double_to_bits(guess) : 0ull;  // NaN -> 0
    if (g > kInfBits) g = kInfBits;
    uint64_t lo, hi;
    if (fits_node(cap, used, bits_to_double(g))) {
        lo = g;
        uint64_t step = 1;
        for (;;) {
            hi = (kInfBits - lo > step) ? lo + step : kInfBits;
            if (!fits_node(cap, used, bits_to_double(hi))) break;
            lo = hi;  // hi == kInfBits cannot fit (checked above), so this terminates
            step <<= 1;
        }
    } else {
        hi = g;
        uint64_t step = 1;
        for (;;) {
            lo = (hi > step) ? hi - step : 0ull;
            if (fits_node(cap, used, bits_to_double(lo))) break;  // lo == 0 fits (checked above)
            hi = lo;
            step <<= 1;
        }
    }
    while (hi - lo > 1) {
        uint64_t mid = lo + ((hi - lo) >> 1);
        if (fits_node(cap, used, bits_to_double(mid))) lo = mid; else hi = mid;
    }
    return bits_to_double(lo);
}

}  // namespace acsfit
