// acsfit_rank.cuh -- builds the per-tick rank tables of the packed-rank scan (RankLayout, acsfit_kernels.cuh).
//
// For every resource dimension the DISTINCT request values of the tick's pod table are collected (a lock-free
// open-addressing hash set on the float64 bit patterns), sorted (one CTA per dimension, bitonic sort in shared
// memory; non-negative doubles order like their bit patterns) and every pod row is rewritten as NW 32-bit words
// of rank fields.  Real clusters request a few dozen distinct quantities per resource, so the tables are tiny;
// when a dimension has more than kRankCap distinct values, or the fields do not fit 128 bits, the caller falls
// back to the float64 compare scan -- results are identical either way.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "acsfit_kernels.cuh"

namespace acsfit {

constexpr int kRankProbes = 256;
constexpr unsigned long long kRankEmpty = 0xFFFFFFFFFFFFFFFFull;  // a NaN pattern: never a request (domain check)

__device__ __forceinline__ unsigned rank_hash(unsigned long long b)
{
    b ^= b >> 33;
    b *= 0xff51afd7ed558ccdull;
    b ^= b >> 29;
    return (unsigned)b & (unsigned)(kRankSlots - 1);
}
__device__ __forceinline__ unsigned long long rank_key(double v)
{
    return v == 0.0 ? 0ull : (unsigned long long)__double_as_longlong(v);  // -0.0 and +0.0 are one value
}

// keys [D][kRankSlots] preset to kRankEmpty; counts [D] zeroed.  counts[d] may overshoot kRankCap: overflow.
__global__ void rank_insert_kernel(const double *__restrict__ req, int64_t rows, int D,
                                   unsigned long long *keys, int *counts)
{
    const int64_t n = rows * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const unsigned long long k = rank_key(req[i]);
        unsigned long long *tab = keys + (size_t)d * kRankSlots;
        unsigned h = rank_hash(k);
        if (*(volatile int *)(counts + d) > kRankCap) continue;  // already overflowed: the table will not be used
        int probe = 0;
        for (; probe < kRankProbes; ++probe, h = (h + 1) & (kRankSlots - 1)) {
            unsigned long long cur = tab[h];
            if (cur == k) break;
            if (cur == kRankEmpty) {
                cur = atomicCAS(tab + h, kRankEmpty, k);
                if (cur == kRankEmpty) {
                    atomicAdd(counts + d, 1);
                    break;
                }
                if (cur == k) break;
            }
        }
        // a long probe sequence means the set is far beyond kRankCap (load <= 1/4 otherwise): give up, bounded work
        if (probe == kRankProbes) atomicMax(counts + d, kRankCap + 1);
    }
}

// one CTA per dimension: gather the set, sort it, write sorted[d][0..U) and each key's rank (1-based) next to it
__global__ void __launch_bounds__(1024)
rank_sort_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ counts, double *sorted,
                 unsigned short *ranks /*[D][kRankSlots]*/)
{
    extern __shared__ unsigned long long sk[];  // kRankCap keys
    __shared__ int fill;
    const int d = blockIdx.x;
    const int U = counts[d];
    if (U > kRankCap) return;  // overflow: the host sees counts[d] and does not use the table
    const unsigned long long *tab = keys + (size_t)d * kRankSlots;
    if (threadIdx.x == 0) fill = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < kRankSlots; i += blockDim.x) {
        const unsigned long long k = tab[i];
        if (k != kRankEmpty) sk[atomicAdd(&fill, 1)] = k;
    }
    __syncthreads();
    int n2 = 1;
    while (n2 < U) n2 <<= 1;
    for (int i = U + threadIdx.x; i < n2; i += blockDim.x) sk[i] = kRankEmpty;  // pads sort to the end
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const unsigned long long a = sk[i], b = sk[j];
                    const bool up = (i & size) == 0;
                    if ((a > b) == up) {
                        sk[i] = b;
                        sk[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < U; i += blockDim.x) {
        const unsigned long long k = sk[i];
        sorted[(size_t)d * kRankCap + i] = __longlong_as_double((long long)k);
        unsigned h = rank_hash(k);
        while (tab[h] != k) h = (h + 1) & (kRankSlots - 1);
        ranks[(size_t)d * kRankSlots + h] = (unsigned short)(i + 1);
    }
}

struct RankFields {
    uint8_t word[16];
    uint8_t shift[16];
};

// packed[row][nw]: field d of a row = 1 + rank of req[row][d] among the sorted distinct values of dimension d
__global__ void rank_pack_kernel(const double *__restrict__ req, int64_t rows, int D,
                                 const unsigned long long *__restrict__ keys, const unsigned short *__restrict__ ranks,
                                 RankFields f, int nw, uint32_t *__restrict__ packed)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int d = 0; d < D; ++d) {
            const unsigned long long k = rank_key(req[(size_t)r * D + d]);
            const unsigned long long *tab = keys + (size_t)d * kRankSlots;
            unsigned h = rank_hash(k);
            while (tab[h] != k) h = (h + 1) & (kRankSlots - 1);  // present by construction
            w[f.word[d]] |= (uint32_t)ranks[(size_t)d * kRankSlots + h] << f.shift[d];
        }
        for (int i = 0; i < nw; ++i) packed[(size_t)r * nw + i] = w[i];
    }
}

}  // namespace acsfit
