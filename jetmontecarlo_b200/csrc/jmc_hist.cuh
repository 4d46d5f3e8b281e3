// Block-privatised FP64 histograms in shared memory (np.histogram restated) and the min/max merge: device helpers
// shared by the replay kernel (jmc_hist), the FP64 fused kernels (jmc_run, JMC_F64) and the process kernels.
#pragma once
#include <climits>
#include <cstdlib>
#include "jmc_common.cuh"
#include "jmc_physics.cuh"

namespace jmc {

// Shared-memory layout of a privatised histogram block:
//   double edges[n_edges]; double sum_w[nb << shift]; double sum_w2[nb << shift]; uint32 count[nb << shift]
// The weighted sums exist in R = 2^shift lane-interleaved copies, slot [bin][lane & (R-1)]: FP64 shared-memory adds are
// compare-and-swap loops on sm_100 (only 32-bit integer adds are native), and a loop runs one round per lane of the
// warp on the same address.  With one copy per lane (R = 32; chosen by the host as large as shared memory allows) no
// two lanes of a warp ever meet, so a round repeats only when another warp updates the same slot in between.
#ifndef JMC_HIST_MODE
#define JMC_HIST_MODE 1          // 1: lane-interleaved copies; 2: __match_any_sync pre-reduction; 0: one copy
#endif
#ifndef JMC_HIST_UNROLL
#define JMC_HIST_UNROLL 4        // samples per thread per trip (their loads are issued together)
#endif
struct HistSmem {
    double* edges;
    double* sum_w;
    double* sum_w2;
    unsigned int* count;
    int shift;
    __device__ HistSmem(void* base, int n_edges, int shift_) : shift(shift_) {
        const int nb = n_edges - 1;
        edges = (double*)base;
        sum_w = edges + n_edges;
        sum_w2 = sum_w + (nb << shift);
        count = (unsigned int*)(sum_w2 + (nb << shift));      // same slots: [bin][lane & (R-1)]
    }
};
static size_t hist_smem_bytes(int n_edges, int shift = 0) {
    const size_t nb = (size_t)n_edges - 1;
    return sizeof(double) * ((size_t)n_edges + 2 * (nb << shift)) + sizeof(unsigned int) * (nb << shift);
}
static int hist_spread_shift(int n_edges, size_t budget) {
    if (JMC_HIST_MODE != 1) return 0;
    static const int max_shift = [] {                    // experiment switch (scripts/acc_shootout.sh)
        const char* e = getenv("JMC_HIST_SPREAD_SHIFT");
        return e ? atoi(e) : 5;
    }();
    int shift = max_shift;
    while (shift > 0 && hist_smem_bytes(n_edges, shift) > budget) --shift;
    return shift;
}

__device__ __forceinline__ void hist_init(HistSmem& h, const double* __restrict__ edges, int n_edges) {
    const int nb = n_edges - 1;
    for (int k = threadIdx.x; k < n_edges; k += blockDim.x) h.edges[k] = edges[k];
    for (int k = threadIdx.x; k < (nb << h.shift); k += blockDim.x) {
        h.sum_w[k] = 0.0;
        h.sum_w2[k] = 0.0;
        h.count[k] = 0u;
    }
    __syncthreads();
}

// (w, w^2) into bin b of the block histogram
__device__ __forceinline__ void hist_add_weight(HistSmem& h, int b, double w) {
    const unsigned lane = threadIdx.x & 31u;
#if JMC_HIST_MODE == 2
    const unsigned mask = __activemask();
    const unsigned peers = __match_any_sync(mask, b);
    const int leader = __ffs(peers) - 1;
    unsigned others = peers & ~(1u << leader);
    double aw = w, aw2 = w * w;
    const double mw = aw, mw2 = aw2;
    while (__any_sync(mask, others != 0u)) {
        const int src = others ? __ffs(others) - 1 : (int)lane;
        const double vw = __shfl_sync(mask, mw, src), vw2 = __shfl_sync(mask, mw2, src);
        if (others) { aw += vw; aw2 += vw2; }
        others &= others - 1u;
    }
    if ((int)lane == leader) {
        atomicAdd(&h.sum_w[b], aw);
        atomicAdd(&h.sum_w2[b], aw2);
    }
#else
    const int slot = (b << h.shift) | (int)(lane & ((1u << h.shift) - 1u));
    atomicAdd(&h.sum_w[slot], w);
    atomicAdd(&h.sum_w2[slot], w * w);
#endif
}

// One sample into the block histogram.  A NaN weight is not added; it lowers
// the poison bin instead (NumPy: NaN in a cumsum reaches every later bin).
template <int MODE = -1>
__device__ __forceinline__ void hist_add(HistSmem& h, int n_edges, double obs, double w, bool has_w,
                                         int* s_poison, const BinGuess& G = BinGuess{0, 0.f, 0.f, 0.f}) {
    const int b = bin_index_guided<MODE>(obs, h.edges, n_edges, G);
    if (has_w && w != w) {
        // obs below the first edge sorts before every bin -> poisons them all;
        // obs above the last edge or NaN sorts after the last bin -> harmless.
        int pb = b;
        if (b < 0) pb = (obs < h.edges[0]) ? 0 : INT_MAX;
        if (pb != INT_MAX) atomicMin(s_poison, pb);
        if (b >= 0) atomicAdd(&h.count[(b << h.shift) | (int)((threadIdx.x & 31u) & ((1u << h.shift) - 1u))], 1u);
        return;
    }
    if (b < 0) return;
    // (native 32-bit shared atomic on the lane's copy of the bin: no two lanes of a warp on one address, and with 16
    // or 32 copies no two on one bank)
    atomicAdd(&h.count[(b << h.shift) | (int)((threadIdx.x & 31u) & ((1u << h.shift) - 1u))], 1u);
    if (has_w && w != 0.0) hist_add_weight(h, b, w);
}

__device__ __forceinline__ void hist_flush(HistSmem& h, int n_edges, double* __restrict__ g_sum_w,
                                           double* __restrict__ g_sum_w2, unsigned long long* __restrict__ g_count,
                                           int* s_poison, int* g_poison) {
    __syncthreads();
    const int nb = n_edges - 1;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        unsigned int c = 0u;
        for (int r = 0; r < (1 << h.shift); ++r) c += h.count[(k << h.shift) + r];
        if (c) {
            atomicAdd(&g_count[k], (unsigned long long)c);
            double sw = 0.0, sw2 = 0.0;
            for (int r = 0; r < (1 << h.shift); ++r) {
                sw += h.sum_w[(k << h.shift) + r];
                sw2 += h.sum_w2[(k << h.shift) + r];
            }
            if (g_sum_w && sw != 0.0) atomicAdd(&g_sum_w[k], sw);
            if (g_sum_w2 && sw2 != 0.0) atomicAdd(&g_sum_w2[k], sw2);
        }
    }
    if (threadIdx.x == 0 && *s_poison != INT_MAX) atomicMin(g_poison, *s_poison);
}

// Equally spaced edges (np.linspace / np.logspace, i.e. every setBins) get an affine first guess of the bin.  The
// block derives it from the edges it has just staged (no host read-back of the edge array, no stream sync).
__device__ __forceinline__ BinGuess derive_bin_guess(const double* s_edges, int n_edges) {
    BinGuess G{0, 0.f, 0.f, 0.f};
    if (n_edges < 8) return G;                  // (uniform over the block: every thread takes the same path)
    const double e0 = s_edges[0], e1 = s_edges[n_edges - 1];
    const double nb = (double)(n_edges - 1);
    // `tight`: every edge within 1e-6 of a bin of the uniform model -- then the FP32 map decides alone wherever its
    // value is further than eps from an integer; eps = twice the sum of the error bounds of the map
    {
        const double d = (e1 - e0) / nb;
        bool ok = d > 0.0 && isfinite(d) && isfinite(e0), tight = true;
        for (int k = threadIdx.x; ok && k < n_edges; k += blockDim.x) {
            const double dev = fabs(s_edges[k] - (e0 + k * d));
            ok = dev < 0.25 * d;
            tight = tight && dev < 1e-6 * d;
        }
        const int all_tight = __syncthreads_and(tight);
        if (__syncthreads_and(ok)) {
            // (float)x and (float)e0: 2^-24 relative each; the product: 2^-24 relative of a value below nb
            const double eps = 2.0 * (1.2e-7 * fmax(fabs(e0), fabs(e1)) / d + 1.2e-7 * nb + 1e-6);
            return BinGuess{1, (float)e0, (float)(1.0 / d), (all_tight && eps < 0.2) ? (float)eps : 0.f};
        }
    }
    // log-uniform edges, or (first = 1) log-uniform edges behind one zero-bin edge
    for (int first = 0; first <= 1; ++first) {
        const int m = n_edges - first;
        if (m < 8) break;
        const double ea = s_edges[first], mb = (double)(m - 1);
        bool ok = ea > 1e-30 && e1 < 1e30, tight = true;      // the FP32 log2 guess must not underflow
        const double l0 = ok ? log2(ea) : 0.0, d = ok ? (log2(e1) - l0) / mb : 0.0;
        ok = ok && d > 0.0 && isfinite(d);
        for (int k = threadIdx.x; ok && k < m; k += blockDim.x) {
            const double dev = fabs(log2(s_edges[first + k]) - (l0 + k * d));
            ok = dev < 0.25 * d;
            tight = tight && dev < 1e-6 * d;
        }
        const int all_tight = __syncthreads_and(tight);
        if (__syncthreads_and(ok)) {
            // __log2f: 2 ulp of the result outside [0.5, 2] (CUDA math API) -> 2 * 2^-23 * |log2 x|, at least 2^-21.4;
            // (float)x: 2^-24 / ln 2; (float)l0: 2^-24 |l0|; the product: 2^-24 relative of a value below nb
            const double span = fmax(fmax(fabs(l0), fabs(log2(e1))), 1.0);
            const double eps = 2.0 * ((2.4e-7 * span + 4e-7 + 1e-7 + 6e-8 * span) / d + 1.2e-7 * mb + 1e-6);
            return BinGuess{first ? 3 : 2, (float)l0, (float)(1.0 / d), (all_tight && eps < 0.2) ? (float)eps : 0.f};
        }
    }
    return G;
}

// ---- min / max with NaN propagation --------------------------------------
// staging: double mn, double mx, int has_nan
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long old = __double_as_longlong(*addr);
    while (v < __longlong_as_double(old)) {
        const unsigned long long assumed = old;
        old = atomicCAS((unsigned long long*)addr, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long old = __double_as_longlong(*addr);
    while (v > __longlong_as_double(old)) {
        const unsigned long long assumed = old;
        old = atomicCAS((unsigned long long*)addr, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// block-reduce per-thread (mn, mx, nan) and merge into d_minmax[0..1]; a NaN
// anywhere makes both NaN (np.min / np.max semantics)
__device__ __forceinline__ void minmax_merge(double mn, double mx, bool has_nan, double* __restrict__ d_minmax) {
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    has_nan = __any_sync(0xffffffffu, has_nan);
    if ((threadIdx.x & 31) == 0) {
        if (has_nan) {
            // NaN is sticky: atomic_min/max never replace it because comparisons fail
            d_minmax[0] = qnan();
            d_minmax[1] = qnan();
        } else {
            atomic_min_double(&d_minmax[0], mn);
            atomic_max_double(&d_minmax[1], mx);
        }
    }
}


// host: argument checks shared by the histogram entry points; *smem = bytes of one accumulator copy per bin
static inline int check_hist_args(const char* who, const double* d_edges, int n_edges, size_t* smem) {
    JMC_REQUIRE(n_edges >= 2, "%s: need at least 2 edges", who);
    JMC_REQUIRE(d_edges != nullptr, "%s: edges is NULL", who);
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    *smem = hist_smem_bytes(n_edges);
    if (*smem > (size_t)di.max_smem_optin)
        return set_error(JMC_ELIMIT, "%s: %d edges need %zu B of shared memory (limit %d B, about 8100 edges)", who,
                         n_edges, *smem, di.max_smem_optin);
    return JMC_OK;
}


}  // namespace jmc
