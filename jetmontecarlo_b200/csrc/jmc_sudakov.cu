// Per-event inverse-transform sampling from a conditional cdf exp(-R(x, cond_i)), R a numerical radiator tabulated on
// a rectangular grid -- the stage between the numerical radiators and the final pdf in the reference's production
// workflow (examples/sudakov_comparisons/sudakov_utils.py:134-238: one samples_from_cdf call per event, i.e. per
// event ~1000 probe points, an interpolant evaluation at each, and a walk through a list of known cdf shapes,
// jetmontecarlo/utils/montecarlo_utils.py:142-318, then scipy's interp1d inversion, :377-399).
//
// Device design: ONE EVENT PER BLOCK.  The event's probe grid, its tabulated cdf and the sort keys live in shared
// memory (24 KB); the radiator table is read through L2 (it is shared by all events).  Every step of the
// reference's per-event walk is a block-wide reduction over the 1000-odd points; an event whose domain has to be cut
// (negligible head, turning point) loops inside its block, exactly as the reference recurses.  Nothing per event
// touches HBM except its condition, its uniform and its sample.  FP64 in the reference's operation order.
// compiled with --fmad=false
#include <cfloat>
#include <cmath>
#include "jmc_common.cuh"
#include "jmc_philox.cuh"
#include "jmc_physics.cuh"

namespace jmc {

static constexpr int kSudThreads = 256;
static constexpr int kSudPoints = 1024;          // >= 1001 probe points, power of two for the bitonic sort
static constexpr double kCdfThreshold = 1e-20;   // "negligible cdf", utils/montecarlo_utils.py:237

struct SudTable {
    const double* gx;
    const double* gy;
    const double* z;       // [nx, ny] row-major
    int nx, ny, mode;      // mode 0: bilinear, linearly extrapolated (RectangularGrid); 1: nearest grid node
};

// index i of the cell [g_i, g_{i+1}] used for v: searchsorted(g, v, 'right') - 1 clamped to [0, n-2]
__device__ __forceinline__ int sud_cell(const double* __restrict__ g, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return max(0, min(lo - 1, n - 2));
}
__device__ __forceinline__ int sud_nearest(const double* __restrict__ g, int n, double v) {
    const int i = sud_cell(g, n, v);
    return (fabs(v - g[i]) <= fabs(g[i + 1] - v)) ? i : i + 1;
}

// R(x, y): what get_2d_interpolation(xs, ys, zs) returns (utils/interpolation.py:248-273)
__device__ __forceinline__ double sud_radiator(const SudTable& T, double x, double y) {
    if (T.mode == 1) return T.z[(size_t)sud_nearest(T.gx, T.nx, x) * T.ny + sud_nearest(T.gy, T.ny, y)];
    const int i = sud_cell(T.gx, T.nx, x), j = sud_cell(T.gy, T.ny, y);
    const double tx = (x - T.gx[i]) / (T.gx[i + 1] - T.gx[i]);
    const double ty = (y - T.gy[j]) / (T.gy[j + 1] - T.gy[j]);
    const double* r0 = T.z + (size_t)i * T.ny + j;
    const double* r1 = r0 + T.ny;
    return r0[0] * (1 - tx) * (1 - ty) + r0[1] * (1 - tx) * ty + r1[0] * tx * (1 - ty) + r1[1] * tx * ty;
}

// np.nan_to_num defaults
__device__ __forceinline__ double sud_nan_to_num(double v) {
    if (v != v) return 0.0;
    if (isinf(v)) return v > 0 ? DBL_MAX : -DBL_MAX;
    return v;
}

// np.linspace(start, stop, num)[k]
__device__ __forceinline__ double sud_linspace(double start, double stop, int num, int k) {
    if (k == num - 1) return stop;
    return k * ((stop - start) / (num - 1)) + start;
}

// block-wide reductions over a predicate / an index (all threads call; results are uniform)
__device__ __forceinline__ int block_min_index(int v, int* s_red) {
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = INT_MAX;
    for (int w = 0; w < kSudThreads / 32; ++w) r = min(r, s_red[w]);
    return r;
}

enum { SUD_OK = 0, SUD_NOT_MONOTONE = 1, SUD_ROUNDS = 2, SUD_SHORT_TABLE = 3 };

// What an event's domain, uniform and result formatting are.  kind 0: all given by the caller's arrays (the general
// entry point).  kinds 1 / 2: the generators of the production workflow, complete in the kernel --
//   1 pre-critical: domain [0, param] (param = z_cut);
//   2 subsequent:   domain [0, cond^param / 2] (param = beta); an event whose upper end is below 1e-10 takes the
//                   underflow value 1e-100 without sampling (sudakov_utils.py:203-211);
// the uniform of event e is uniforms[e] if given, else the first 53-bit Philox uniform of stream (seed, e); an
// infinite sample becomes 0 with weight 0 (sudakov_utils.py:117-120), every other weight is 1.
struct SudGen {
    int kind;
    double param;
    uint64_t seed;
    double* weights;
};

__global__ void __launch_bounds__(kSudThreads)
conditional_inverse_kernel(SudTable T, const double* __restrict__ cond, const double* __restrict__ dom_lo,
                           const double* __restrict__ dom_hi, const double* __restrict__ uniforms, uint64_t n,
                           int catch_turnpoint, int force_monotone, int max_rounds, double* __restrict__ samples,
                           int* __restrict__ status, SudGen gen) {
    __shared__ double s_x[kSudPoints], s_c[kSudPoints];      // probe points / cdf values (then: the sorted table)
    __shared__ double s_aux[kSudPoints];                      // merge scratch: the two halves of the mixed grid
    __shared__ int s_key[kSudPoints];                         // original position (stable tie-break)
    __shared__ int s_red[kSudThreads / 32];
    const int tid = threadIdx.x;
    for (uint64_t ev = blockIdx.x; ev < n; ev += gridDim.x) {
        const double c = cond[ev];
        double lo = gen.kind == 0 ? dom_lo[ev] : 0.0;
        double hi = gen.kind == 0 ? dom_hi[ev] : (gen.kind == 1 ? gen.param : np_pow(c, gen.param) / 2.);
        double r;
        if (uniforms) {
            r = uniforms[ev];
        } else {
            double unused;
            philox_uniform53x2(ev, 0u, gen.seed, r, unused);
        }
        int result_status = SUD_ROUNDS;
        double result = nan("");
        // (NumPy's `upper < 1e-10` is false for NaN: such an event is sampled, as in the reference)
        const bool underflow = gen.kind == 2 && hi < 1e-10;
        if (underflow) { result = 1e-100; result_status = SUD_OK; }
        for (int round = 0; round < max_rounds && !underflow; ++round) {
            // ---- probe points (utils/montecarlo_utils.py:147-154, utils/interpolation.py:18-41)
            int M;
            __syncthreads();
            if (lo < 0) {
                M = 1000;
                for (int k = tid; k < M; k += kSudThreads) s_x[k] = sud_linspace(lo, hi, 1000, k);
            } else {
                const bool from_zero = lo == 0;
                const double l = from_zero ? lo + 1e-20 : lo;
                const int off = from_zero ? 1 : 0;
                M = 1000 + off;
                // lin_log_mixed_list(l, hi, 1000): the logarithmic half without its first point and the linear half
                // without its last point, merged in order.  s_aux[0..500] = log part, s_aux[512..1010] = linear part
                const double ll = log10(l), lh = log10(hi);
                for (int k = tid; k < 501; k += kSudThreads) s_aux[k] = pow(10., sud_linspace(ll, lh, 502, k + 1));
                for (int k = tid; k < 499; k += kSudThreads) s_aux[512 + k] = sud_linspace(l, hi, 500, k);
                __syncthreads();
                // rank of every element in the merged sequence = own index + elements of the other list before it
                for (int k = tid; k < 501; k += kSudThreads) {
                    const double v = s_aux[k];
                    int a = 0, b = 499;
                    while (a < b) { const int m = (a + b) >> 1; if (s_aux[512 + m] < v) a = m + 1; else b = m; }
                    s_x[off + k + a] = v;
                }
                for (int k = tid; k < 499; k += kSudThreads) {
                    const double v = s_aux[512 + k];
                    int a = 0, b = 501;
                    while (a < b) { const int m = (a + b) >> 1; if (s_aux[m] <= v) a = m + 1; else b = m; }
                    s_x[off + k + a] = v;
                }
                if (from_zero && tid == 0) s_x[0] = 0.0;
            }
            __syncthreads();
            // ---- the conditional cdf on the probe points: raw = exp(-R(x, cond)), vals = nan_to_num(raw)
            for (int k = tid; k < M; k += kSudThreads) s_c[k] = sud_nan_to_num(exp(-1. * sud_radiator(T, s_x[k], c)));
            __syncthreads();
            // ---- shape of the table (utils/montecarlo_utils.py:157-300)
            int first_bad = INT_MAX, first_above = INT_MAX, not_one = 0, not_negligible = 0;
            for (int k = tid; k < M; k += kSudThreads) {
                const double v = s_c[k];
                if (k < M - 1 && !(v <= s_c[k + 1])) first_bad = min(first_bad, k);
                if (v > kCdfThreshold) first_above = min(first_above, k);
                not_one |= !(v == 1.);
                not_negligible |= !(v < kCdfThreshold);
            }
            first_bad = block_min_index(first_bad, s_red);
            first_above = block_min_index(first_above, s_red);
            const bool always_one = !__syncthreads_or(not_one), negligible = !__syncthreads_or(not_negligible);
            const bool all_mono = first_bad == INT_MAX;
            // is the table monotone from the first non-negligible point on?
            int bad_after = 0;
            if (!all_mono && first_above != INT_MAX)
                for (int k = tid; k < M - 1; k += kSudThreads) bad_after |= (k >= first_above && !(s_c[k] <= s_c[k + 1]));
            const bool monotone_after = first_above != INT_MAX && !__syncthreads_or(bad_after);
            // the raw cdf where the table first decreases (the reference re-evaluates the cdf there)
            double raw_bad = 0.0;
            if (!all_mono) raw_bad = exp(-1. * sud_radiator(T, s_x[first_bad], c));
            const bool starts_at_one = !all_mono && first_bad == 0 && raw_bad == 1.;
            int n_keep = 0;
            if (all_mono) {
                n_keep = M;                                   // a monotone table is already sorted by cdf value
            } else if (always_one || starts_at_one) {
                result = 0.0; result_status = SUD_OK; break;  // zero bin
            } else if (negligible) {
                result = hi; result_status = SUD_OK; break;   // overflow bin
            } else if (monotone_after) {
                lo = s_x[first_above];                        // drop the negligible head and try again
                continue;
            } else if (raw_bad == 1. && catch_turnpoint) {
                hi = s_x[first_bad];                          // valid up to the turning point only
                continue;
            } else if (force_monotone) {
                // keep the entries after which the cdf does not decrease (the last entry is dropped), sort them by
                // cdf value -- stable -- as interp1d does
                __syncthreads();
                double v_keep[kSudPoints / kSudThreads], x_keep[kSudPoints / kSudThreads];
                bool keep[kSudPoints / kSudThreads];
#pragma unroll
                for (int q = 0; q < kSudPoints / kSudThreads; ++q) {
                    const int k = tid + q * kSudThreads;
                    keep[q] = k < M - 1 && s_c[k] <= s_c[k + 1];
                    v_keep[q] = k < M ? s_c[k] : 0.0;
                    x_keep[q] = k < M ? s_x[k] : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < kSudPoints / kSudThreads; ++q) {
                    const int k = tid + q * kSudThreads;
                    s_c[k] = keep[q] ? v_keep[q] : INFINITY;  // dropped entries sort behind every kept one
                    s_x[k] = x_keep[q];
                    s_key[k] = keep[q] ? k : kSudPoints + k;
                }
                {
                    int cnt = 0;
#pragma unroll
                    for (int q = 0; q < kSudPoints / kSudThreads; ++q) cnt += keep[q];
                    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                    if ((tid & 31) == 0) s_red[tid >> 5] = cnt;
                    __syncthreads();
                    n_keep = 0;
                    for (int w = 0; w < kSudThreads / 32; ++w) n_keep += s_red[w];
                    __syncthreads();
                }
                // bitonic sort of (cdf value, original position) with the x values riding along
                for (int size = 2; size <= kSudPoints; size <<= 1) {
                    for (int st = size >> 1; st > 0; st >>= 1) {
                        for (int k = tid; k < kSudPoints; k += kSudThreads) {
                            const int p = k ^ st;
                            if (p > k) {
                                const bool up = (k & size) == 0;
                                const double ck = s_c[k], cp = s_c[p];
                                const int kk = s_key[k], kp = s_key[p];
                                const bool gt = ck > cp || (ck == cp && kk > kp);
                                if (gt == up) {
                                    s_c[k] = cp; s_c[p] = ck;
                                    s_key[k] = kp; s_key[p] = kk;
                                    const double t = s_x[k]; s_x[k] = s_x[p]; s_x[p] = t;
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
            } else {
                result_status = SUD_NOT_MONOTONE; break;
            }
            // ---- x(r) from the table sorted by cdf value: linear between the neighbours of r, linear beyond the ends
            // (scipy interp1d(cdf, x, fill_value='extrapolate'), utils/montecarlo_utils.py:377-399)
            if (n_keep < 2) { result_status = SUD_SHORT_TABLE; break; }
            if (tid == 0) {
                int a = 0, b = n_keep;                        // searchsorted(c, r) 'left'
                while (a < b) { const int m = (a + b) >> 1; if (s_c[m] < r) a = m + 1; else b = m; }
                const int idx = min(max(a, 1), n_keep - 1);
                const double c_lo = s_c[idx - 1], c_hi = s_c[idx], x_lo = s_x[idx - 1], x_hi = s_x[idx];
                // the two-weight form of current SciPy's interp1d
                s_aux[0] = ((r - c_lo) / (c_hi - c_lo)) * x_hi + ((c_hi - r) / (c_hi - c_lo)) * x_lo;
            }
            __syncthreads();
            result = s_aux[0];
            result_status = SUD_OK;
            break;
        }
        if (tid == 0) {
            if (gen.kind != 0) {
                const bool lost = isinf(result);
                samples[ev] = lost ? 0.0 : result;
                if (gen.weights) gen.weights[ev] = lost ? 0.0 : 1.0;
            } else {
                samples[ev] = result;
                if (gen.weights) gen.weights[ev] = 1.0;
            }
            if (result_status != SUD_OK) atomicMax(status, result_status);
        }
        __syncthreads();
    }
}

}  // namespace jmc

using namespace jmc;

static int launch_conditional(const double* d_grid_x, int nx, const double* d_grid_y, int ny, const double* d_table,
                              int interpolation, const double* d_cond, const double* d_dom_lo, const double* d_dom_hi,
                              const double* d_uniforms, uint64_t n, int catch_turnpoint, int force_monotone,
                              double* d_samples, SudGen gen, void* stream) {
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    cudaStream_t st = (cudaStream_t)stream;
    int* d_status;
    int rc = scratch(sizeof(int), (void**)&d_status, stream);
    if (rc) return rc;
    JMC_CUDA(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    SudTable T{d_grid_x, d_grid_y, d_table, nx, ny, interpolation};
    const uint64_t cap = (uint64_t)di.sm_count * 8;
    const int grid = (int)(n < cap ? n : cap);
    conditional_inverse_kernel<<<grid, kSudThreads, 0, st>>>(T, d_cond, d_dom_lo, d_dom_hi, d_uniforms, n, catch_turnpoint,
                                                             force_monotone, 64, d_samples, d_status, gen);
    JMC_LAUNCHED("conditional_inverse_kernel");
    int h_status = 0;
    JMC_CUDA(cudaMemcpyAsync(&h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    JMC_CUDA(cudaStreamSynchronize(st));
    if (h_status == SUD_NOT_MONOTONE)
        return set_error(JMC_EDOMAIN, "conditional_samples_from_cdf: the cdf of an event is not monotone on its domain and "
                                      "neither catch_turnpoint nor force_monotone resolves it");
    if (h_status == SUD_ROUNDS)
        return set_error(JMC_EDOMAIN, "conditional_samples_from_cdf: domains still shrinking after 64 rounds");
    if (h_status == SUD_SHORT_TABLE)
        return set_error(JMC_EDOMAIN, "conditional_samples_from_cdf: a cdf table needs at least two points");
    return JMC_OK;
}

static int check_table(const char* who, const double* d_grid_x, int nx, const double* d_grid_y, int ny,
                       const double* d_table, int interpolation) {
    JMC_REQUIRE(nx >= 2 && ny >= 2, "%s: the radiator grid needs at least 2 x 2 points", who);
    JMC_REQUIRE(interpolation == 0 || interpolation == 1, "%s: interpolation must be 0 (RectangularGrid) or 1 (Nearest)", who);
    JMC_REQUIRE(d_grid_x && d_grid_y && d_table, "%s: NULL radiator table", who);
    return JMC_OK;
}

extern "C" int jmc_conditional_inverse_transform(const double* d_grid_x, int nx, const double* d_grid_y, int ny,
                                                 const double* d_table, int interpolation, const double* d_cond,
                                                 const double* d_dom_lo, const double* d_dom_hi,
                                                 const double* d_uniforms, uint64_t n, int catch_turnpoint,
                                                 int force_monotone, double* d_samples, void* stream) {
    int rc = check_table("jmc_conditional_inverse_transform", d_grid_x, nx, d_grid_y, ny, d_table, interpolation);
    if (rc) return rc;
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_cond && d_dom_lo && d_dom_hi && d_uniforms && d_samples, "jmc_conditional_inverse_transform: NULL array");
    return launch_conditional(d_grid_x, nx, d_grid_y, ny, d_table, interpolation, d_cond, d_dom_lo, d_dom_hi, d_uniforms, n,
                              catch_turnpoint, force_monotone, d_samples, SudGen{0, 0.0, 0ull, nullptr}, stream);
}

extern "C" int jmc_sudakov_generate(int emission, const double* d_grid_x, int nx, const double* d_grid_y, int ny,
                                    const double* d_table, int interpolation, const double* d_theta_crit, double param,
                                    const double* d_uniforms, uint64_t seed, uint64_t n, double* d_samples,
                                    double* d_weights, void* stream) {
    int rc = check_table("jmc_sudakov_generate", d_grid_x, nx, d_grid_y, ny, d_table, interpolation);
    if (rc) return rc;
    JMC_REQUIRE(emission == JMC_SUDAKOV_PRECRITICAL || emission == JMC_SUDAKOV_SUBSEQUENT,
                "jmc_sudakov_generate: emission must be JMC_SUDAKOV_PRECRITICAL or JMC_SUDAKOV_SUBSEQUENT");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_theta_crit && d_samples && d_weights, "jmc_sudakov_generate: NULL array");
    return launch_conditional(d_grid_x, nx, d_grid_y, ny, d_table, interpolation, d_theta_crit, nullptr, nullptr, d_uniforms,
                              n, 0, 1, d_samples, SudGen{emission, param, seed, d_weights}, stream);
}
