// Generic Monte Carlo process: hit-or-miss sampling of a box of kinematic variables with a phase-space constraint,
// weight = jacobian x differential cross section, and the pairwise observables / energy-weighted correlators built
// on it.  What it replaces: MonteCarloProcess.addSample in a Python loop (jetmontecarlo/montecarlo/process.py:109-191),
// the one concrete process of the reference, e+ e- -> hadrons at LO (examples/processes/ee2hadrons_LO.py), its
// pairwise observables (:75-150) and Npoint_MonteCarloEWOC (jetmontecarlo/montecarlo/ewoc.py:11-75).
//
// Device design: sample i owns the Philox stream of index i; trial t of the hit-or-miss loop is block t of that
// stream (two 53-bit uniforms), so a sample and its number of rejected trials are pure functions of (seed, i) and any
// split of the job over launches or GPUs draws the same points.  The fused kernel never stores a sample: it forms the
// three parton pairs of the event, their observable and energy weights, and bins them (FP64, NumPy operation order,
// the block histogram of jmc_hist.cuh).  compiled with --fmad=false
#include <cmath>
#include <cstring>
#include "jmc_common.cuh"
#include "jmc_hist.cuh"
#include "jmc_philox.cuh"
#include "jmc_physics.cuh"

namespace jmc {

static constexpr int kProcThreads = 256;
static constexpr uint32_t kMaxTrials = 1u << 20;      // a box whose acceptance is below 1e-6 is a caller error

struct ProcParams {
    int kind, is_log;
    double log_eps, lo0, lo1, width0, width1, log_width0, log_width1;
    double prefactor;
};

// ref: utils/montecarlo_utils.py:12-55 as called at montecarlo/process.py:124-130
__device__ __forceinline__ double proc_coordinate(const ProcParams& p, double lo, double width, double log_width,
                                                  double u) {
    return p.is_log ? lo + exp(p.log_eps * u + log_width) : lo + u * width;
}

// e+ e- -> q qbar g: 0 < x_i < 1 for x_1 = 1 - a, x_2 = 1 - b, x_3 = 2 - x_1 - x_2.
// ref: examples/processes/ee2hadrons_LO.py:17-21
__device__ __forceinline__ bool ee2h_in_phasespace(double a, double b) {
    const double x_1 = 1 - a, x_2 = 1 - b;
    const double x_3 = 2 - x_1 - x_2;
    return 0 < x_1 && x_1 < 1 && 0 < x_2 && x_2 < 1 && 0 < x_3 && x_3 < 1;
}
// The reference's expression verbatim: its second variable is the kinematic variable 1 - x_2 itself (the "1 -" is
// missing at examples/processes/ee2hadrons_LO.py:28); a deterministic function of the inputs, reproduced.
__device__ __forceinline__ double ee2h_xsec(double a, double b, double prefactor) {
    const double x_1 = 1 - a, x_2 = b;
    return prefactor * (x_1 * x_1 + x_2 * x_2) / ((1. - x_1) * (1. - x_2));
}

// One accepted phase-space point of stream `idx`: coordinates, weight = jacobian * d(sigma), rejected trials.
__device__ __forceinline__ bool proc_draw(const ProcParams& p, uint64_t idx, uint64_t seed, double& a, double& b,
                                          double& w, uint32_t& rejected) {
    for (uint32_t t = 0; t < kMaxTrials; ++t) {
        double u0, u1;
        philox_uniform53x2(idx, t, seed, u0, u1);
        a = proc_coordinate(p, p.lo0, p.width0, p.log_width0, u0);
        b = proc_coordinate(p, p.lo1, p.width1, p.log_width1, u1);
        if (ee2h_in_phasespace(a, b)) {
            double jac = 1.;
            if (p.is_log) { jac *= a; jac *= b; }
            w = jac * ee2h_xsec(a, b, p.prefactor);
            rejected = t;
            return true;
        }
    }
    rejected = kMaxTrials;
    return false;
}

// x ** w as NumPy evaluates it for the usual small exponents (1 and 2 are exact; np.power squares by multiplication)
__device__ __forceinline__ double small_pow(double x, double w) {
    if (w == 1.0) return x;
    if (w == 2.0) return x * x;
    if (w == 0.0) return 1.0;
    return pow(x, w);
}

// pairwise observable of partons with energy fractions (x_i, x_j).  ref: examples/processes/ee2hadrons_LO.py:95-150
__device__ __forceinline__ double pair_observable(int id, double x_i, double x_j, double kappa) {
    const double cos_ij = 2. / (x_i * x_j) - 2. / x_i - 2. / x_j + 1;
    switch (id) {
        case JMC_PAIR_COSTHETA: return cos_ij;
        case JMC_PAIR_THETA: return acos(cos_ij);
        case JMC_PAIR_M2: return -(1 - x_i - x_j);
        case JMC_PAIR_GECF: return pow(2.0, kappa / 2 - 2) * x_i * x_j * pow(1 - cos_ij, kappa / 2);
        case JMC_PAIR_Z_I: return x_i / 2;
        default: return x_j / 2;
    }
}

// block-reduced count of rejected trials
__device__ __forceinline__ void add_rejected(unsigned long long local, unsigned long long* g_rejected) {
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(g_rejected, local);
}

__global__ void __launch_bounds__(kProcThreads)
process_draw_kernel(ProcParams p, uint64_t n, uint64_t seed, uint64_t first, double* __restrict__ samples,
                      double* __restrict__ weights, unsigned long long* g_rejected, int* g_error) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long rej = 0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a, b, w;
        uint32_t r;
        if (!proc_draw(p, first + i, seed, a, b, w, r)) atomicExch(g_error, 1);
        rej += r;
        if (samples) { samples[2 * i] = a; samples[2 * i + 1] = b; }
        if (weights) weights[i] = w;
    }
    add_rejected(rej, g_rejected);
}

// the three parton pairs of every event, pair-major: [pair (1,2) of all events | pair (2,3) | pair (3,1)], as
// ee2hLO_PairwiseObservable.update concatenates them (examples/processes/ee2hadrons_LO.py:80-92)
__global__ void __launch_bounds__(kProcThreads)
process_pairs_kernel(const double* __restrict__ samples, uint64_t n, int obs_id, double kappa, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x_1 = 1 - samples[2 * i], x_2 = 1 - samples[2 * i + 1];
        const double x_3 = 2 - x_1 - x_2;
        out[i] = pair_observable(obs_id, x_1, x_2, kappa);
        out[n + i] = pair_observable(obs_id, x_2, x_3, kappa);
        out[2 * n + i] = pair_observable(obs_id, x_3, x_1, kappa);
    }
}

// fused: generate -> three pairs -> observable, energy weights -> histogram (and / or min, max of the observable)
__global__ void __launch_bounds__(kProcThreads)
process_ewoc_kernel(ProcParams p, uint64_t n, uint64_t seed, uint64_t first, int obs_id, double kappa, double w1, double w2,
                    const double* __restrict__ edges, int n_edges, double* __restrict__ g_sum_w,
                    double* __restrict__ g_sum_w2, unsigned long long* __restrict__ g_count, int* g_poison,
                    double* __restrict__ d_minmax, unsigned long long* g_rejected, int* g_error, int shift) {
    extern __shared__ __align__(16) double s_raw[];
    __shared__ int s_poison;
    HistSmem h(s_raw, n_edges > 0 ? n_edges : 1, shift);
    const bool do_hist = edges != nullptr;
    if (threadIdx.x == 0) s_poison = INT_MAX;
    BinGuess G{0, 0.f, 0.f};
    if (do_hist) {
        hist_init(h, edges, n_edges);
        G = derive_bin_guess(h.edges, n_edges);
    } else {
        __syncthreads();
    }
    double mn = INFINITY, mx = -INFINITY;
    bool has_nan = false;
    unsigned long long rej = 0ull;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a, b, w;
        uint32_t r;
        if (!proc_draw(p, first + i, seed, a, b, w, r)) atomicExch(g_error, 1);
        rej += r;
        const double x_1 = 1 - a, x_2 = 1 - b;
        const double x_3 = 2 - x_1 - x_2;
        auto pair = [&](double x_i, double x_j) {
            const double obs = pair_observable(obs_id, x_i, x_j, kappa);
            // z1**w1 * z2**w2 * weights (montecarlo/ewoc.py:23-27), z_i = x_i / 2
            const double wk = small_pow(x_i / 2, w1) * small_pow(x_j / 2, w2) * w;
            if (do_hist) hist_add(h, n_edges, obs, wk, true, &s_poison, G);
            if (d_minmax) {
                has_nan |= (obs != obs);
                mn = fmin(mn, obs);
                mx = fmax(mx, obs);
            }
        };
        pair(x_1, x_2);
        pair(x_2, x_3);
        pair(x_3, x_1);
    }
    if (do_hist) hist_flush(h, n_edges, g_sum_w, g_sum_w2, g_count, &s_poison, g_poison);
    if (d_minmax) minmax_merge(mn, mx, has_nan, d_minmax);
    add_rejected(rej, g_rejected);
}

__global__ void process_poison_kernel(int* g_poison, int nb, double* g_sum_w, double* g_sum_w2) {
    const int pb = *g_poison;
    if (pb < 0 || pb >= nb) return;
    for (int k = pb + threadIdx.x; k < nb; k += blockDim.x) { g_sum_w[k] = qnan(); g_sum_w2[k] = qnan(); }
}

static int make_proc_params(const jmc_process* pr, ProcParams* out) {
    JMC_REQUIRE(pr != nullptr, "process is NULL");
    JMC_REQUIRE(pr->kind == JMC_PROCESS_EE2HADRONS_LO, "unknown process kind %d", pr->kind);
    JMC_REQUIRE(pr->method == JMC_LIN || pr->method == JMC_LOG, "%dis not a supported sample method.", pr->method);
    ProcParams p;
    std::memset(&p, 0, sizeof(p));
    p.kind = pr->kind;
    p.is_log = pr->method == JMC_LOG;
    if (p.is_log) {
        JMC_REQUIRE(pr->epsilon > 0.0 && pr->epsilon < 1.0, "epsilon=%g is invalid for log sampling.", pr->epsilon);
        p.log_eps = std::log(pr->epsilon);
    }
    for (int k = 0; k < 2; ++k) JMC_REQUIRE(pr->hi[k] > pr->lo[k], "kinematic bound %d is empty", k);
    p.lo0 = pr->lo[0]; p.width0 = pr->hi[0] - pr->lo[0]; p.log_width0 = std::log(pr->hi[0] - pr->lo[0]);
    p.lo1 = pr->lo[1]; p.width1 = pr->hi[1] - pr->lo[1]; p.log_width1 = std::log(pr->hi[1] - pr->lo[1]);
    p.prefactor = pr->prefactor;
    *out = p;
    return JMC_OK;
}

static int proc_grid(uint64_t n) {
    DeviceInfo di;
    if (device_info(&di)) return -1;
    const uint64_t want = (n + kProcThreads - 1) / kProcThreads, cap = (uint64_t)di.sm_count * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// error flag + rejected counter live in the per-stream scratch; the flag is read back (the reference's loop would
// simply never terminate on a box without accepted points)
static int finish_process(int* d_error, cudaStream_t st) {
    int h_err = 0;
    JMC_CUDA(cudaMemcpyAsync(&h_err, d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
    JMC_CUDA(cudaStreamSynchronize(st));
    if (h_err) return set_error(JMC_EDOMAIN, "process: no phase-space point accepted in %u trials", kMaxTrials);
    return JMC_OK;
}

}  // namespace jmc

using namespace jmc;

extern "C" int jmc_process_sample(const jmc_process* pr, uint64_t n, uint64_t seed, uint64_t first_index,
                                  double* d_samples, double* d_weights, uint64_t* d_rejected, void* stream) {
    ProcParams p;
    int rc = make_proc_params(pr, &p);
    if (rc) return rc;
    JMC_REQUIRE(d_rejected != nullptr, "jmc_process_sample: the rejected-trials counter is NULL");
    if (n == 0) return JMC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int* d_error;
    rc = scratch(sizeof(int), (void**)&d_error, stream);
    if (rc) return rc;
    JMC_CUDA(cudaMemsetAsync(d_error, 0, sizeof(int), st));
    const int grid = proc_grid(n);
    if (grid < 0) return JMC_ENODEV;
    process_draw_kernel<<<grid, kProcThreads, 0, st>>>(p, n, seed, first_index, d_samples, d_weights,
                                                        (unsigned long long*)d_rejected, d_error);
    JMC_LAUNCHED("process_draw_kernel");
    return finish_process(d_error, st);
}

extern "C" int jmc_process_pairs(const jmc_process* pr, const double* d_samples, uint64_t n, int obs_id, double kappa,
                                 double* d_obs, void* stream) {
    ProcParams p;
    int rc = make_proc_params(pr, &p);
    if (rc) return rc;
    JMC_REQUIRE(obs_id >= JMC_PAIR_COSTHETA && obs_id <= JMC_PAIR_Z_J, "jmc_process_pairs: unknown observable %d", obs_id);
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_samples && d_obs, "jmc_process_pairs: NULL array");
    const int grid = proc_grid(n);
    if (grid < 0) return JMC_ENODEV;
    process_pairs_kernel<<<grid, kProcThreads, 0, (cudaStream_t)stream>>>(d_samples, n, obs_id, kappa, d_obs);
    JMC_LAUNCHED("process_pairs_kernel");
    return JMC_OK;
}

extern "C" int jmc_process_ewoc(const jmc_process* pr, uint64_t n, uint64_t seed, uint64_t first_index, int obs_id,
                                double kappa, double w1, double w2, const double* d_edges, int n_edges,
                                double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax,
                                uint64_t* d_rejected, void* stream) {
    ProcParams p;
    int rc = make_proc_params(pr, &p);
    if (rc) return rc;
    JMC_REQUIRE(obs_id >= JMC_PAIR_COSTHETA && obs_id <= JMC_PAIR_Z_J, "jmc_process_ewoc: unknown observable %d", obs_id);
    JMC_REQUIRE(d_rejected != nullptr, "jmc_process_ewoc: the rejected-trials counter is NULL");
    size_t smem = 0;
    int shift = 0;
    if (d_edges) {
        rc = check_hist_args("jmc_process_ewoc", d_edges, n_edges, &smem);
        if (rc) return rc;
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_process_ewoc: histogram outputs missing");
        shift = smem <= 28 * 1024 ? hist_spread_shift(n_edges, 28 * 1024) : 0;
        smem = hist_smem_bytes(n_edges, shift);
    } else {
        JMC_REQUIRE(d_minmax != nullptr, "jmc_process_ewoc: neither edges nor minmax requested");
        n_edges = 0;
    }
    if (n == 0) return JMC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* scr;
    rc = scratch(256, (void**)&scr, stream);
    if (rc) return rc;
    int* d_error = (int*)scr;
    int* d_poison = (int*)(scr + 64);
    JMC_CUDA(cudaMemsetAsync(d_error, 0, sizeof(int), st));
    JMC_CUDA(cudaMemsetAsync(d_poison, 0x7f, sizeof(int), st));
    JMC_CUDA(cudaFuncSetAttribute(process_ewoc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = proc_grid(n);
    if (grid < 0) return JMC_ENODEV;
    process_ewoc_kernel<<<grid, kProcThreads, smem, st>>>(p, n, seed, first_index, obs_id, kappa, w1, w2, d_edges, n_edges,
                                                         d_sum_w, d_sum_w2, (unsigned long long*)d_count, d_poison, d_minmax,
                                                         (unsigned long long*)d_rejected, d_error, shift);
    JMC_LAUNCHED("process_ewoc_kernel");
    if (d_edges) {
        process_poison_kernel<<<1, 256, 0, st>>>(d_poison, n_edges - 1, d_sum_w, d_sum_w2);
        JMC_LAUNCHED("process_poison_kernel");
    }
    return finish_process(d_error, st);
}
