// Host-side plumbing shared by the translation units of libjmc_b200.so:
// error strings, launch accounting, per-device scratch, launch geometry.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/jmc_b200.h"

namespace jmc {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();

#define JMC_CUDA(call)                                         \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return ::jmc::cuda_fail(_e, #call); \
    } while (0)

#define JMC_REQUIRE(cond, ...)                                 \
    do {                                                       \
        if (!(cond)) return ::jmc::set_error(JMC_EINVAL, __VA_ARGS__); \
    } while (0)

// after a kernel launch
#define JMC_LAUNCHED(name)                                     \
    do {                                                       \
        ::jmc::count_launch();                                 \
        cudaError_t _e = cudaGetLastError();                   \
        if (_e != cudaSuccess) return ::jmc::cuda_fail(_e, name); \
    } while (0)

struct DeviceInfo {
    int device;
    int sm_count;
    int max_smem_optin;
};
int device_info(DeviceInfo* out);

// small device scratch per (host thread, device, stream): poison flags, FP32 edges, staged batch parameters
int scratch(size_t bytes, void** ptr, void* stream);
int pinned_scratch(size_t bytes, void** ptr);

// Shared-memory accumulation of (sum w, sum w^2).  FP64 (and FP32) shared-memory atomic adds are compare-and-swap
// loops on sm_100 (ATOMS.CAST.SPIN); only 32-bit integer ones are native.  Measured alternative: interleaving the
// two sums and updating both with ONE 128-bit CAS (ATOMS.CAS.128 exists) is 3-4x SLOWER on 100-bin histograms
// where every sample carries a weight (C1: 4.8e10 vs 2.1e11 samples/s; the wide CAS serialises harder under
// contention), so the sums stay in two arrays with one 64-bit CAS loop each.

// Derived constants of an integrand, computed once per launch on the host.
struct RunParams {
    jmc_integrand g;
    double log_eps;       // ln(epsilon)
    double lw_crit_z;     // ln(1/2 - z_cut)
    double lw_radius;     // ln(radius)
    double lw_half;       // ln(1/2)
    double lw_pre;        // ln(z_cut)
    double lw_simple;     // ln(hi - lo)
};
int make_run_params(const jmc_integrand* g, RunParams* out);

// An FP32 job prepared on the host (jmc_fast.cu): kernel parameters + FP32 edges in the compare domain.
struct FastPlan;
int fast_plan_make(const jmc_integrand* g, const double* h_edges, int n_edges, FastPlan** out);   // h_edges NULL: min/max pass
void fast_plan_free(FastPlan* p);
const float* fast_plan_edges(const FastPlan* p);                                                   // n_edges floats (host)
int fast_plan_run(const FastPlan* plan, const float* d_edges_f, int* d_poison, uint64_t n, uint64_t seed,
                  uint64_t first_index, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax,
                  void* stream, bool defer_poison = false);
// jmc_finalize that also applies (and re-arms) the NaN flag a fused run left in *d_poison (jmc_exact.cu)
int finalize_poisoned(double* d_sum_w, double* d_sum_w2, const uint64_t* d_count, const double* d_edges, int n_edges,
                      double area, uint64_t n_total, int bc_kind, double bc_value, double* d_density,
                      double* d_density_err, double* d_integral, double* d_integral_err, int* d_poison, void* stream);

}  // namespace jmc
