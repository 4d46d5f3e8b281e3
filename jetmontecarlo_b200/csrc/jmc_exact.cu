// FP64 kernels of libjmc_b200.so (compiled with --fmad=false so that the
// arithmetic is IEEE FP64 in NumPy operation order):
//   - array functions and whole-integrand replay (jmc_eval_fn, jmc_eval_integrand)
//   - Philox samplers on request (jmc_sample, jmc_uniforms)
//   - np.histogram restated: bin index, (sum w, sum w^2, count) with shared-
//     memory privatisation, NaN poisoning, min/max
//   - density / cumulative-integral epilogue with a block prefix scan
//   - the exact (FP64) mode of the fused generate->weight->bin kernel
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "jmc_common.cuh"
#include "jmc_integrand.cuh"
#include "jmc_hist.cuh"

namespace jmc {

static constexpr int kThreads = 256;

// ---------------------------------------------------------------------------
// array functions
// ---------------------------------------------------------------------------
struct FnArgs {
    const double* a[4];
    int64_t stride[4];
    double scalar[4];
};

__device__ __forceinline__ double fn_arg(const FnArgs& A, int k, uint64_t i) {
    return A.a[k] ? A.a[k][i * A.stride[k]] : A.scalar[k];
}

// One instantiation per function id: the register allocation of the simple functions (products, observables,
// fixed-coupling forms) is no longer that of the running-coupling radiators (the single switch kernel of round 1
// needed 226 registers for all 23 of them), and the dead branches are gone.
template <int FN, bool ACTIVE>
__device__ __forceinline__ double eval_one(const jmc_fn_params& p, double x0, double x1, double x2, double x3,
                                           const CritConsts* cc) {
    switch (FN) {
        case JMC_FN_ALPHA_S: return alpha_running(x0, x1);
        case JMC_FN_SPLITTING: return splitting(x0, p.jet, p.acc);
        case JMC_FN_RADIATOR_WEIGHT: return weight_radiator(x0, x1, p.jet, p.fixed_coupling, p.acc);
        case JMC_FN_CRIT_PDF_FC: return pdf_crit_fc(x0, x1, p.z_cut, p.jet);
        case JMC_FN_CRIT_RAD_FC: return rad_crit_fc(x0, p.z_cut, p.jet);
        case JMC_FN_SUB_RAD_FC: return rad_sub_fc(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_RADPRIME_FC: return radprime_sub_fc(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_PDF_FC: return pdf_sub_fc(x0, p.beta, p.jet, x1);
        case JMC_FN_PRE_RAD_FC: return rad_pre_fc(x0, x1, p.z_cut, p.jet, x2);
        case JMC_FN_CRIT_RAD_RC: return rad_crit_rc<ACTIVE>(x0, p.z_cut, p.jet, cc);
        case JMC_FN_CRIT_RADPRIME_RC: return radprime_crit_rc(x0, p.z_cut, p.jet);
        // (the *_surely_nan tests return early with the NaN the full evaluation would produce, see jmc_integrand.cuh)
        case JMC_FN_CRIT_PDF_RC: return crit_rc_surely_nan(x1, p.z_cut) ? qnan() : pdf_crit_rc<ACTIVE>(x0, x1, p.z_cut, p.jet, cc);
        case JMC_FN_SUB_RAD_RC: return rad_sub_rc<ACTIVE>(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_RADPRIME_RC: return radprime_sub_rc(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_PDF_RC: return sub_rc_surely_nan(x0, p.beta, x1) ? qnan() : pdf_sub_rc<ACTIVE>(x0, p.beta, p.jet, x1);
        case JMC_FN_PRE_WEIGHT: return weight_precritical(x0, x1, p.z_cut, p.jet);
        case JMC_FN_C_GROOMED: return ecf_groomed(x0, x1, p.z_cut, p.beta, x2, x3, p.acc);
        case JMC_FN_C_UNGROOMED: return ecf_ungroomed(x0, x1, p.beta, p.acc);
        case JMC_FN_TWO_EMISSION:
            return (!p.fixed_coupling && (crit_rc_surely_nan(x1, p.z_cut)
                                          || sub_rc_surely_nan(csub_verbatim(x2, x3, p.beta), p.beta, x1)))
                       ? qnan() : weight_two_emission<ACTIVE>(x0, x1, x2, x3, p.z_cut, p.beta, p.jet, p.fixed_coupling, cc);
        case JMC_FN_PRE_WEIGHT_ZEROBIN: return weight_precritical_zerobin(x0, x1, p.z_cut, p.jet, p.epsilon);
        case JMC_FN_ALPHA_SPLITTING:
            return (p.fixed_coupling ? JMC_ALPHA_FIXED : alpha_running(x0, x1)) * splitting(x0, p.jet, p.acc);
        case JMC_FN_MUL: return x0 * x1;
        case JMC_FN_MASK_FINITE: return (isfinite(x0) && isfinite(x1)) ? x0 : 0.0;
        case JMC_FN_EWOC_WEIGHT: {       // NumPy order: (z1**w1 * z2**w2) * w; small integer powers as np.power forms them
            auto pw = [](double x, double e) { return e == 1.0 ? x : (e == 2.0 ? x * x : (e == 0.0 ? 1.0 : pow(x, e))); };
            return pw(x0, x3) * pw(x1, p.beta) * x2;
        }
        case JMC_FN_SUB_SC1: return sub_sc1(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_SC2: return sub_sc2(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_SC3: return sub_sc3(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_HC1: return sub_hc1(x0, p.beta, p.jet, x1);
        case JMC_FN_SUB_HC2: return sub_hc2(x0, p.beta, p.jet, x1);
        case JMC_FN_CRIT_SC1: return crit_sc1(x0, p.z_cut, p.jet, x1);
        case JMC_FN_CRIT_SC2: return crit_sc2(x0, p.z_cut, p.jet, x1);
        case JMC_FN_CRIT_SC3: return crit_sc3(x0, p.z_cut, p.jet, x1);
        case JMC_FN_CRIT_HC1: return crit_hc1(x0, p.z_cut, p.jet, x1);
        case JMC_FN_CRIT_HC2: return crit_hc2(x0, p.z_cut, p.jet, x1);
        default: return qnan();
    }
}
// number of arguments a function reads (the others are never loaded)
template <int FN>
struct FnArity {
    static constexpr int value =
        (FN == JMC_FN_SPLITTING || FN == JMC_FN_CRIT_RAD_FC || FN == JMC_FN_CRIT_RAD_RC || FN == JMC_FN_CRIT_RADPRIME_RC) ? 1
        : (FN == JMC_FN_PRE_RAD_FC) ? 3
        : (FN == JMC_FN_C_GROOMED || FN == JMC_FN_TWO_EMISSION || FN == JMC_FN_EWOC_WEIGHT) ? 4 : 2;
};

// "the reference's value is certainly NaN" for the running-coupling functions (the tests of jmc_integrand.cuh): the
// cheap part of eval_one, used by the compacting form of the kernel to decide which elements need the radiators
template <int FN>
__device__ __forceinline__ bool eval_surely_nan(const jmc_fn_params& p, double x0, double x1, double x2, double x3) {
    if (FN == JMC_FN_CRIT_PDF_RC) return crit_rc_surely_nan(x1, p.z_cut);
    if (FN == JMC_FN_SUB_PDF_RC) return sub_rc_surely_nan(x0, p.beta, x1);
    if (FN == JMC_FN_TWO_EMISSION)
        return !p.fixed_coupling && (crit_rc_surely_nan(x1, p.z_cut) || sub_rc_surely_nan(csub_verbatim(x2, x3, p.beta), p.beta, x1));
    return false;
}

// COMPACT (running-coupling pdfs and twoEmissionWeight): most elements of a C3 sample set lie below the Landau scale
// (99.7 %) and are NaN after a few operations; the others need ~60 FP64 transcendentals.  Evaluated in place, one
// warp in eleven runs the radiator code with one or two live lanes and the kernel is an order of magnitude off its
// 40 B / element HBM roofline.  A warp writes the NaNs at once, appends the arguments and index of the other
// elements to its shared-memory queue and evaluates 32 of them at a time with every lane busy.
static constexpr int kFnQueueSlots = 64;

template <int FN, bool ACTIVE, bool COMPACT>
__global__ void __launch_bounds__(kThreads) eval_fn_kernel(jmc_fn_params p, FnArgs A, uint64_t n,
                                                            double* __restrict__ out) {
    constexpr int NA = FnArity<FN>::value;
    constexpr bool CRIT_RC = FN == JMC_FN_CRIT_RAD_RC || FN == JMC_FN_CRIT_PDF_RC || FN == JMC_FN_TWO_EMISSION;
    // region-boundary values of the critical radiator: functions of (z_c, jet), once per thread
    CritConsts consts{0.0, 0.0, 0.0};
    // (fixed_coupling selects a branch of twoEmissionWeight only; the two radiator functions have no such argument)
    const bool wanted = CRIT_RC && ACTIVE && (FN != JMC_FN_TWO_EMISSION || !p.fixed_coupling);
    if (wanted) consts = make_crit_consts(p.z_cut, p.jet);
    const CritConsts* cc = wanted ? &consts : nullptr;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (!COMPACT) {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const double x0 = fn_arg(A, 0, i), x1 = NA > 1 ? fn_arg(A, 1, i) : 0.0, x2 = NA > 2 ? fn_arg(A, 2, i) : 0.0,
                         x3 = NA > 3 ? fn_arg(A, 3, i) : 0.0;
            out[i] = eval_one<FN, ACTIVE>(p, x0, x1, x2, x3, cc);
        }
        return;
    }
    __shared__ double s_q[kThreads / 32][4][kFnQueueSlots];
    __shared__ unsigned long long s_qi[kThreads / 32][kFnQueueSlots];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int qn = 0;                                          // warp-uniform
    auto drain = [&](int slot) {
        out[s_qi[warp][slot]] = eval_one<FN, ACTIVE>(p, s_q[warp][0][slot], s_q[warp][1][slot], s_q[warp][2][slot],
                                                     s_q[warp][3][slot], cc);
    };
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n; i0 += stride) {
        const uint64_t i = i0 + lane;
        const bool valid = i < n;
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        bool active = false;
        if (valid) {
            x0 = fn_arg(A, 0, i);
            x1 = NA > 1 ? fn_arg(A, 1, i) : 0.0;
            x2 = NA > 2 ? fn_arg(A, 2, i) : 0.0;
            x3 = NA > 3 ? fn_arg(A, 3, i) : 0.0;
            active = !eval_surely_nan<FN>(p, x0, x1, x2, x3);
            if (!active) out[i] = qnan();
        }
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (!m) continue;
        if (active) {
            const int slot = qn + __popc(m & ((1u << lane) - 1u));
            s_q[warp][0][slot] = x0;
            s_q[warp][1][slot] = x1;
            s_q[warp][2][slot] = x2;
            s_q[warp][3][slot] = x3;
            s_qi[warp][slot] = i;
        }
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            qn -= 32;
            drain(qn + (int)lane);
            __syncwarp();
        }
    }
    __syncwarp();
    if ((int)lane < qn) drain((int)lane);
}

template <int FN>
static void launch_eval_fn(int fn, const jmc_fn_params& p, const FnArgs& A, uint64_t n, double* out, int grid,
                           cudaStream_t st) {
    // the running-coupling functions evaluate the active region only where its derivation holds (beta > 1, z_c < 1/2);
    // other parameters get the all-regions instantiation
    constexpr bool RC = FN == JMC_FN_CRIT_RAD_RC || FN == JMC_FN_CRIT_PDF_RC || FN == JMC_FN_SUB_RAD_RC
                        || FN == JMC_FN_SUB_PDF_RC || FN == JMC_FN_TWO_EMISSION;
    constexpr bool CAN_COMPACT = FN == JMC_FN_CRIT_PDF_RC || FN == JMC_FN_SUB_PDF_RC || FN == JMC_FN_TWO_EMISSION;
    const bool sub_fn = FN == JMC_FN_SUB_RAD_RC || FN == JMC_FN_SUB_PDF_RC || FN == JMC_FN_TWO_EMISSION;
    const bool crit_fn = FN == JMC_FN_CRIT_RAD_RC || FN == JMC_FN_CRIT_PDF_RC || FN == JMC_FN_TWO_EMISSION;
    const bool regular = (!sub_fn || p.beta > 1.0) && (!crit_fn || p.z_cut < 0.5);
    if (fn == FN) {
        if (RC && !regular) eval_fn_kernel<FN, !RC, false><<<grid, kThreads, 0, st>>>(p, A, n, out);
        else if (CAN_COMPACT && !(FN == JMC_FN_TWO_EMISSION && p.fixed_coupling))
            eval_fn_kernel<FN, true, CAN_COMPACT><<<grid, kThreads, 0, st>>>(p, A, n, out);
        else eval_fn_kernel<FN, true, false><<<grid, kThreads, 0, st>>>(p, A, n, out);
    } else if constexpr (FN + 1 < JMC_FN_COUNT) {
        launch_eval_fn<FN + 1>(fn, p, A, n, out, grid, st);
    }
}

static int grid_for(uint64_t n, int threads, int blocks_per_sm = 8) {
    DeviceInfo di;
    if (device_info(&di) != 0) return -1;
    uint64_t want = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)di.sm_count * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace jmc

using namespace jmc;

extern "C" int jmc_eval_fn(int fn, const jmc_fn_params* p, uint64_t n, const double* d_a0, int64_t stride0,
                           const double* d_a1, int64_t stride1, const double* d_a2, int64_t stride2,
                           const double* d_a3, int64_t stride3, double* d_out, void* stream) {
    JMC_REQUIRE(p != nullptr, "jmc_eval_fn: params is NULL");
    JMC_REQUIRE(fn >= 0 && fn < JMC_FN_COUNT, "jmc_eval_fn: unknown function id %d", fn);
    JMC_REQUIRE(p->jet == JMC_QUARK || p->jet == JMC_GLUON, "jmc_eval_fn: jet must be quark or gluon");
    JMC_REQUIRE(p->acc >= JMC_ACC_LL && p->acc <= JMC_ACC_MLL, "jmc_eval_fn: invalid accuracy");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_out != nullptr, "jmc_eval_fn: out is NULL");
    FnArgs A;
    A.a[0] = d_a0; A.a[1] = d_a1; A.a[2] = d_a2; A.a[3] = d_a3;
    A.stride[0] = stride0; A.stride[1] = stride1; A.stride[2] = stride2; A.stride[3] = stride3;
    A.scalar[0] = p->s0; A.scalar[1] = p->s1; A.scalar[2] = p->s2; A.scalar[3] = p->s3;
    // documented scalar stand-ins
    if ((fn >= JMC_FN_SUB_RAD_FC && fn <= JMC_FN_SUB_PDF_FC) || (fn >= JMC_FN_SUB_RAD_RC && fn <= JMC_FN_SUB_PDF_RC))
        A.scalar[1] = p->max_radius;
    if (fn == JMC_FN_C_GROOMED) { A.scalar[2] = p->z_pre; A.scalar[3] = p->f_soft; }
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    launch_eval_fn<0>(fn, *p, A, n, d_out, grid, (cudaStream_t)stream);
    JMC_LAUNCHED("eval_fn_kernel");
    return JMC_OK;
}

// ---------------------------------------------------------------------------
// sampling on request
// ---------------------------------------------------------------------------
namespace jmc {

// plumbing the Python mirror would otherwise ask an array library for: a constant into an FP64 array
__global__ void __launch_bounds__(kThreads) fill_kernel(double* __restrict__ out, uint64_t n, double value) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = value;
}

// out[r * n + i] = in[i]: np.tile(in, reps)
__global__ void __launch_bounds__(kThreads) tile_kernel(const double* __restrict__ in, uint64_t n, int reps,
                                                         double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = in[i];
        for (int r = 0; r < reps; ++r) out[(uint64_t)r * n + i] = v;
    }
}

// ---- packing for the collectives of a sharded job (jetmontecarlo_b200/distributed.py) ----
// [sum w | sum w^2 | count as FP64] <-> the three histograms; counts are exact in FP64 up to 2^53
__global__ void __launch_bounds__(kThreads) pack_hist_kernel(const double* __restrict__ sum_w, const double* __restrict__ sum_w2,
                                                              const unsigned long long* __restrict__ count, uint64_t nb,
                                                              double* __restrict__ buf, int unpack,
                                                              double* __restrict__ o_sum_w, double* __restrict__ o_sum_w2,
                                                              unsigned long long* __restrict__ o_count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nb; k += stride) {
        if (!unpack) {
            buf[k] = sum_w[k];
            buf[nb + k] = sum_w2[k];
            buf[2 * nb + k] = (double)count[k];
        } else {
            o_sum_w[k] = buf[k];
            o_sum_w2[k] = buf[nb + k];
            o_count[k] = (unsigned long long)llrint(buf[2 * nb + k]);
        }
    }
}
// rows (min, max) <-> rows (-min, max, has-NaN) with NaN -> -inf: ONE MAX all-reduce merges minima, maxima and NaN flags
__global__ void __launch_bounds__(kThreads) pack_minmax_kernel(double* __restrict__ mm, uint64_t n, double* __restrict__ buf,
                                                                int unpack) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        if (!unpack) {
            const double lo = mm[2 * k], hi = mm[2 * k + 1];
            const bool bad = lo != lo || hi != hi;
            buf[3 * k] = lo != lo ? -INFINITY : -lo;
            buf[3 * k + 1] = hi != hi ? -INFINITY : hi;
            buf[3 * k + 2] = bad ? 1.0 : 0.0;
        } else {
            const bool bad = buf[3 * k + 2] > 0.0;
            mm[2 * k] = bad ? qnan() : -buf[3 * k];
            mm[2 * k + 1] = bad ? qnan() : buf[3 * k + 1];
        }
    }
}

// number of finite entries
__global__ void __launch_bounds__(kThreads) count_finite_kernel(const double* __restrict__ in, uint64_t n,
                                                                 unsigned long long* __restrict__ count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += isfinite(in[i]) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void __launch_bounds__(kThreads) uniforms_kernel(uint64_t n, uint64_t seed, uint64_t first, int n_uniform,
                                                             double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int n_blocks = (n_uniform + 1) / 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        for (int b = 0; b < n_blocks; ++b) {
            double u0, u1;
            philox_uniform53x2(first + i, (uint32_t)b, seed, u0, u1);
            out[i * n_uniform + 2 * b] = u0;
            if (2 * b + 1 < n_uniform) out[i * n_uniform + 2 * b + 1] = u1;
        }
    }
}

__global__ void __launch_bounds__(kThreads) sample_kernel(RunParams p, int sampler_kind, uint64_t n, uint64_t seed,
                                                           uint64_t first, double* __restrict__ samples,
                                                           double* __restrict__ jac) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool is_log = p.g.method == JMC_LOG;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double u0, u1;
        philox_uniform53x2(first + i, 0u, seed, u0, u1);
        double j;
        if (sampler_kind == JMC_SAMPLER_CRITICAL) {
            double z, th;
            sample_critical(p, u0, u1, z, th);
            samples[2 * i] = z; samples[2 * i + 1] = th;
            j = is_log ? (z - p.g.z_cut) * th : 1.0;
        } else if (sampler_kind == JMC_SAMPLER_UNGROOMED) {
            double z, th;
            sample_ungroomed(p, u0, u1, z, th);
            samples[2 * i] = z; samples[2 * i + 1] = th;
            j = is_log ? z * th : 1.0;
        } else if (sampler_kind == JMC_SAMPLER_PRECRITICAL) {
            const double z = sample_precritical(p, u0);
            samples[i] = z;
            j = is_log ? z : 1.0;
        } else {
            const double x = sample_simple(p, u0);
            samples[i] = x;
            j = is_log ? x : 1.0;
        }
        if (jac) jac[i] = j;
    }
}

}  // namespace jmc

extern "C" int jmc_uniforms(uint64_t n, uint64_t seed, uint64_t first_index, int n_uniform, double* d_out,
                            void* stream) {
    JMC_REQUIRE(n_uniform >= 1 && n_uniform <= 64, "jmc_uniforms: n_uniform must be in 1..64");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_out != nullptr, "jmc_uniforms: out is NULL");
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    uniforms_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, seed, first_index, n_uniform, d_out);
    JMC_LAUNCHED("uniforms_kernel");
    return JMC_OK;
}

extern "C" int jmc_fill(double* d_out, uint64_t n, double value, void* stream) {
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_out != nullptr, "jmc_fill: out is NULL");
    if (value == 0.0) {                                  // +0.0 is all-zero bytes: no kernel needed
        JMC_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * n, (cudaStream_t)stream));
        return JMC_OK;
    }
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    fill_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_out, n, value);
    JMC_LAUNCHED("fill_kernel");
    return JMC_OK;
}

extern "C" int jmc_tile(const double* d_in, uint64_t n, int reps, double* d_out, void* stream) {
    JMC_REQUIRE(reps >= 1, "jmc_tile: reps must be positive");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_in && d_out, "jmc_tile: NULL array");
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    tile_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_in, n, reps, d_out);
    JMC_LAUNCHED("tile_kernel");
    return JMC_OK;
}

extern "C" int jmc_count_finite(const double* d_in, uint64_t n, uint64_t* h_count, void* stream) {
    JMC_REQUIRE(h_count != nullptr, "jmc_count_finite: count is NULL");
    *h_count = 0;
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_in != nullptr, "jmc_count_finite: NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_count;
    int rc = scratch(sizeof(unsigned long long), (void**)&d_count, stream);
    if (rc) return rc;
    JMC_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    count_finite_kernel<<<grid, kThreads, 0, st>>>(d_in, n, d_count);
    JMC_LAUNCHED("count_finite_kernel");
    unsigned long long h = 0;
    JMC_CUDA(cudaMemcpyAsync(&h, d_count, sizeof(h), cudaMemcpyDeviceToHost, st));
    JMC_CUDA(cudaStreamSynchronize(st));
    *h_count = h;
    return JMC_OK;
}

extern "C" int jmc_pack_histograms(double* d_sum_w, double* d_sum_w2, uint64_t* d_count, uint64_t nb, double* d_buf,
                                   int unpack, void* stream) {
    if (nb == 0) return JMC_OK;
    JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count && d_buf, "jmc_pack_histograms: NULL array");
    const int grid = grid_for(nb, kThreads);
    if (grid < 0) return JMC_ENODEV;
    pack_hist_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_sum_w, d_sum_w2, (const unsigned long long*)d_count, nb,
                                                                  d_buf, unpack, d_sum_w, d_sum_w2,
                                                                  (unsigned long long*)d_count);
    JMC_LAUNCHED("pack_hist_kernel");
    return JMC_OK;
}

extern "C" int jmc_pack_minmax(double* d_minmax, uint64_t n_pairs, double* d_buf, int unpack, void* stream) {
    if (n_pairs == 0) return JMC_OK;
    JMC_REQUIRE(d_minmax && d_buf, "jmc_pack_minmax: NULL array");
    const int grid = grid_for(n_pairs, kThreads);
    if (grid < 0) return JMC_ENODEV;
    pack_minmax_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_minmax, n_pairs, d_buf, unpack);
    JMC_LAUNCHED("pack_minmax_kernel");
    return JMC_OK;
}

extern "C" int jmc_memset(void* d_ptr, int byte_value, uint64_t n_bytes, void* stream) {
    if (n_bytes == 0) return JMC_OK;
    JMC_REQUIRE(d_ptr != nullptr, "jmc_memset: pointer is NULL");
    JMC_CUDA(cudaMemsetAsync(d_ptr, byte_value, n_bytes, (cudaStream_t)stream));
    return JMC_OK;
}

static int sampler_to_integrand(const jmc_sampler* s, jmc_integrand* g) {
    JMC_REQUIRE(s != nullptr, "sampler is NULL");
    JMC_REQUIRE(s->kind >= JMC_SAMPLER_SIMPLE && s->kind <= JMC_SAMPLER_UNGROOMED, "unknown sampler kind %d", s->kind);
    *g = jmc_integrand{};
    g->method = s->method;
    g->epsilon = s->epsilon;
    g->z_cut = s->zc;
    g->radius = s->radius;
    g->lo = s->lo;
    g->hi = s->hi;
    g->beta = 2.0;
    g->theta_scale = 1.0;
    g->fixed_coupling = 1;
    switch (s->kind) {
        case JMC_SAMPLER_SIMPLE: g->kind = JMC_KIND_SIMPLE; break;
        case JMC_SAMPLER_CRITICAL: g->kind = JMC_KIND_CRIT; break;
        case JMC_SAMPLER_PRECRITICAL: g->kind = JMC_KIND_PRE_CRIT_SUB; g->radius = 1.0; break;
        default: g->kind = JMC_KIND_RADIATOR; break;
    }
    return JMC_OK;
}

extern "C" int jmc_sample(const jmc_sampler* s, uint64_t n, uint64_t seed, uint64_t first_index, double* d_samples,
                          double* d_jac, void* stream) {
    jmc_integrand g;
    int rc = sampler_to_integrand(s, &g);
    if (rc) return rc;
    RunParams p;
    rc = make_run_params(&g, &p);
    if (rc) return rc;
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_samples != nullptr, "jmc_sample: samples is NULL");
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    sample_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, s->kind, n, seed, first_index, d_samples, d_jac);
    JMC_LAUNCHED("sample_kernel");
    return JMC_OK;
}

// inverse-transform sampling from a tabulated cdf (scipy interp1d semantics: linear, extrapolating)
namespace jmc {
__global__ void __launch_bounds__(kThreads)
inverse_transform_kernel(const double* __restrict__ cdf, const double* __restrict__ xs, int nt, uint64_t n,
                         uint64_t seed, uint64_t first, double* __restrict__ out) {
    extern __shared__ double s_tab[];
    double* s_c = s_tab;
    double* s_x = s_tab + nt;
    for (int k = threadIdx.x; k < nt; k += blockDim.x) { s_c[k] = cdf[k]; s_x[k] = xs[k]; }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double u, unused;
        philox_uniform53x2(first + i, 0u, seed, u, unused);
        int lo = 0, hi = nt;                   // searchsorted(cdf, u, 'left')
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_c[mid] < u) lo = mid + 1; else hi = mid;
        }
        int idx = lo < 1 ? 1 : (lo > nt - 1 ? nt - 1 : lo);
        const double c_lo = s_c[idx - 1], c_hi = s_c[idx], x_lo = s_x[idx - 1], x_hi = s_x[idx];
        const double slope = (x_hi - x_lo) / (c_hi - c_lo);
        out[i] = slope * (u - c_lo) + x_lo;
    }
}
}  // namespace jmc

extern "C" int jmc_inverse_transform(const double* d_cdf_vals, const double* d_x_vals, int n_table, uint64_t n,
                                     uint64_t seed, uint64_t first_index, double* d_out, void* stream) {
    JMC_REQUIRE(n_table >= 2, "jmc_inverse_transform: need at least two table points");
    JMC_REQUIRE(d_cdf_vals && d_x_vals, "jmc_inverse_transform: NULL table");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_out != nullptr, "jmc_inverse_transform: out is NULL");
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    const size_t smem = sizeof(double) * 2 * (size_t)n_table;
    if (smem > (size_t)di.max_smem_optin)
        return set_error(JMC_ELIMIT, "jmc_inverse_transform: table of %d points exceeds shared memory", n_table);
    JMC_CUDA(cudaFuncSetAttribute(inverse_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_for(n, kThreads);
    inverse_transform_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(d_cdf_vals, d_x_vals, n_table, n, seed,
                                                                              first_index, d_out);
    JMC_LAUNCHED("inverse_transform_kernel");
    return JMC_OK;
}

// ---------------------------------------------------------------------------
// whole-integrand replay
// ---------------------------------------------------------------------------
namespace jmc {

template <int KIND>
__global__ void __launch_bounds__(kThreads)
eval_integrand_kernel(RunParams p, uint64_t n, const double* __restrict__ crit, const double* __restrict__ sub,
                      const double* __restrict__ pre, const double* __restrict__ zb_u,
                      const double* __restrict__ edges, int n_edges, double* __restrict__ jac_out,
                      double* __restrict__ obs_out, double* __restrict__ w_out, int32_t* __restrict__ bin_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        SampleSet s;
        s.z_c = s.th_c = s.z_s = s.th_s = s.z_pre = s.zb_u = 0.0;
        if (crit) { s.z_c = crit[2 * i]; s.th_c = crit[2 * i + 1]; }
        if (sub) { s.z_s = sub[2 * i]; s.th_s = sub[2 * i + 1]; }
        if (pre) s.z_pre = pre[i];
        if (zb_u) s.zb_u = zb_u[i]; else s.zb_u = 2.0;   // no zero-bin draw -> never zeroed
        double jac, obs, wj;
        integrand_eval<KIND>(p, s, jac, obs, wj);
        if (jac_out) jac_out[i] = jac;
        if (obs_out) obs_out[i] = obs;
        if (w_out) w_out[i] = wj;
        if (bin_out) bin_out[i] = bin_index(obs, edges, n_edges);
    }
}

}  // namespace jmc

extern "C" int jmc_eval_integrand(const jmc_integrand* g, uint64_t n, const double* d_crit, const double* d_sub,
                                  const double* d_pre, const double* d_zb_u, const double* d_edges, int n_edges,
                                  double* d_jac, double* d_obs, double* d_weight, int32_t* d_bin, void* stream) {
    RunParams p;
    int rc = make_run_params(g, &p);
    if (rc) return rc;
    JMC_REQUIRE(g->kind != JMC_KIND_SHOWER, "jmc_eval_integrand: the shower has no sample-replay form");
    if (d_bin) JMC_REQUIRE(d_edges != nullptr && n_edges >= 2, "jmc_eval_integrand: bins requested without edges");
    const bool need_crit = g->kind == JMC_KIND_CRIT || g->kind == JMC_KIND_CRIT_SUB || g->kind == JMC_KIND_PRE_CRIT_SUB;
    const bool need_sub = g->kind == JMC_KIND_RADIATOR || g->kind == JMC_KIND_CRIT_SUB || g->kind == JMC_KIND_PRE_CRIT_SUB;
    const bool need_pre = g->kind == JMC_KIND_PRE_CRIT_SUB || g->kind == JMC_KIND_SIMPLE;
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(!need_crit || d_crit, "jmc_eval_integrand: critical samples missing");
    JMC_REQUIRE(!need_sub || d_sub, "jmc_eval_integrand: subsequent/ungroomed samples missing");
    JMC_REQUIRE(!need_pre || d_pre, "jmc_eval_integrand: pre-critical/simple samples missing");
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    cudaStream_t st = (cudaStream_t)stream;
#define JMC_LAUNCH_EVAL(K)                                                                                      \
    eval_integrand_kernel<K><<<grid, kThreads, 0, st>>>(p, n, d_crit, d_sub, d_pre, d_zb_u, d_edges, n_edges, \
                                                         d_jac, d_obs, d_weight, d_bin)
    switch (g->kind) {
        case JMC_KIND_RADIATOR: JMC_LAUNCH_EVAL(JMC_KIND_RADIATOR); break;
        case JMC_KIND_CRIT: JMC_LAUNCH_EVAL(JMC_KIND_CRIT); break;
        case JMC_KIND_CRIT_SUB: JMC_LAUNCH_EVAL(JMC_KIND_CRIT_SUB); break;
        case JMC_KIND_PRE_CRIT_SUB: JMC_LAUNCH_EVAL(JMC_KIND_PRE_CRIT_SUB); break;
        default: JMC_LAUNCH_EVAL(JMC_KIND_SIMPLE); break;
    }
#undef JMC_LAUNCH_EVAL
    JMC_LAUNCHED("eval_integrand_kernel");
    return JMC_OK;
}

// ---------------------------------------------------------------------------
// np.histogram restated
// ---------------------------------------------------------------------------
namespace jmc {

__global__ void __launch_bounds__(kThreads) bin_index_kernel(const double* __restrict__ obs, uint64_t n,
                                                              const double* __restrict__ edges, int n_edges,
                                                              int32_t* __restrict__ bin) {
    extern __shared__ double s_edges[];
    for (int k = threadIdx.x; k < n_edges; k += blockDim.x) s_edges[k] = edges[k];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        bin[i] = bin_index(obs[i], s_edges, n_edges);
}

// HBM-streaming: 16 B per sample (observable + weight).  Every thread issues the loads of JMC_HIST_UNROLL pairs of
// consecutive samples (128-bit loads when the arrays are 16-byte aligned: VEC) before it bins the first one.
// the streaming loop of hist_kernel for one (load width, guess mode, weights or none): chosen once per block
template <bool VEC, int MODE, bool HAS_W>
__device__ __forceinline__ void hist_stream(HistSmem& h, const double* __restrict__ obs, const double* __restrict__ w,
                                            uint64_t n, int n_edges, int* s_poison, const BinGuess& G) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    constexpr int U = JMC_HIST_UNROLL;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const uint64_t n2 = n >> 1;                          // pairs of consecutive samples
        const double2* obs2 = (const double2*)obs;
        const double2* w2 = (const double2*)w;
        for (; i + (U - 1) * stride < n2; i += U * stride) {
            double2 o[U], ww[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                o[k] = __ldcs(obs2 + i + k * stride);
                ww[k] = HAS_W ? __ldcs(w2 + i + k * stride) : make_double2(1.0, 1.0);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                hist_add<MODE>(h, n_edges, o[k].x, ww[k].x, HAS_W, s_poison, G);
                hist_add<MODE>(h, n_edges, o[k].y, ww[k].y, HAS_W, s_poison, G);
            }
        }
        for (; i < n2; i += stride) {
            const double2 o = obs2[i], ww = HAS_W ? w2[i] : make_double2(1.0, 1.0);
            hist_add<MODE>(h, n_edges, o.x, ww.x, HAS_W, s_poison, G);
            hist_add<MODE>(h, n_edges, o.y, ww.y, HAS_W, s_poison, G);
        }
        if ((n & 1ull) && blockIdx.x == 0 && threadIdx.x == 0)
            hist_add<MODE>(h, n_edges, obs[n - 1], HAS_W ? w[n - 1] : 1.0, HAS_W, s_poison, G);
    } else {
        for (; i + (U - 1) * stride < n; i += U * stride) {
            double o[U], ww[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                o[k] = __ldcs(obs + i + k * stride);
                ww[k] = HAS_W ? __ldcs(w + i + k * stride) : 1.0;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) hist_add<MODE>(h, n_edges, o[k], ww[k], HAS_W, s_poison, G);
        }
        for (; i < n; i += stride) hist_add<MODE>(h, n_edges, obs[i], HAS_W ? w[i] : 1.0, HAS_W, s_poison, G);
    }
}

template <bool VEC, bool HAS_W>
__global__ void __launch_bounds__(1024)
hist_kernel(const double* __restrict__ obs, const double* __restrict__ w, uint64_t n,
            const double* __restrict__ edges, int n_edges, double* __restrict__ g_sum_w,
            double* __restrict__ g_sum_w2, unsigned long long* __restrict__ g_count, int* g_poison, int shift) {
    extern __shared__ __align__(16) double s_raw[];
    __shared__ int s_poison;
    HistSmem h(s_raw, n_edges, shift);
    if (threadIdx.x == 0) s_poison = INT_MAX;
    hist_init(h, edges, n_edges);
    const BinGuess G = derive_bin_guess(h.edges, n_edges);
    // np.logspace (mode 2) / np.linspace (mode 1) edges are what every large job has: their loops carry no per-sample
    // look at the mode
    if (G.mode == 2) hist_stream<VEC, 2, HAS_W>(h, obs, w, n, n_edges, &s_poison, G);
    else if (G.mode == 1) hist_stream<VEC, 1, HAS_W>(h, obs, w, n, n_edges, &s_poison, G);
    else if (G.mode == 3) hist_stream<VEC, 3, HAS_W>(h, obs, w, n, n_edges, &s_poison, G);
    else hist_stream<VEC, 0, HAS_W>(h, obs, w, n, n_edges, &s_poison, G);
    hist_flush(h, n_edges, g_sum_w, g_sum_w2, g_count, &s_poison, g_poison);
}

// NaN-fill every bin at or above the poison bin
__global__ void poison_kernel(const int* g_poison, int nb, double* g_sum_w, double* g_sum_w2) {
    const int pb = *g_poison;
    if (pb < 0 || pb >= nb) return;
    for (int k = pb + blockIdx.x * blockDim.x + threadIdx.x; k < nb; k += gridDim.x * blockDim.x) {
        if (k < 0) continue;
        if (g_sum_w) g_sum_w[k] = qnan();
        if (g_sum_w2) g_sum_w2[k] = qnan();
    }
}

// HBM-streaming, 8 B per sample: 128-bit loads, four of them in flight per thread (the arrays start 16-byte aligned
// after `head` leading elements, which block 0 takes one by one together with the odd element at the end)
__global__ void __launch_bounds__(kThreads) minmax_kernel(const double* __restrict__ x, uint64_t n,
                                                           double* __restrict__ d_minmax) {
    double mn = INFINITY, mx = -INFINITY;
    bool has_nan = false;
    auto take = [&](double v) {
        has_nan |= (v != v);
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    };
    const uint64_t head = (((uintptr_t)x & 15u) && n) ? 1 : 0;
    const double2* x2 = (const double2*)(x + head);
    const uint64_t n2 = (n - head) >> 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldcs(x2 + i + k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) { take(v[k].x); take(v[k].y); }
    }
    for (; i < n2; i += stride) {
        const double2 v = x2[i];
        take(v.x);
        take(v.y);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (head) take(x[0]);
        if ((n - head) & 1ull) take(x[n - 1]);
    }
    minmax_merge(mn, mx, has_nan, d_minmax);
}

__global__ void minmax_init_kernel(double* d_minmax) {
    d_minmax[0] = INFINITY;
    d_minmax[1] = -INFINITY;
}

}  // namespace jmc

extern "C" int jmc_bin_index(const double* d_obs, uint64_t n, const double* d_edges, int n_edges, int32_t* d_bin,
                             void* stream) {
    JMC_REQUIRE(n_edges >= 2, "jmc_bin_index: need at least 2 edges");
    JMC_REQUIRE(d_edges != nullptr, "jmc_bin_index: edges is NULL");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_obs && d_bin, "jmc_bin_index: NULL array");
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    const size_t smem = sizeof(double) * (size_t)n_edges;
    if (smem > (size_t)di.max_smem_optin) return set_error(JMC_ELIMIT, "jmc_bin_index: %d edges exceed shared memory", n_edges);
    JMC_CUDA(cudaFuncSetAttribute(bin_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_for(n, kThreads);
    bin_index_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(d_obs, n, d_edges, n_edges, d_bin);
    JMC_LAUNCHED("bin_index_kernel");
    return JMC_OK;
}

extern "C" int jmc_hist(const double* d_obs, const double* d_weights, uint64_t n, const double* d_edges, int n_edges,
                        double* d_sum_w, double* d_sum_w2, uint64_t* d_count, void* stream) {
    size_t smem;
    int rc = check_hist_args("jmc_hist", d_edges, n_edges, &smem);
    if (rc) return rc;
    JMC_REQUIRE(d_count != nullptr, "jmc_hist: count is NULL");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_obs != nullptr, "jmc_hist: obs is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    int* d_poison;
    rc = scratch(sizeof(int), (void**)&d_poison, (void*)st);
    if (rc) return rc;
    JMC_CUDA(cudaMemsetAsync(d_poison, 0x7f, sizeof(int), st));   // 0x7f7f7f7f: "no poison"
    // big histograms (the reference's 5000 radiator bins take 140 KB) leave room for one block per SM: give that
    // block 1024 threads so that 32 warps share the histogram instead of 8
    const int threads = smem > 72 * 1024 ? 1024 : kThreads;
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    // (measured, 99 log bins, every weight non-zero: 16 copies 3.6 TB/s, 32 copies 3.1 TB/s -- the 50 KB per block of
    // the latter halve the resident threads and with them the loads in flight -- 8 copies 2.9, 4 copies 2.5, 1 copy 2.0)
    const int shift = hist_spread_shift(n_edges, threads == 1024 ? (size_t)di.max_smem_optin - 1024 : 33 * 1024);
    smem = hist_smem_bytes(n_edges, shift);
    const bool vec = (((uintptr_t)d_obs | (uintptr_t)d_weights) & 15u) == 0;      // (a NULL weight array is aligned)
    auto kern = vec ? (d_weights ? hist_kernel<true, true> : hist_kernel<true, false>)
                    : (d_weights ? hist_kernel<false, true> : hist_kernel<false, false>);
    JMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = grid_for(vec ? (n + 1) / 2 : n, threads, per_sm);
    kern<<<grid, threads, smem, st>>>(d_obs, d_weights, n, d_edges, n_edges, d_sum_w, d_sum_w2,
                                      (unsigned long long*)d_count, d_poison, shift);
    JMC_LAUNCHED("hist_kernel");
    if (d_weights) {
        poison_kernel<<<1, 256, 0, st>>>(d_poison, n_edges - 1, d_sum_w, d_sum_w2);
        JMC_LAUNCHED("poison_kernel");
    }
    return JMC_OK;
}

extern "C" int jmc_minmax(const double* d_x, uint64_t n, double* d_minmax, void* stream) {
    JMC_REQUIRE(d_minmax != nullptr, "jmc_minmax: output is NULL");
    JMC_REQUIRE(n > 0, "jmc_minmax: zero-size array has no minimum");   // np.min raises ValueError
    JMC_REQUIRE(d_x != nullptr, "jmc_minmax: input is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    minmax_init_kernel<<<1, 1, 0, st>>>(d_minmax);
    JMC_LAUNCHED("minmax_init_kernel");
    const int grid = grid_for((n + 1) / 2, kThreads);
    minmax_kernel<<<grid, kThreads, 0, st>>>(d_x, n, d_minmax);
    JMC_LAUNCHED("minmax_kernel");
    return JMC_OK;
}

// ---------------------------------------------------------------------------
// 2-D histogram and epilogue (integrator_2d; used by the pre-critical radiator, generation.py:143-250)
// ---------------------------------------------------------------------------
namespace jmc {

__global__ void __launch_bounds__(kThreads)
hist2d_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ w, uint64_t n,
              const double* __restrict__ ex, int nxe, const double* __restrict__ ey, int nye,
              double* __restrict__ g_sum_w, double* __restrict__ g_sum_w2, unsigned long long* __restrict__ g_count) {
    extern __shared__ double s_e[];
    double* s_ex = s_e;
    double* s_ey = s_e + nxe;
    for (int k = threadIdx.x; k < nxe; k += blockDim.x) s_ex[k] = ex[k];
    for (int k = threadIdx.x; k < nye; k += blockDim.x) s_ey[k] = ey[k];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int nby = nye - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int bx = bin_index(x[i], s_ex, nxe), by = bin_index(y[i], s_ey, nye);
        if (bx < 0 || by < 0) continue;
        const size_t b = (size_t)bx * nby + by;
        // (nx-1)(ny-1) bins do not fit shared memory in general: L2-resident global atomics (REDG.F64 / .64)
        atomicAdd(&g_count[b], 1ull);
        if (w) {
            const double wi = w[i];
            atomicAdd(&g_sum_w[b], wi);
            atomicAdd(&g_sum_w2[b], wi * wi);
        }
    }
}

// density[iy][ix] (transposed, as the reference stores it) and running sums along x then y
__global__ void density2d_kernel(const double* __restrict__ sum_w, const double* __restrict__ sum_w2,
                                 const unsigned long long* __restrict__ count, const double* __restrict__ ex, int nxe,
                                 const double* __restrict__ ey, int nye, double norm, double* __restrict__ density,
                                 double* __restrict__ density_err) {
    const int nbx = nxe - 1, nby = nye - 1;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nbx * nby; t += gridDim.x * blockDim.x) {
        const int ix = t / nby, iy = t % nby;
        // areafactor = area / (N * binArea)                          integrator.py:562-566
        const double bin_area = (ex[ix + 1] - ex[ix]) * (ey[iy + 1] - ey[iy]);
        const double f = norm / bin_area;
        double d = sum_w[t] * f, e = sqrt(sum_w2[t]) * f;
        if (count[t] == 0ull) { d = 0.0; e = 0.0; }
        density[(size_t)iy * nbx + ix] = d;
        density_err[(size_t)iy * nbx + ix] = e;
    }
}

// multidim_cumsum of density*binArea on the [nby, nbx] (transposed) array: cumsum along the last axis (x), then
// along y; `reverse` flips both axes (np.flip before and after).  One thread per line; the arrays are tiny.
__global__ void cumsum2d_kernel(const double* __restrict__ density, const double* __restrict__ density_err,
                                const double* __restrict__ ex, int nxe, const double* __restrict__ ey, int nye,
                                int bc_kind, double bc_value, double* __restrict__ integral,
                                double* __restrict__ integral_err, int pass) {
    const int nbx = nxe - 1, nby = nye - 1;
    const bool reverse = bc_kind != JMC_BC_FIRST;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (pass == 0) {                    // along x for each row iy
        if (t >= nby) return;
        const int iy = t;
        double a = 0.0, b = 0.0;
        for (int j = 0; j < nbx; ++j) {
            const int ix = reverse ? nbx - 1 - j : j;
            const double area = (ex[ix + 1] - ex[ix]) * (ey[iy + 1] - ey[iy]);
            a += density[(size_t)iy * nbx + ix] * area;
            b += density_err[(size_t)iy * nbx + ix] * area;
            integral[(size_t)iy * nbx + ix] = a;
            integral_err[(size_t)iy * nbx + ix] = b;
        }
    } else {                            // along y for each column ix, then apply the boundary condition
        if (t >= nbx) return;
        const int ix = t;
        double a = 0.0, b = 0.0;
        for (int j = 0; j < nby; ++j) {
            const int iy = reverse ? nby - 1 - j : j;
            a += integral[(size_t)iy * nbx + ix];
            b += integral_err[(size_t)iy * nbx + ix];
            integral_err[(size_t)iy * nbx + ix] = b;
            integral[(size_t)iy * nbx + ix] = bc_kind == JMC_BC_LAST_MINUS ? bc_value - a : bc_value + a;
        }
    }
}

}  // namespace jmc

extern "C" int jmc_hist2d(const double* d_x, const double* d_y, const double* d_weights, uint64_t n,
                          const double* d_edges_x, int nx_edges, const double* d_edges_y, int ny_edges,
                          double* d_sum_w, double* d_sum_w2, uint64_t* d_count, void* stream) {
    JMC_REQUIRE(nx_edges >= 2 && ny_edges >= 2, "jmc_hist2d: need at least 2 edges per axis");
    JMC_REQUIRE(d_edges_x && d_edges_y && d_count, "jmc_hist2d: NULL argument");
    JMC_REQUIRE(!d_weights || (d_sum_w && d_sum_w2), "jmc_hist2d: weighted histogram outputs missing");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_x && d_y, "jmc_hist2d: NULL input array");
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    const size_t smem = sizeof(double) * ((size_t)nx_edges + ny_edges);
    if (smem > (size_t)di.max_smem_optin) return set_error(JMC_ELIMIT, "jmc_hist2d: edges exceed shared memory");
    JMC_CUDA(cudaFuncSetAttribute(hist2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_for(n, kThreads);
    hist2d_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(d_x, d_y, d_weights, n, d_edges_x, nx_edges, d_edges_y,
                                                                   ny_edges, d_sum_w, d_sum_w2,
                                                                   (unsigned long long*)d_count);
    JMC_LAUNCHED("hist2d_kernel");
    return JMC_OK;
}

extern "C" int jmc_finalize2d(const double* d_sum_w, const double* d_sum_w2, const uint64_t* d_count,
                              const double* d_edges_x, int nx_edges, const double* d_edges_y, int ny_edges,
                              double area, uint64_t n_total, int bc_kind, double bc_value, double* d_density,
                              double* d_density_err, double* d_integral, double* d_integral_err, void* stream) {
    JMC_REQUIRE(nx_edges >= 2 && ny_edges >= 2, "jmc_finalize2d: need at least 2 edges per axis");
    JMC_REQUIRE(bc_kind >= JMC_BC_FIRST && bc_kind <= JMC_BC_LAST_MINUS, "jmc_finalize2d: invalid boundary condition");
    JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count && d_edges_x && d_edges_y, "jmc_finalize2d: NULL input");
    JMC_REQUIRE(d_density && d_density_err, "jmc_finalize2d: NULL output");
    JMC_REQUIRE((d_integral == nullptr) == (d_integral_err == nullptr), "jmc_finalize2d: integral and its error go together");
    cudaStream_t st = (cudaStream_t)stream;
    const int nbx = nx_edges - 1, nby = ny_edges - 1;
    const int cells = nbx * nby;
    density2d_kernel<<<(cells + 255) / 256, 256, 0, st>>>(d_sum_w, d_sum_w2, (const unsigned long long*)d_count,
                                                           d_edges_x, nx_edges, d_edges_y, ny_edges,
                                                           area / (double)n_total, d_density, d_density_err);
    JMC_LAUNCHED("density2d_kernel");
    if (d_integral) {
        cumsum2d_kernel<<<(nby + 127) / 128, 128, 0, st>>>(d_density, d_density_err, d_edges_x, nx_edges, d_edges_y,
                                                            ny_edges, bc_kind, bc_value, d_integral, d_integral_err, 0);
        JMC_LAUNCHED("cumsum2d_kernel");
        cumsum2d_kernel<<<(nbx + 127) / 128, 128, 0, st>>>(d_density, d_density_err, d_edges_x, nx_edges, d_edges_y,
                                                            ny_edges, bc_kind, bc_value, d_integral, d_integral_err, 1);
        JMC_LAUNCHED("cumsum2d_kernel");
    }
    return JMC_OK;
}

// ---------------------------------------------------------------------------
// fused path, exact (FP64) mode
// ---------------------------------------------------------------------------
namespace jmc {

// COMPACT (running-coupling CRIT / CRIT_SUB): 65 % / 99.7 % of the samples lie below the Landau scale and have weight
// NaN -> 0; the rest need the ~60 FP64 transcendentals of the radiators.  Evaluated in place, 9 % of the CRIT_SUB
// warps would run that code with one or two live lanes -- 70 % of the kernel's time.  Instead a warp appends the
// samples that passed the NaN test to its shared-memory queue (ballot + popc) and evaluates 32 of them at a time with
// every lane busy; the others go straight to the count histogram.  Sums are order-independent up to FP64 rounding,
// as they already are between blocks.
static constexpr int kExactQueueFields = 5, kExactQueueSlots = 64;       // z_c, th_c, z_s, th_s, obs
static constexpr size_t kExactQueueBytes = sizeof(double) * kExactQueueFields * kExactQueueSlots * (kThreads / 32);

template <int KIND, bool ACTIVE, bool COMPACT>
__global__ void __launch_bounds__(kThreads)
run_exact_kernel(RunParams p, uint64_t n, uint64_t seed, uint64_t first, const double* __restrict__ edges,
                 int n_edges, double* __restrict__ g_sum_w, double* __restrict__ g_sum_w2,
                 unsigned long long* __restrict__ g_count, int* g_poison, double* __restrict__ d_minmax, int shift,
                 int queue_offset) {
    extern __shared__ __align__(16) double s_raw[];
    __shared__ int s_poison;
    HistSmem h(s_raw, n_edges > 0 ? n_edges : 1, shift);
    const bool do_hist = edges != nullptr;
    if (threadIdx.x == 0) s_poison = INT_MAX;
    BinGuess G{0, 0.f, 0.f, 0.f};
    if (do_hist) {
        hist_init(h, edges, n_edges);
        G = derive_bin_guess(h.edges, n_edges);
    } else {
        __syncthreads();
    }
    constexpr bool CRIT_RC = KIND == JMC_KIND_CRIT || KIND == JMC_KIND_CRIT_SUB;
    CritConsts consts{0.0, 0.0, 0.0};
    if (CRIT_RC && ACTIVE && !p.g.fixed_coupling) consts = make_crit_consts(p.g.z_cut, p.g.jet);
    const CritConsts* cc = (CRIT_RC && ACTIVE) ? &consts : nullptr;
    double mn = INFINITY, mx = -INFINITY;
    bool has_nan = false;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (!COMPACT) {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            SampleSet s;
            s.z_c = s.th_c = s.z_s = s.th_s = s.z_pre = 0.0;
            s.zb_u = 2.0;
            draw_exact<KIND>(p, first + i, seed, s);
            double jac, obs, wj;
            integrand_eval<KIND, ACTIVE>(p, s, jac, obs, wj, cc, do_hist);
            if (do_hist) hist_add(h, n_edges, obs, wj, true, &s_poison, G);
            if (d_minmax) {
                has_nan |= (obs != obs);
                mn = fmin(mn, obs);
                mx = fmax(mx, obs);
            }
        }
    } else {
        const unsigned lane = threadIdx.x & 31u;
        double* q = s_raw + queue_offset + (threadIdx.x >> 5) * (kExactQueueFields * kExactQueueSlots);
        int qn = 0;                                      // warp-uniform
        auto drain = [&](int slot) {
            const double wj = integrand_heavy_rc<KIND, ACTIVE>(p, q[slot], q[kExactQueueSlots + slot],
                                                               q[2 * kExactQueueSlots + slot], q[3 * kExactQueueSlots + slot], cc);
            hist_add(h, n_edges, q[4 * kExactQueueSlots + slot], wj, true, &s_poison, G);
        };
        // whole warps step together (the queue is warp-collective); lanes past n are idle in the last trip
        for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n; i0 += stride) {
            const uint64_t i = i0 + lane;
            const bool valid = i < n;
            SampleSet s;
            s.z_c = s.th_c = s.z_s = s.th_s = s.z_pre = 0.0;
            s.zb_u = 2.0;
            double obs = 0.0;
            bool active = false;
            if (valid) {
                draw_exact<KIND>(p, first + i, seed, s);
                active = integrand_light_rc<KIND>(p, s, obs);
                if (d_minmax) {
                    has_nan |= (obs != obs);
                    mn = fmin(mn, obs);
                    mx = fmax(mx, obs);
                }
                // certainly NaN: nan_to_num makes it a zero weight (count only), else it poisons like any NaN
                if (do_hist && !active) hist_add(h, n_edges, obs, p.g.nan_to_num ? 0.0 : qnan(), true, &s_poison, G);
            }
            if (!do_hist) continue;
            const unsigned m = __ballot_sync(0xffffffffu, active);
            if (!m) continue;
            if (active) {
                const int slot = qn + __popc(m & ((1u << lane) - 1u));
                q[slot] = s.z_c;
                q[kExactQueueSlots + slot] = s.th_c;
                q[2 * kExactQueueSlots + slot] = s.z_s;
                q[3 * kExactQueueSlots + slot] = s.th_s;
                q[4 * kExactQueueSlots + slot] = obs;
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                drain(qn + (int)lane);
                __syncwarp();
            }
        }
        if (do_hist) {
            __syncwarp();
            if ((int)lane < qn) drain((int)lane);
        }
    }
    if (do_hist) hist_flush(h, n_edges, g_sum_w, g_sum_w2, g_count, &s_poison, g_poison);
    if (d_minmax) minmax_merge(mn, mx, has_nan, d_minmax);
}

template <int KIND>
static int launch_run_exact(const RunParams& p, uint64_t n, uint64_t seed, uint64_t first, const double* d_edges,
                            int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, int* d_poison,
                            double* d_minmax, size_t smem, cudaStream_t st) {
    // active-region evaluation of the running-coupling radiators where its derivation holds (jmc_physics.cuh)
    constexpr bool HAS_RC = KIND == JMC_KIND_CRIT || KIND == JMC_KIND_CRIT_SUB;
    const bool regular = !HAS_RC || p.g.fixed_coupling || (p.g.z_cut < 0.5 && (KIND != JMC_KIND_CRIT_SUB || p.g.beta > 1.0));
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    // (a block histogram close to the shared-memory limit leaves no room for the queues: evaluated in place then)
    const bool compact = HAS_RC && !p.g.fixed_coupling && regular
                         && smem + kExactQueueBytes + 16 <= (size_t)di.max_smem_optin;
    const size_t queue = compact ? kExactQueueBytes : 0;
    const int shift = (d_edges && smem + queue <= 64 * 1024) ? hist_spread_shift(n_edges, 64 * 1024 - queue) : 0;
    if (d_edges) smem = hist_smem_bytes(n_edges, shift);
    const int queue_offset = (int)((smem + 15) / 16 * 2);          // in doubles, 16-byte aligned
    if (compact) smem = (size_t)queue_offset * sizeof(double) + queue;
    auto kern = compact ? run_exact_kernel<KIND, true, HAS_RC>
                        : (regular ? run_exact_kernel<KIND, true, false> : run_exact_kernel<KIND, !HAS_RC, false>);
    JMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    static const int waves = [] {                         // experiment switch: blocks per resident slot
        const char* e = getenv("JMC_EXACT_WAVES");
        return e ? std::max(1, atoi(e)) : 1;
    }();
    const int grid = grid_for(n, kThreads, per_sm * waves);
    if (grid < 0) return JMC_ENODEV;
    kern<<<grid, kThreads, smem, st>>>(p, n, seed, first, d_edges, n_edges, d_sum_w, d_sum_w2,
                                       (unsigned long long*)d_count, d_poison, d_minmax, shift, queue_offset);
    JMC_LAUNCHED("run_exact_kernel");
    return JMC_OK;
}

int run_exact(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, const double* d_edges,
              int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, void* stream) {
    RunParams p;
    int rc = make_run_params(g, &p);
    if (rc) return rc;
    size_t smem = 0;
    if (d_edges) {
        rc = check_hist_args("jmc_run", d_edges, n_edges, &smem);
        if (rc) return rc;
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_run: histogram outputs missing");
    } else {
        JMC_REQUIRE(d_minmax != nullptr, "jmc_run: neither edges nor minmax requested");
    }
    if (n == 0) return JMC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int* d_poison;
    rc = scratch(sizeof(int), (void**)&d_poison, (void*)st);
    if (rc) return rc;
    JMC_CUDA(cudaMemsetAsync(d_poison, 0x7f, sizeof(int), st));
#define JMC_RUN(K) \
    rc = launch_run_exact<K>(p, n, seed, first_index, d_edges, n_edges, d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, smem, st)
    switch (g->kind) {
        case JMC_KIND_RADIATOR: JMC_RUN(JMC_KIND_RADIATOR); break;
        case JMC_KIND_CRIT: JMC_RUN(JMC_KIND_CRIT); break;
        case JMC_KIND_CRIT_SUB: JMC_RUN(JMC_KIND_CRIT_SUB); break;
        case JMC_KIND_PRE_CRIT_SUB: JMC_RUN(JMC_KIND_PRE_CRIT_SUB); break;
        case JMC_KIND_SIMPLE: JMC_RUN(JMC_KIND_SIMPLE); break;
        default: return set_error(JMC_EINVAL, "jmc_run: kind %d has no FP64 sampler form", g->kind);
    }
#undef JMC_RUN
    if (rc) return rc;
    if (d_edges && !g->nan_to_num) {
        poison_kernel<<<1, 256, 0, st>>>(d_poison, n_edges - 1, d_sum_w, d_sum_w2);
        JMC_LAUNCHED("poison_kernel");
    }
    return JMC_OK;
}

// samples on request: per-sample samples / observable / weight*jac of the FP64 path
template <int KIND>
__global__ void __launch_bounds__(kThreads)
dump_exact_kernel(RunParams p, uint64_t n, uint64_t seed, uint64_t first, double* __restrict__ samples,
                  double* __restrict__ obs_out, double* __restrict__ w_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        SampleSet s;
        s.z_c = s.th_c = s.z_s = s.th_s = s.z_pre = 0.0;
        s.zb_u = 2.0;
        draw_exact<KIND>(p, first + i, seed, s);
        double jac, obs, wj;
        integrand_eval<KIND>(p, s, jac, obs, wj);
        if (samples) {
            double* o = samples + 8 * i;
            o[0] = s.z_c; o[1] = s.th_c; o[2] = s.z_s; o[3] = s.th_s; o[4] = s.z_pre; o[5] = s.zb_u;
            o[6] = s.z_c - p.g.z_cut; o[7] = jac;
        }
        if (obs_out) obs_out[i] = obs;
        if (w_out) w_out[i] = wj;
    }
}

int dump_exact(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, double* d_samples,
               double* d_obs, double* d_weight, void* stream) {
    RunParams p;
    int rc = make_run_params(g, &p);
    if (rc) return rc;
    if (n == 0) return JMC_OK;
    const int grid = grid_for(n, kThreads);
    if (grid < 0) return JMC_ENODEV;
    cudaStream_t st = (cudaStream_t)stream;
#define JMC_DUMP(K) dump_exact_kernel<K><<<grid, kThreads, 0, st>>>(p, n, seed, first_index, d_samples, d_obs, d_weight)
    switch (g->kind) {
        case JMC_KIND_RADIATOR: JMC_DUMP(JMC_KIND_RADIATOR); break;
        case JMC_KIND_CRIT: JMC_DUMP(JMC_KIND_CRIT); break;
        case JMC_KIND_CRIT_SUB: JMC_DUMP(JMC_KIND_CRIT_SUB); break;
        case JMC_KIND_PRE_CRIT_SUB: JMC_DUMP(JMC_KIND_PRE_CRIT_SUB); break;
        case JMC_KIND_SIMPLE: JMC_DUMP(JMC_KIND_SIMPLE); break;
        default: return set_error(JMC_EINVAL, "jmc_dump: kind %d has no FP64 sampler form", g->kind);
    }
#undef JMC_DUMP
    JMC_LAUNCHED("dump_exact_kernel");
    return JMC_OK;
}

// ---------------------------------------------------------------------------
// density / cumulative integral epilogue
// ---------------------------------------------------------------------------
static constexpr int kScanThreads = 1024;

// Exclusive prefix of per-thread totals across the block (two arrays at once).
__device__ __forceinline__ void block_exclusive_scan2(double& a, double& b, double* s_a, double* s_b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double ia = a, ib = b;   // inclusive within the warp
    for (int o = 1; o < 32; o <<= 1) {
        const double ta = __shfl_up_sync(0xffffffffu, ia, o);
        const double tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        double wa = s_a[lane], wb = s_b[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const double ta = __shfl_up_sync(0xffffffffu, wa, o);
            const double tb = __shfl_up_sync(0xffffffffu, wb, o);
            if (lane >= o) { wa += ta; wb += tb; }
        }
        s_a[lane] = wa; s_b[lane] = wb;
    }
    __syncthreads();
    const double oa = warp ? s_a[warp - 1] : 0.0, ob = warp ? s_b[warp - 1] : 0.0;
    a = oa + (ia - a);
    b = ob + (ib - b);
}

// One block.  Bins are processed in scan order: forward for a first-bin
// boundary condition, reversed for a last-bin one (integrator.py:162-192).
// FROM_HIST: density and its error are first derived from the histograms
// (integrator.py:113-124); otherwise they are inputs (integrate() alone).
// poison (single job only): the fused kernels' NaN flag -- bins at or above *poison hold NaN sums, as NumPy's
// cumsum-based weighted histogram would (applied here instead of by a kernel of its own; the flag is re-armed)
template <bool FROM_HIST>
__global__ void __launch_bounds__(kScanThreads)
finalize_kernel(const double* __restrict__ sum_w, const double* __restrict__ sum_w2,
                const unsigned long long* __restrict__ count, const double* __restrict__ edges, int n_edges,
                double area, const double* __restrict__ area_per_job, double n_total, int bc_kind, double bc_value,
                double* __restrict__ density, double* __restrict__ density_err, double* __restrict__ integral,
                double* __restrict__ integral_err, int* poison = nullptr) {
    __shared__ double s_a[32], s_b[32];
    const int nb = n_edges - 1;
    if (FROM_HIST && poison) {
        const int pb = *poison;
        __syncthreads();
        if (threadIdx.x == 0) *poison = 0x7f7f7f7f;
        if (pb >= 0 && pb < nb) {
            for (int k = pb + threadIdx.x; k < nb; k += blockDim.x) {
                const_cast<double*>(sum_w)[k] = qnan();
                const_cast<double*>(sum_w2)[k] = qnan();
            }
            __syncthreads();
        }
    }
    // blockIdx.x = job of a batch (jmc_finalize_batch); a single job is a batch of one
    {
        const size_t job = blockIdx.x;
        if (FROM_HIST) { sum_w += job * nb; sum_w2 += job * nb; count += job * nb; }
        edges += job * n_edges;
        density += job * nb; density_err += job * nb;
        if (integral) { integral += job * (nb + 1); integral_err += job * nb; }
        if (area_per_job) area = area_per_job[job];
    }
    const int per = (nb + kScanThreads - 1) / kScanThreads;
    const bool reverse = bc_kind != JMC_BC_FIRST;
    const double norm = area / n_total;
    const int begin = threadIdx.x * per;
    const int end = min(begin + per, nb);
    // pass 1: density and per-thread totals of density*width, err*width
    double ta = 0.0, tb = 0.0;
    for (int j = begin; j < end; ++j) {
        const int k = reverse ? nb - 1 - j : j;
        const double width = edges[k + 1] - edges[k];
        double d, e;
        if (FROM_HIST) {
            d = (sum_w[k] / width) * norm;
            e = (sqrt(sum_w2[k]) / width) * norm;
            if (count[k] == 0ull) { d = 0.0; e = 0.0; }
            density[k] = d;
            density_err[k] = e;
        } else {
            d = density[k];
            e = density_err[k];
        }
        ta += d * width;
        tb += e * width;
    }
    if (integral == nullptr) return;   // density only
    block_exclusive_scan2(ta, tb, s_a, s_b);
    // pass 2: running sums from the thread's offset
    for (int j = begin; j < end; ++j) {
        const int k = reverse ? nb - 1 - j : j;
        const double width = edges[k + 1] - edges[k];
        ta += density[k] * width;
        tb += density_err[k] * width;
        integral_err[k] = tb;
        if (bc_kind == JMC_BC_FIRST) integral[k + 1] = bc_value + ta;
        else if (bc_kind == JMC_BC_LAST_PLUS) integral[k] = bc_value + ta;
        else integral[k] = bc_value - ta;
    }
    if (threadIdx.x == 0) {
        if (bc_kind == JMC_BC_FIRST) integral[0] = bc_value;
        else integral[nb] = bc_value;
    }
}

}  // namespace jmc

extern "C" int jmc_finalize(const double* d_sum_w, const double* d_sum_w2, const uint64_t* d_count,
                            const double* d_edges, int n_edges, double area, uint64_t n_total, int bc_kind,
                            double bc_value, double* d_density, double* d_density_err, double* d_integral,
                            double* d_integral_err, void* stream) {
    JMC_REQUIRE(n_edges >= 2, "jmc_finalize: need at least 2 edges");
    JMC_REQUIRE(n_total > 0, "jmc_finalize: zero samples (the reference divides by len(observables) here)");
    JMC_REQUIRE(bc_kind >= JMC_BC_FIRST && bc_kind <= JMC_BC_LAST_MINUS, "jmc_finalize: invalid boundary condition");
    JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count && d_edges, "jmc_finalize: NULL input");
    JMC_REQUIRE(d_density && d_density_err, "jmc_finalize: NULL output");
    JMC_REQUIRE((d_integral == nullptr) == (d_integral_err == nullptr), "jmc_finalize: integral and its error go together");
    finalize_kernel<true><<<1, kScanThreads, 0, (cudaStream_t)stream>>>(
        d_sum_w, d_sum_w2, (const unsigned long long*)d_count, d_edges, n_edges, area, nullptr, (double)n_total, bc_kind,
        bc_value, d_density, d_density_err, d_integral, d_integral_err);
    JMC_LAUNCHED("finalize_kernel");
    return JMC_OK;
}

namespace jmc {
// jmc_finalize that also applies (and re-arms) the NaN flag of a fused run: one launch less per job
int finalize_poisoned(double* d_sum_w, double* d_sum_w2, const uint64_t* d_count, const double* d_edges, int n_edges,
                      double area, uint64_t n_total, int bc_kind, double bc_value, double* d_density,
                      double* d_density_err, double* d_integral, double* d_integral_err, int* d_poison, void* stream) {
    JMC_REQUIRE(n_total > 0, "jmc_finalize: zero samples (the reference divides by len(observables) here)");
    JMC_REQUIRE(bc_kind >= JMC_BC_FIRST && bc_kind <= JMC_BC_LAST_MINUS, "jmc_finalize: invalid boundary condition");
    finalize_kernel<true><<<1, kScanThreads, 0, (cudaStream_t)stream>>>(
        d_sum_w, d_sum_w2, (const unsigned long long*)d_count, d_edges, n_edges, area, nullptr, (double)n_total, bc_kind,
        bc_value, d_density, d_density_err, d_integral, d_integral_err, d_poison);
    JMC_LAUNCHED("finalize_kernel");
    return JMC_OK;
}
}  // namespace jmc

extern "C" int jmc_finalize_batch(int n_jobs, const double* d_sum_w, const double* d_sum_w2, const uint64_t* d_count,
                                  const double* d_edges, int n_edges, const double* d_area, uint64_t n_total,
                                  int bc_kind, double bc_value, double* d_density, double* d_density_err,
                                  double* d_integral, double* d_integral_err, void* stream) {
    JMC_REQUIRE(n_jobs >= 1, "jmc_finalize_batch: need at least one job");
    JMC_REQUIRE(n_edges >= 2, "jmc_finalize_batch: need at least 2 edges");
    JMC_REQUIRE(bc_kind >= JMC_BC_FIRST && bc_kind <= JMC_BC_LAST_MINUS, "jmc_finalize_batch: invalid boundary condition");
    JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count && d_edges && d_area, "jmc_finalize_batch: NULL input");
    JMC_REQUIRE(d_density && d_density_err && d_integral && d_integral_err, "jmc_finalize_batch: NULL output");
    finalize_kernel<true><<<n_jobs, kScanThreads, 0, (cudaStream_t)stream>>>(
        d_sum_w, d_sum_w2, (const unsigned long long*)d_count, d_edges, n_edges, 0.0, d_area, (double)n_total, bc_kind,
        bc_value, d_density, d_density_err, d_integral, d_integral_err);
    JMC_LAUNCHED("finalize_kernel");
    return JMC_OK;
}

extern "C" int jmc_cumulative(const double* d_density, const double* d_density_err, const double* d_edges,
                              int n_edges, int bc_kind, double bc_value, double* d_integral,
                              double* d_integral_err, void* stream) {
    JMC_REQUIRE(n_edges >= 2, "jmc_cumulative: need at least 2 edges");
    JMC_REQUIRE(bc_kind >= JMC_BC_FIRST && bc_kind <= JMC_BC_LAST_MINUS, "jmc_cumulative: invalid boundary condition");
    JMC_REQUIRE(d_density && d_density_err && d_edges && d_integral && d_integral_err, "jmc_cumulative: NULL argument");
    finalize_kernel<false><<<1, kScanThreads, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, nullptr, d_edges, n_edges, 1.0, nullptr, 1.0, bc_kind, bc_value, (double*)d_density,
        (double*)d_density_err, d_integral, d_integral_err);
    JMC_LAUNCHED("finalize_kernel");
    return JMC_OK;
}
