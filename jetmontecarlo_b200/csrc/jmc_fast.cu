// FP32 log-space mode of the fused generate -> weight -> bin kernel: the
// throughput path.  (Compiled with FMA contraction; jmc_exact.cu is the
// NumPy-order FP64 twin used for sample-by-sample parity.)
//
// Design (B200-first, nothing here exists in the reference, which loops in
// Python over random.random()):
//  * one Philox-4x32-10 block per sample gives four 24-bit uniforms; all
//    log-sampled coordinates are carried as base-2 logarithms
//        l2x = u * log2(eps) + log2(width)
//    so the only transcendental instructions are MUFU.EX2 / MUFU.LG2 on
//    values of moderate size, nothing underflows in FP32 (theta^beta down to
//    1e-40 stays a logarithm), and the log-bin index is an FMA + floor;
//  * weights are evaluated on the ACTIVE branch only.  The reference
//    evaluates every region of the running-coupling radiators and masks them,
//    which yields NaN below the Landau scale (then np.nan_to_num -> 0); here
//    a handful of compares on the logarithms decide "the reference would be
//    NaN" before any radiator arithmetic is done (derivation in DESIGN.md);
//  * jacobian x weight products are assembled in the exponent (one EX2), so
//    the catastrophic cancellation of the reference's (z - z_c) never occurs:
//    delta = z - z_c is the sampled quantity;
//  * histogram: per-block shared-memory bins; counts by native 32-bit shared
//    atomics, weighted sums by FP64 shared atomics issued only for non-zero
//    weights; one flush of global FP64 atomics per block.
//
// ref for the formulas: analytics/radiators/{fixedcoupling,running_coupling}.py,
// numerics/{observables,weights}.py, numerics/radiators/samplers.py (file:line
// next to each function below).
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>
#include "jmc_common.cuh"
#include "jmc_philox.cuh"

namespace jmc {

static constexpr int kFastThreads = 256;
#define JMC_LN2F 0.69314718055994531f
#define JMC_INV_LN2F 1.4426950408889634f

// QCD constants (double on the host, see jmc_physics.cuh for provenance)
static constexpr double H_CF = 8. / 6., H_CA = 3., H_TF = .5, H_NF = 5.;
static constexpr double H_BETA0 = 0.6100939485189322, H_MZ = 91.19, H_PT = 3000., H_AMZ = 0.118;
static constexpr double H_ALPHA = 0.0785101418528313, H_MU = 0.0003333333333333333;
static constexpr double H_PI = 3.141592653589793;

struct FastParams {
    int kind, jet, fixed, obs_mll, split_acc, nan_to_num;
    // sampling, base-2 logs
    float L2;          // log2(eps)
    float l2w_cz;      // log2(1/2 - z_cut)
    float l2R;         // log2(radius) (+ log2(theta_scale) for RADIATOR)
    float l2zc;        // log2(z_cut)
    float zc;
    // observable (C_groomed / C_ungroomed)
    float beta, f, zcl;      // zcl = z_cut - z_pre (scalar z_pre kinds)
    float c0;                // z_cut - f*zcl : delta + c0 = zmin - f*zcl
    float one_m;             // 1 - (1-f)*zcl
    float omf_zcl;           // (1-f)*zcl
    float omf_zc;            // (1-f)*z_cut
    float l2_onorm;          // -2 log2(1 - z_cut)
    // fixed coupling
    float g;                 // 2 C_R alpha_fixed / pi
    float A;                 // ln(1/(2 z_cut)) * g         (theta exponent of critPDF_fc)
    float c1;                // C_R alpha_fixed / (beta pi)
    float ln_zc_eps;         // ln(z_cut / eps)             (zero bin)
    // running coupling, shared
    float ca;                // 2 * 0.118 * beta0
    float lnPT, lnMZ, lnMU, beta0;
    float cn;                // C_R / (pi beta0)            (colour_norm with coeff 2 C_R)
    float inv_2pib0;         // 1 / (2 pi beta0)            (colour_norm with coeff b_bar)
    float twoCR;
    // running coupling, critical radiator (alpha = alpha_fixed)
    float k, lzc, ln2zc, L_r1, L_r2, one_kMU, ln_one_kMU, W_half, W_zc, W_a, S1, S2, H1, cnb;
    int crit_always_nan;
    // running coupling, subsequent radiator
    float lnknee, bbar_knee;
    // binning
    int dom_log;             // 1: compare log2(obs) with log2(edges); 0: compare obs with edges
    int uniform;             // edges equally spaced in that domain
    int n_edges;
    float b0, inv_d;
};

__device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lnf_(float x) { return JMC_LN2F * lg2(x); }
__device__ __forceinline__ float rcp(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float wlog(float x) { return x * lnf_(x); }     // W(x) = x ln x, qcd_utils.py:211-216
__device__ __forceinline__ float fnan() { return __int_as_float(0x7fc00000); }

// z * splitting(z): the jacobian z of log sampling folded in.  ref: analytics/qcd_utils.py:124-190
__device__ __forceinline__ float z_times_p(float z, int jet) {
    const float omz = 1.0f - z;
    if (jet == JMC_QUARK) return (float)H_CF * (1.0f + omz * omz);
    return (float)H_CA * (2.0f * omz + z * z * omz) + (float)(H_TF * H_NF) * z * (z * z + omz * omz);
}
__device__ __forceinline__ float p_of(float z, int jet) { return z_times_p(z, jet) * rcp(z); }
// reduced splitting function P(z) + P(1-z)
__device__ __forceinline__ float p_bar(float z, int jet) { return p_of(z, jet) + p_of(1.0f - z, jet); }

// b_bar(z_c): ref analytics/qcd_utils.py:223-236
__device__ __forceinline__ float b_bar_f(float x, int jet) {
    const float l = lnf_(2.0f - 2.0f * x);
    if (jet == JMC_QUARK) return (float)H_CF * (-3.0f + 6.0f * x + 4.0f * l) * 0.5f;
    const float q = 1.0f + (x - 1.0f) * x;
    return (2.0f * x - 1.0f) * (-4.0f * (float)(H_NF * H_TF) * q + (float)H_CA * (11.0f + 2.0f * (x - 1.0f) * x))
               * (1.0f / 6.0f)
           + 2.0f * (float)H_CA * l;
}

// one-loop coupling from ln(mu/GeV), frozen below 1 GeV.  ref: analytics/qcd_utils.py:102-119
__device__ __forceinline__ float alpha_from_lnmu(const FastParams& P, float ln_mu) {
    return (float)H_AMZ * rcp(1.0f + P.ca * (fmaxf(ln_mu, 0.0f) - P.lnMZ));
}

// critRadAnalytic on the active branch; false = the reference returns NaN.
// L = ln(theta).  ref: analytics/radiators/running_coupling.py:157-289
__device__ __forceinline__ bool crit_rad_rc(const FastParams& P, float L, float& R) {
    const float a3 = 1.0f + P.k * (P.lzc + L);            // 1 + Lambda(z_c theta): the smallest log argument
    if (P.crit_always_nan || !(a3 > 0.0f)) return false;
    const float a2 = 1.0f + P.k * (L - JMC_LN2F);         // 1 + Lambda(theta/2)
    float soft, hard;
    if (L > P.L_r1) {                                     // theta > MU_NP / z_c
        soft = P.cn / P.k * (P.W_half - P.W_zc - wlog(a2) + wlog(a3));
    } else if (L > P.L_r2) {                              // 2 MU_NP < theta < MU_NP / z_c
        const float t = L + P.lzc - P.lnMU;
        soft = P.S1 + P.cn * ((P.W_a - wlog(a2) + P.k * t * (1.0f + P.ln_one_kMU)) / P.k
                              + P.k * t * t / (2.0f * P.one_kMU));
    } else {                                              // theta < 2 MU_NP
        soft = P.S1 + P.S2 + P.cn * (-P.k * P.ln2zc * (P.L_r2 - L) / P.one_kMU);
    }
    if (L > P.L_r2) hard = -P.cnb * lnf_(1.0f + P.k * L);
    else hard = P.H1 + P.cnb * P.k * (P.L_r2 - L) / P.one_kMU;
    R = soft + hard;
    return true;
}

// subRadAnalytic and subRadPrimeAnalytic with maxRadius = theta_crit (per-sample
// alpha) on the active branch.  lc = ln(C / theta_c^beta), Lc = ln(theta_c).
// Outputs R and rp = R' * csub (the 1/c of R' is assembled in the exponent by
// the caller).  false = the reference returns NaN.
// ref: analytics/radiators/running_coupling.py:11-148, 350-380
__device__ __forceinline__ bool sub_rc(const FastParams& P, float lc, float Lc, float& R, float& rp) {
    const float aR = alpha_from_lnmu(P, P.lnPT + Lc);
    const float k = 2.0f * aR * P.beta0;
    const float one_kMU = 1.0f + k * P.lnMU;              // 1 + Lambda(MU_NP, alpha(theta_c))
    const float aC = 1.0f + k * lc;                       // 1 + Lambda(C)
    // log(2 - 2C) in b_bar(C) needs C < 1; the rest are the Landau conditions
    if (!(one_kMU > 0.0f) || !(aC > 0.0f) || !(lc < 0.0f)) return false;
    const float beta = P.beta, ib = rcp(beta), bm1 = beta - 1.0f;
    const float C = ex2(lc * JMC_INV_LN2F);
    const float pref = beta / (k * bm1);
    const float q = 1.0f - k * JMC_LN2F * (1.0f - ib);    // 1 + Lambda(1/2)(1 - 1/beta)
    const float a_half = 1.0f - k * JMC_LN2F;             // 1 + Lambda(1/2)
    const float w1h = wlog(a_half);
    const float ln_one_kMU = lnf_(one_kMU);
    const float aC2 = 1.0f + k * (lc - bm1 * JMC_LN2F) * ib;   // 1 + Lambda(C / 2^(beta-1)) / beta
    auto sc1 = [&](float x) {
        return P.cn * pref * (w1h * (1.0f - ib) - wlog(1.0f + k * (x - bm1 * JMC_LN2F) * ib) + wlog(1.0f + k * x) * ib);
    };
    auto sc2 = [&](float x) {
        const float d = P.lnMU - x;
        return P.cn * (pref * (wlog(q + k * P.lnMU * ib) - wlog(q + k * x * ib) - k * d * ib * (1.0f + ln_one_kMU))
                       + k * d * d / (2.0f * bm1));
    };
    float soft, hard;
    const float bC = b_bar_f(C, P.jet);
    if (lc > P.lnMU) {
        soft = sc1(lc);
    } else if (lc > P.lnknee) {
        soft = sc1(P.lnMU) + sc2(lc);
    } else {
        const float l2c = JMC_LN2F + lc, l2m = JMC_LN2F + P.lnMU;
        soft = sc1(P.lnMU) + sc2(P.lnknee) + P.cn * ((l2c * l2c * ib - beta * l2m * l2m) * k * 0.5f);
    }
    if (lc > P.lnknee) hard = bC * P.inv_2pib0 * (lnf_(a_half) - lnf_(aC2));
    else hard = P.bbar_knee * P.inv_2pib0 * (lnf_(a_half) - ln_one_kMU)
                + bC * P.inv_2pib0 * k * (P.lnknee - lc) / (one_kMU * beta);
    R = (C < 0.5f) ? soft + hard : 0.0f;
    // derivative
    const float lm = fmaxf(lc, fminf((beta * P.lnMU - lc) / bm1, -JMC_LN2F));
    const float sc = lnf_(aC2 / (1.0f + k * (lc + bm1 * lm) * ib)) * beta / (bm1 * k) + (lm - lc) / one_kMU;
    const float hc = rcp(1.0f + k * fmaxf(-bm1 * ib * JMC_LN2F + lc * ib, P.lnMU));
    rp = aR * (P.twoCR * sc + bC * hc) / ((float)H_PI * beta);
    return true;
}

struct FastSample {            // what a dump shows for one sample
    float z_c, th_c, z_s, th_s, z_pre, zb_u, delta;
};

// One sample: returns log2(observable) and weight*jacobian (0 where the
// reference's nan_to_num'd weight is 0; NaN if nan_to_num is off).
template <int KIND>
__device__ __forceinline__ void eval_fast(const FastParams& P, uint64_t idx, uint64_t seed, float& l2obs, float& w,
                                          FastSample* dump) {
    uint32_t r[4];
    philox4x32_10(idx, 0u, seed, r);
    const float u0 = u32_to_unit_float(r[0]), u1 = u32_to_unit_float(r[1]);
    const float u2 = u32_to_unit_float(r[2]), u3 = u32_to_unit_float(r[3]);
    const float nanw = P.nan_to_num ? 0.0f : fnan();

    if (KIND == JMC_KIND_RADIATOR) {
        // ungroomedSampler (samplers.py:204-226), C_ungroomed (observables.py:82-109),
        // radiatorWeight (weights.py:15-26); theta_em = theta * theta_scale (generation.py:303-335)
        const float l2z = u0 * P.L2 - 1.0f;
        const float l2t = u1 * P.L2 + P.l2R;             // log2(theta_em)
        const float z = ex2(l2z);
        l2obs = l2z + P.beta * l2t;
        if (P.obs_mll) l2obs += lg2(1.0f - z);
        float zp;
        if (P.split_acc == JMC_ACC_LL) zp = P.twoCR;
        else if (P.split_acc == JMC_ACC_LL_NONSING) zp = z_times_p(z, P.jet);
        else zp = z_times_p(z, P.jet) + z * p_of(1.0f - z, P.jet);
        const float alpha = P.fixed ? (float)H_ALPHA : alpha_from_lnmu(P, P.lnPT + JMC_LN2F * (l2z + l2t));
        w = zp * alpha * (float)(1.0 / H_PI);            // splitting * alpha / (theta pi) * z theta
        if (dump) { dump->z_s = z; dump->th_s = ex2(l2t); }
        return;
    }

    // critical emission: z = z_c + delta, delta log-sampled (samplers.py:38-62)
    const float l2d = u0 * P.L2 + P.l2w_cz;
    const float l2tc = u1 * P.L2 + P.l2R;
    const float delta = ex2(l2d);
    const float z = P.zc + delta;
    const float l2z = lg2(z);
    const float Lc = JMC_LN2F * l2tc;
    if (dump) { dump->z_c = z; dump->th_c = ex2(l2tc); dump->delta = delta; }

    if (KIND == JMC_KIND_CRIT) {
        // C_groomed (observables.py:15-79), criticalEmissionWeight (weights.py:31-41)
        const float a = delta + P.c0;
        const float b = P.obs_mll ? (1.0f - z) - P.omf_zcl : P.one_m;
        l2obs = lg2(a * b) + P.beta * l2tc + P.l2_onorm;
        if (P.fixed) {
            w = P.g * ex2(l2d - l2z + P.A * l2tc);       // 2 C_R a/(pi theta z) theta^A * (z - z_c) theta
        } else {
            float R;
            if (!crit_rad_rc(P, Lc, R)) { w = nanw; return; }
            const float alpha = alpha_from_lnmu(P, P.lnPT + JMC_LN2F * l2z + Lc);
            w = p_bar(z, P.jet) * delta * alpha * (float)(1.0 / H_PI) * ex2(-R * JMC_INV_LN2F);
        }
        return;
    }

    // subsequent emission from the ungroomed sampler
    const float l2zs = u2 * P.L2 - 1.0f;
    const float l2ts = u3 * P.L2 + P.l2R;
    if (dump) { dump->z_s = ex2(l2zs); dump->th_s = ex2(l2ts); }

    if (KIND == JMC_KIND_CRIT_SUB) {
        // twoEmissionWeight (weights.py:63-99) with csub = z_s th_s^b th_s^b as written at :74
        const float a = delta + P.c0;
        const float b = P.obs_mll ? (1.0f - z) - P.omf_zcl : P.one_m;
        const float l2cc = lg2(a * b) + P.beta * l2tc + P.l2_onorm;
        const float l2cs = l2zs + 2.0f * P.beta * l2ts;
        l2obs = fmaxf(l2cc, l2cs);
        const float lc = JMC_LN2F * (l2cs - P.beta * l2tc);          // ln(csub / theta_c^beta)
        if (P.fixed) {
            // critPDF_fc * subPDF_fc(csub, maxRadius=theta_c) * th_s^b th_c^b * jacobians
            const float x = JMC_LN2F + lc;                           // ln(2 csub / theta_c^beta)
            if (!(x < 0.0f)) { w = 0.0f; return; }
            const float e2 = l2d - l2z + (P.A + P.beta) * l2tc + (1.0f - P.beta) * l2ts - P.c1 * x * x * JMC_INV_LN2F;
            w = P.g * 2.0f * P.c1 * (-x) * ex2(e2);
        } else {
            float Rc, Rs, rp;
            if (!sub_rc(P, lc, Lc, Rs, rp) || !crit_rad_rc(P, Lc, Rc)) { w = nanw; return; }
            const float alpha = alpha_from_lnmu(P, P.lnPT + JMC_LN2F * l2z + Lc);
            const float lin = p_bar(z, P.jet) * delta * alpha * (float)(1.0 / H_PI) * rp;
            w = lin * ex2(P.beta * l2tc + (1.0f - P.beta) * l2ts - (Rc + Rs) * JMC_INV_LN2F);
        }
        return;
    }

    if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        // all-emission integrand with the zero bin
        // (examples/tests/comparison_tests/comparison_utils_probabilistic.py:82-98, 384-474)
        uint32_t r2[4];
        philox4x32_10(idx, 1u, seed, r2);
        const float u4 = u32_to_unit_float(r2[0]), u5 = u32_to_unit_float(r2[1]);
        const float E = -P.g * Lc;                                   // 2 C_R a/pi ln(R0/theta_c)
        const float cdf = ex2(-E * P.ln_zc_eps * JMC_INV_LN2F);
        const float l2zp = u4 * P.L2 + P.l2zc;
        const bool zero_bin = u5 < cdf;
        const float zp = zero_bin ? 0.0f : ex2(l2zp);
        const float zcl = P.zc - zp;
        // zmin - f (z_cut - z_pre) = delta + (1-f) z_cut + f z_pre: summed without passing through z_cut - z_pre,
        // which would swallow a z_pre below the FP32 spacing of z_cut
        const float a = delta + P.omf_zc + P.f * zp;
        const float b = P.obs_mll ? (1.0f - z) - (1.0f - P.f) * zcl : 1.0f - (1.0f - P.f) * zcl;
        const float l2cc = lg2(a * b) + P.beta * l2tc + P.l2_onorm;
        const float l2cs = l2zs + P.beta * l2tc;                     // c_sub = z_s theta_c^beta
        l2obs = fmaxf(l2cc, l2cs);
        const float x = JMC_LN2F * (1.0f + l2zs);                    // ln(2 z_s) < 0
        const float pre = zero_bin ? cdf : E * ex2(-E * (P.l2zc - l2zp));
        const float e2 = l2d - l2z + P.A * l2tc + l2ts - P.c1 * x * x * JMC_INV_LN2F;
        w = P.g * 2.0f * P.c1 * (-x) * pre * ex2(e2);
        if (dump) { dump->z_pre = zp; dump->zb_u = u5; }
        return;
    }
}

// ---- binning ---------------------------------------------------------------
// s_edges: edges in the compare domain (log2 or linear), FP32.
__device__ __forceinline__ int bin_fast(const FastParams& P, float l2obs, const float* __restrict__ s_edges) {
    const float v = P.dom_log ? l2obs : ex2(l2obs);
    const int ne = P.n_edges;
    if (!(v >= s_edges[0]) || !(v <= s_edges[ne - 1])) return -1;
    int idx;
    if (P.uniform) {
        idx = (int)((v - P.b0) * P.inv_d);
        idx = max(0, min(idx, ne - 2));
        // FP32 rounding of the affine map can be off by one: settle against the edges themselves
        if (v < s_edges[idx]) --idx;
        else if (idx < ne - 2 && v >= s_edges[idx + 1]) ++idx;
    } else {
        int lo = 0, hi = ne;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_edges[mid] <= v) lo = mid + 1; else hi = mid;
        }
        idx = min(lo - 1, ne - 2);
    }
    return idx;
}

template <int KIND>
__global__ void __launch_bounds__(kFastThreads)
run_fast_kernel(FastParams P, uint64_t n, uint64_t seed, uint64_t first, const float* __restrict__ g_edges,
                double* __restrict__ g_sum_w, double* __restrict__ g_sum_w2,
                unsigned long long* __restrict__ g_count, int* g_poison, double* __restrict__ d_minmax) {
    extern __shared__ double s_raw[];
    __shared__ int s_poison;
    const int ne = P.n_edges, nb = ne > 0 ? ne - 1 : 0;
    double* s_sum_w = s_raw;
    double* s_sum_w2 = s_sum_w + nb;
    unsigned int* s_count = (unsigned int*)(s_sum_w2 + nb);
    float* s_edges = (float*)(s_count + nb);
    const bool do_hist = g_edges != nullptr;
    if (threadIdx.x == 0) s_poison = INT_MAX;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) { s_sum_w[k] = 0.0; s_sum_w2[k] = 0.0; s_count[k] = 0u; }
    for (int k = threadIdx.x; k < ne; k += blockDim.x) s_edges[k] = g_edges[k];
    __syncthreads();
    float mn = INFINITY, mx = -INFINITY;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float l2obs, w;
        eval_fast<KIND>(P, first + i, seed, l2obs, w, nullptr);
        if (do_hist) {
            const int b = bin_fast(P, l2obs, s_edges);
            if (w != w) {                                // NaN weight (nan_to_num off): NumPy poisons bins >= b
                const float v = P.dom_log ? l2obs : ex2(l2obs);
                const int pb = b >= 0 ? b : (v < s_edges[0] ? 0 : INT_MAX);
                if (pb != INT_MAX) atomicMin(&s_poison, pb);
                if (b >= 0) atomicAdd(&s_count[b], 1u);
            } else if (b >= 0) {
                atomicAdd(&s_count[b], 1u);
                if (w != 0.0f) {
                    const double wd = (double)w;
                    atomicAdd(&s_sum_w[b], wd);
                    atomicAdd(&s_sum_w2[b], wd * wd);
                }
            }
        }
        mn = fminf(mn, l2obs);
        mx = fmaxf(mx, l2obs);
    }
    __syncthreads();
    if (do_hist) {
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            const unsigned int c = s_count[k];
            if (c) {
                atomicAdd(&g_count[k], (unsigned long long)c);
                const double sw = s_sum_w[k], sw2 = s_sum_w2[k];
                if (sw != 0.0) atomicAdd(&g_sum_w[k], sw);
                if (sw2 != 0.0) atomicAdd(&g_sum_w2[k], sw2);
            }
        }
        if (threadIdx.x == 0 && s_poison != INT_MAX) atomicMin(g_poison, s_poison);
    }
    if (d_minmax) {
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((threadIdx.x & 31) == 0 && mn <= mx) {
            // merge exp2 of the log-domain extremes into the caller's [min, max]
            const double dmn = exp2((double)mn), dmx = exp2((double)mx);
            unsigned long long* a0 = (unsigned long long*)&d_minmax[0];
            unsigned long long* a1 = (unsigned long long*)&d_minmax[1];
            unsigned long long old = *a0;
            while (dmn < __longlong_as_double(old)) {
                const unsigned long long as = old;
                old = atomicCAS(a0, as, (unsigned long long)__double_as_longlong(dmn));
                if (old == as) break;
            }
            old = *a1;
            while (dmx > __longlong_as_double(old)) {
                const unsigned long long as = old;
                old = atomicCAS(a1, as, (unsigned long long)__double_as_longlong(dmx));
                if (old == as) break;
            }
        }
    }
}

__global__ void poison_fast_kernel(const int* g_poison, int nb, double* g_sum_w, double* g_sum_w2) {
    const int pb = *g_poison;
    if (pb < 0 || pb >= nb) return;
    for (int k = pb + blockIdx.x * blockDim.x + threadIdx.x; k < nb; k += gridDim.x * blockDim.x) {
        g_sum_w[k] = __longlong_as_double(0x7ff8000000000000LL);
        g_sum_w2[k] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

// samples on request: per-sample samples / observable / weight of the FP32 path
template <int KIND>
__global__ void __launch_bounds__(kFastThreads)
dump_fast_kernel(FastParams P, uint64_t n, uint64_t seed, uint64_t first, double* __restrict__ samples,
                 double* __restrict__ obs, double* __restrict__ weight) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        FastSample s = {0.f, 0.f, 0.f, 0.f, 0.f, 2.f, 0.f};
        float l2obs, w;
        eval_fast<KIND>(P, first + i, seed, l2obs, w, &s);
        if (samples) {
            double* o = samples + 8 * i;
            o[0] = s.z_c; o[1] = s.th_c; o[2] = s.z_s; o[3] = s.th_s; o[4] = s.z_pre; o[5] = s.zb_u;
            o[6] = s.delta; o[7] = 0.0;
        }
        if (obs) obs[i] = exp2((double)l2obs);
        if (weight) weight[i] = (double)w;
    }
}

// ---- host: derive FastParams in double -------------------------------------
static double h_wlog(double x) { return x * std::log(x); }
static double h_bbar(double zc, int jet) {
    if (jet == JMC_QUARK) return H_CF * (-3. + 6. * zc + 4. * std::log(2. - 2. * zc)) / 2.;
    return ((-1. + 2. * zc) * (-4. * H_NF * H_TF * (1. + (-1. + zc) * zc) + H_CA * (11. + 2. * (-1. + zc) * zc)) / 6.
            + 2. * H_CA * std::log(2. - 2. * zc));
}

static int make_fast_params(const jmc_integrand* g, const double* h_edges, int n_edges, FastParams* out,
                            std::vector<float>* edges_f) {
    RunParams rp;
    int rc = make_run_params(g, &rp);
    if (rc) return rc;
    if (g->method != JMC_LOG)
        return set_error(JMC_EINVAL, "jmc_run: the FP32 path is the log-sampling path; use JMC_F64 for 'lin' sampling");
    if (g->kind == JMC_KIND_SIMPLE)
        return set_error(JMC_EINVAL, "jmc_run: simpleSampler integrands run in JMC_F64");
    if (!g->fixed_coupling && !(g->beta > 1.0) && g->kind == JMC_KIND_CRIT_SUB)
        return set_error(JMC_EINVAL, "jmc_run: running-coupling subsequent radiator needs beta > 1 (beta = 1 is singular "
                                     "in the reference as well)");
    FastParams P;
    std::memset(&P, 0, sizeof(P));
    const double ln2 = std::log(2.0);
    const double CR = g->jet == JMC_QUARK ? H_CF : H_CA;
    P.kind = g->kind; P.jet = g->jet; P.fixed = g->fixed_coupling; P.obs_mll = g->obs_acc == JMC_ACC_MLL;
    P.split_acc = g->split_acc; P.nan_to_num = g->nan_to_num;
    P.L2 = (float)std::log2(g->epsilon);
    P.l2w_cz = (float)std::log2(0.5 - g->z_cut);
    P.l2R = (float)(std::log2(g->radius) + (g->kind == JMC_KIND_RADIATOR ? std::log2(g->theta_scale) : 0.0));
    P.l2zc = g->z_cut > 0 ? (float)std::log2(g->z_cut) : 0.f;
    P.zc = (float)g->z_cut;
    P.beta = (float)g->beta; P.f = (float)g->f_soft;
    const double zcl = g->z_cut - g->z_pre;
    P.zcl = (float)zcl;
    P.c0 = (float)(g->z_cut - g->f_soft * zcl);
    P.one_m = (float)(1.0 - (1.0 - g->f_soft) * zcl);
    P.omf_zcl = (float)((1.0 - g->f_soft) * zcl);
    P.omf_zc = (float)((1.0 - g->f_soft) * g->z_cut);
    P.l2_onorm = (float)(-2.0 * std::log2(1.0 - g->z_cut));
    const double gg = 2.0 * CR * H_ALPHA / H_PI;
    P.g = (float)gg;
    P.A = g->z_cut > 0 ? (float)(std::log(1.0 / (2.0 * g->z_cut)) * gg) : 0.f;
    P.c1 = (float)(CR * H_ALPHA / (g->beta * H_PI));
    P.ln_zc_eps = g->z_cut > 0 ? (float)std::log(g->z_cut / g->epsilon) : 0.f;
    P.ca = (float)(2.0 * H_AMZ * H_BETA0);
    P.lnPT = (float)std::log(H_PT); P.lnMZ = (float)std::log(H_MZ); P.lnMU = (float)std::log(H_MU);
    P.beta0 = (float)H_BETA0;
    P.cn = (float)(CR / (H_PI * H_BETA0));
    P.inv_2pib0 = (float)(1.0 / (2.0 * H_PI * H_BETA0));
    P.twoCR = (float)(2.0 * CR);
    if (g->z_cut > 0) {
        const double k = 2.0 * H_ALPHA * H_BETA0, lzc = std::log(g->z_cut), lnMU = std::log(H_MU);
        const double cn = CR / (H_PI * H_BETA0), cnb = h_bbar(g->z_cut, g->jet) / (2.0 * H_PI * H_BETA0);
        const double one_kMU = 1.0 + k * lnMU;
        P.k = (float)k; P.lzc = (float)lzc; P.ln2zc = (float)std::log(2.0 * g->z_cut);
        const double L1 = lnMU - lzc, L2 = lnMU + ln2;
        P.L_r1 = (float)L1; P.L_r2 = (float)L2;
        P.one_kMU = (float)one_kMU; P.ln_one_kMU = (float)std::log(one_kMU);
        const double W_half = h_wlog(1.0 - k * ln2), W_zc = h_wlog(1.0 + k * lzc);
        const double a_Wa = 1.0 + k * (lnMU - ln2 - lzc);
        const double W_a = h_wlog(a_Wa);
        auto sc1 = [&](double L) {
            return cn / k * (W_half - W_zc - h_wlog(1.0 + k * (L - ln2)) + h_wlog(1.0 + k * (lzc + L)));
        };
        auto sc2 = [&](double L) {
            const double t = L + lzc - lnMU;
            return cn * ((W_a - h_wlog(1.0 + k * (L - ln2)) + k * t * (1.0 + std::log(one_kMU))) / k
                         + k * t * t / (2.0 * one_kMU));
        };
        P.W_half = (float)W_half; P.W_zc = (float)W_zc; P.W_a = (float)W_a;
        P.S1 = (float)sc1(L1); P.S2 = (float)sc2(L2);
        P.H1 = (float)(-cnb * std::log(1.0 + k * L2));
        P.cnb = (float)cnb;
        // constants that are NaN make the reference NaN everywhere (tiny z_cut)
        P.crit_always_nan = !(1.0 + k * lzc > 0.0) || !(a_Wa > 0.0) || !(one_kMU > 0.0) || !(1.0 + k * L2 > 0.0)
                            || !std::isfinite(P.S1) || !std::isfinite(P.S2);
    }
    const double knee = std::pow(H_MU, g->beta) * std::pow(2.0, g->beta - 1.0);
    P.lnknee = (float)std::log(knee);
    P.bbar_knee = (float)h_bbar(knee, g->jet);
    // ---- binning domain
    P.n_edges = n_edges;
    if (h_edges) {
        bool positive = true, increasing = true;
        for (int k = 0; k < n_edges; ++k) {
            positive = positive && h_edges[k] > 0.0;
            if (k) increasing = increasing && h_edges[k] > h_edges[k - 1];
        }
        if (!increasing) return set_error(JMC_EINVAL, "jmc_run: bin edges must increase strictly");
        P.dom_log = positive;
        edges_f->resize(n_edges);
        std::vector<double> dom(n_edges);
        for (int k = 0; k < n_edges; ++k) dom[k] = positive ? std::log2(h_edges[k]) : h_edges[k];
        const double d = (dom[n_edges - 1] - dom[0]) / (n_edges - 1);
        bool uniform = true;
        for (int k = 0; k < n_edges; ++k) uniform = uniform && std::fabs(dom[k] - (dom[0] + k * d)) <= 1e-9 * std::fabs(d) * n_edges + 1e-12;
        P.uniform = uniform;
        P.b0 = (float)dom[0];
        P.inv_d = (float)(1.0 / d);
        for (int k = 0; k < n_edges; ++k) (*edges_f)[k] = (float)dom[k];
        for (int k = 1; k < n_edges; ++k)
            if (!((*edges_f)[k] > (*edges_f)[k - 1]))
                return set_error(JMC_ELIMIT, "jmc_run: bin edges %d and %d coincide in FP32; use JMC_F64 for bins this fine",
                                 k - 1, k);
    }
    *out = P;
    return JMC_OK;
}

static size_t fast_smem_bytes(int n_edges) {
    const size_t nb = n_edges > 0 ? (size_t)n_edges - 1 : 0;
    return sizeof(double) * 2 * nb + sizeof(unsigned int) * nb + sizeof(float) * (size_t)(n_edges > 0 ? n_edges : 0);
}

template <int KIND>
static int launch_fast(const FastParams& P, uint64_t n, uint64_t seed, uint64_t first, const float* d_edges_f,
                       double* d_sum_w, double* d_sum_w2, uint64_t* d_count, int* d_poison, double* d_minmax,
                       cudaStream_t st) {
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    const size_t smem = fast_smem_bytes(d_edges_f ? P.n_edges : 0);
    if (smem > (size_t)di.max_smem_optin)
        return set_error(JMC_ELIMIT, "jmc_run: %d edges need %zu B of shared memory (limit %d B)", P.n_edges, smem,
                         di.max_smem_optin);
    JMC_CUDA(cudaFuncSetAttribute(run_fast_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, run_fast_kernel<KIND>, kFastThreads, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t want = (n + kFastThreads - 1) / kFastThreads;
    const uint64_t cap = (uint64_t)di.sm_count * per_sm;      // persistent: one wave of resident blocks
    const int grid = (int)(want < cap ? want : cap);
    FastParams Q = P;
    if (!d_edges_f) Q.n_edges = 0;
    run_fast_kernel<KIND><<<grid, kFastThreads, smem, st>>>(Q, n, seed, first, d_edges_f, d_sum_w, d_sum_w2,
                                                             (unsigned long long*)d_count, d_poison, d_minmax);
    JMC_LAUNCHED("run_fast_kernel");
    return JMC_OK;
}

int run_fast(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, const double* d_edges,
             int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (d_edges) {
        JMC_REQUIRE(n_edges >= 2, "jmc_run: need at least 2 edges");
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_run: histogram outputs missing");
    } else {
        JMC_REQUIRE(d_minmax != nullptr, "jmc_run: neither edges nor minmax requested");
    }
    // The FP32 kernel bins against FP32 edges in its compare domain (log2 or linear); they are derived on the
    // host from the caller's FP64 edges: one synchronous 8*n_edges byte read-back per launch.
    std::vector<double> h_edges;
    std::vector<float> edges_f;
    if (d_edges) {
        h_edges.resize(n_edges);
        JMC_CUDA(cudaMemcpyAsync(h_edges.data(), d_edges, sizeof(double) * n_edges, cudaMemcpyDeviceToHost, st));
        JMC_CUDA(cudaStreamSynchronize(st));
    }
    FastParams P;
    int rc = make_fast_params(g, d_edges ? h_edges.data() : nullptr, d_edges ? n_edges : 0, &P, &edges_f);
    if (rc) return rc;
    if (n == 0) return JMC_OK;
    char* scr;
    rc = scratch(256 + sizeof(float) * (size_t)(n_edges > 0 ? n_edges : 1), (void**)&scr);
    if (rc) return rc;
    int* d_poison = (int*)scr;
    float* d_edges_f = nullptr;
    JMC_CUDA(cudaMemsetAsync(d_poison, 0x7f, sizeof(int), st));
    if (d_edges) {
        d_edges_f = (float*)(scr + 256);
        JMC_CUDA(cudaMemcpyAsync(d_edges_f, edges_f.data(), sizeof(float) * n_edges, cudaMemcpyHostToDevice, st));
        JMC_CUDA(cudaStreamSynchronize(st));     // edges_f goes out of scope at return
    }
#define JMC_RUNF(K) rc = launch_fast<K>(P, n, seed, first_index, d_edges_f, d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, st)
    switch (g->kind) {
        case JMC_KIND_RADIATOR: JMC_RUNF(JMC_KIND_RADIATOR); break;
        case JMC_KIND_CRIT: JMC_RUNF(JMC_KIND_CRIT); break;
        case JMC_KIND_CRIT_SUB: JMC_RUNF(JMC_KIND_CRIT_SUB); break;
        case JMC_KIND_PRE_CRIT_SUB: JMC_RUNF(JMC_KIND_PRE_CRIT_SUB); break;
        default: return set_error(JMC_EINVAL, "jmc_run: kind %d has no FP32 form", g->kind);
    }
#undef JMC_RUNF
    if (rc) return rc;
    if (d_edges && !g->nan_to_num) {
        poison_fast_kernel<<<1, 256, 0, st>>>(d_poison, n_edges - 1, d_sum_w, d_sum_w2);
        JMC_LAUNCHED("poison_fast_kernel");
    }
    return JMC_OK;
}

int dump_fast(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, double* d_samples,
              double* d_obs, double* d_weight, void* stream) {
    FastParams P;
    std::vector<float> unused;
    int rc = make_fast_params(g, nullptr, 0, &P, &unused);
    if (rc) return rc;
    if (n == 0) return JMC_OK;
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    uint64_t want = (n + kFastThreads - 1) / kFastThreads;
    const uint64_t cap = (uint64_t)di.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = (cudaStream_t)stream;
#define JMC_DUMPF(K) dump_fast_kernel<K><<<grid, kFastThreads, 0, st>>>(P, n, seed, first_index, d_samples, d_obs, d_weight)
    switch (g->kind) {
        case JMC_KIND_RADIATOR: JMC_DUMPF(JMC_KIND_RADIATOR); break;
        case JMC_KIND_CRIT: JMC_DUMPF(JMC_KIND_CRIT); break;
        case JMC_KIND_CRIT_SUB: JMC_DUMPF(JMC_KIND_CRIT_SUB); break;
        case JMC_KIND_PRE_CRIT_SUB: JMC_DUMPF(JMC_KIND_PRE_CRIT_SUB); break;
        default: return set_error(JMC_EINVAL, "jmc_dump: kind %d has no FP32 form", g->kind);
    }
#undef JMC_DUMPF
    JMC_LAUNCHED("dump_fast_kernel");
    return JMC_OK;
}

}  // namespace jmc
