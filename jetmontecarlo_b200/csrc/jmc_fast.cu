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
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include <atomic>
#include "jmc_common.cuh"
#include "jmc_philox.cuh"

namespace jmc {

static constexpr int kFastThreads = 256, kFastThreadsBig = 1024;
static constexpr size_t kBigSmemThreshold = 72 * 1024;     // block histograms above this use the 1024-thread kernels
static constexpr int kQueueFields = 5, kQueueSlots = 64;   // per-warp compaction queue (see run_fast_kernel)
// resident blocks per SM the register allocation must allow (3 x 256 threads: <= 85 registers) and sample pairs per
// loop trip; swept with scripts/fast_shootout.sh.
// The rarely executed running-coupling radiator code may spill; the per-sample light path must not.
#ifndef JMC_FAST_MIN_BLOCKS
#define JMC_FAST_MIN_BLOCKS 3
#endif
#ifndef JMC_FAST_PAIRS
#define JMC_FAST_PAIRS 1         // sample pairs per loop trip, kinds with >= 4 uniforms per sample
#endif
#ifndef JMC_FAST_PAIRS_COMPACT
#define JMC_FAST_PAIRS_COMPACT 2 // the same for the warp-compacted kinds (running-coupling critical x subsequent)
#endif
#ifndef JMC_FAST_PAIRS_2W
#define JMC_FAST_PAIRS_2W 2      // sample pairs per loop trip, kinds with 2 uniforms per sample (one Philox block per pair)
#endif
#ifndef JMC_FAST_PIN_KEYS
#define JMC_FAST_PIN_KEYS 1      // keep the Philox round keys in registers across the loop
#endif
#ifndef JMC_FAST_CHUNK
#define JMC_FAST_CHUNK 32        // units (warp trips) a warp takes per request from the global work counter
#endif
#ifndef JMC_ACC_MODE
#define JMC_ACC_MODE 3           // shared-memory accumulation of (sum w, sum w^2), see fast_body::accumulate
#endif
// accumulator types of the block histogram per mode: 0 / 2 / 3 FP64 both, 4 FP32 both, 5 FP64 sum w + FP32 sum w^2
#if JMC_ACC_MODE == 4
typedef float acc_w_t;
typedef float acc_w2_t;
#elif JMC_ACC_MODE == 5
typedef double acc_w_t;
typedef float acc_w2_t;
#else
typedef double acc_w_t;
typedef double acc_w2_t;
#endif
#define JMC_LN2F 0.69314718055994531f
#define JMC_INV_LN2F 1.4426950408889634f

// QCD constants (double on the host, see jmc_physics.cuh for provenance)
static constexpr double H_CF = 8. / 6., H_CA = 3., H_TF = .5, H_NF = 5.;
static constexpr double H_BETA0 = 0.6100939485189322, H_MZ = 91.19, H_PT = 3000., H_AMZ = 0.118;
static constexpr double H_ALPHA = 0.0785101418528313, H_MU = 0.0003333333333333333;
static constexpr double H_PI = 3.141592653589793;

struct FastParams {
    int kind, jet, fixed, obs_mll, split_acc, nan_to_num, count_weighted;
    // sampling, base-2 logs
    float L2;          // log2(eps)
    float sL2;         // 2^-32 log2(eps): maps a 32-bit word straight to u * log2(eps)
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
    float aK, c0K;           // (1 - z_cut)^-2 and c0 (1 - z_cut)^-2: the observable's normalisation folded into `a`
    float bm0;               // 1 - z_cut - (1-f)*zcl : b = bm0 - delta (MLL observable)
    float omb;               // 1 - beta
    float act_c2, act_c3;    // thresholds of the "1 + Lambda(C, alpha(theta_c)) > 0" test in base-2 logs
    float wT1, wT2, wT3, wT4; // the same tests on float(random word), see active_from_words()
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
    float l2tc_min;          // log2 of the smallest theta_c at which the CRIT_SUB weight is not NaN
    // binning
    int dom_log;             // 1: compare log2(obs) with log2(edges); 0: compare obs with edges
    int uniform;             // edges equally spaced in that domain
    int n_edges;
    float b0, inv_d;
    float mb0_inv_d;         // -b0 * inv_d
    float e_first, e_last;   // first / last edge in the compare domain
    float nb_f;              // number of bins as a float
    int spread_shift;        // log2 of the number of lane-interleaved copies of each bin's accumulator
    int checked_only;        // small job: every trip through the checked loop (see fast_body)
};

__device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lnf_(float x) { return JMC_LN2F * lg2(x); }
__device__ __forceinline__ float rcp(float x) { return __fdividef(1.0f, x); }
// ln(1 + t), accurate for small |t| where lg2.approx(1 + t) only has absolute accuracy
__device__ __forceinline__ float log1p_f(float t) {
    if (fabsf(t) < 0.03f) return t * (1.0f - t * (0.5f - t * (0.33333334f - t * 0.25f)));
    return lnf_(1.0f + t);
}
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
    // ln(aC2 / aM) with aM = 1 + k (lc + (beta-1) lm)/beta and aC2 - aM = -k (beta-1)(ln2 + lm)/beta formed
    // directly: at m = 1/2 the ratio is exactly 1 and R' changes sign nearby, so no cancellation is allowed here
    const float aM = 1.0f + k * (lc + bm1 * lm) * ib;
    const float sc = log1p_f(-k * bm1 * (JMC_LN2F + lm) * ib / aM) * beta / (bm1 * k) + (lm - lc) / one_kMU;
    const float hc = rcp(1.0f + k * fmaxf(-bm1 * ib * JMC_LN2F + lc * ib, P.lnMU));
    rp = aR * (P.twoCR * sc + bC * hc) / ((float)H_PI * beta);
    return true;
}

// The draws of one sample, kept in registers between the two stages below.
struct Draw {
    float l2d, l2tc, l2zs, l2ts, l2zp;   // base-2 logs of delta, theta_c, z_s, theta_s, z_pre
    float delta, z;                      // delta = z_crit - z_cut (sampled), z = z_cut + delta
    float u5, cdf;                       // zero-bin uniform and zero-bin probability (PRE_CRIT_SUB)
    float lcl2;                          // log2(csub / theta_c^beta)
};

// uniform in [0,1] from a 32-bit word folded into the log-sampling FMA:
// l2x = u * log2(eps) + log2(width) = float(w) * (2^-32 log2 eps) + log2(width)
__device__ __forceinline__ float l2_sample(uint32_t w, float sL2, float l2w) {
    return fmaf(__uint2float_rn(w), sL2, l2w);
}

// ---- sample streams of the FP32 mode -------------------------------------------------------------------------
// Pair p = (global sample index >> 1) owns the Philox blocks (counter = (p_lo, p_hi, slot, 0)); sample s = index & 1:
//   2-uniform kinds (RADIATOR, CRIT):  slot 0, words 2s, 2s+1                       one block per TWO samples
//   4-uniform kind (CRIT_SUB):         slot s, words 0..3
//   6-uniform kind (PRE_CRIT_SUB):     slot s, words 0..3 + slot 2, words 2s, 2s+1   three blocks per two samples
// A thread always evaluates both samples of a pair, so no generated word is thrown away.
template <int KIND>
struct PairWords {
    static constexpr bool FOUR = KIND == JMC_KIND_CRIT_SUB || KIND == JMC_KIND_PRE_CRIT_SUB;
    uint32_t a[4];
    uint32_t b[FOUR ? 4 : 1];
    uint32_t c[KIND == JMC_KIND_PRE_CRIT_SUB ? 4 : 1];
};
struct Words {
    uint32_t r[4];       // critical (z, theta) then subsequent (z, theta); 2-uniform kinds use r[0], r[1]
    uint32_t e[2];       // pre-critical z, zero-bin uniform
};
template <int KIND>
__device__ __forceinline__ void draw_pair(uint32_t p_lo, uint32_t p_hi, const PhiloxKeys& keys, PairWords<KIND>& w) {
    philox4x32_10(p_lo, p_hi, 0u, keys, w.a);
    if (PairWords<KIND>::FOUR) philox4x32_10(p_lo, p_hi, 1u, keys, w.b);
    if (KIND == JMC_KIND_PRE_CRIT_SUB) philox4x32_10(p_lo, p_hi, 2u, keys, w.c);
}
// words of sample s of the pair (s is a compile-time constant in the hot loop: no register indexing)
template <int KIND>
__device__ __forceinline__ Words sample_words(const PairWords<KIND>& pw, int s) {
    Words w;
    if (PairWords<KIND>::FOUR) {
#pragma unroll
        for (int k = 0; k < 4; ++k) w.r[k] = s ? pw.b[k] : pw.a[k];
    } else {
        w.r[0] = s ? pw.a[2] : pw.a[0];
        w.r[1] = s ? pw.a[3] : pw.a[1];
        w.r[2] = w.r[3] = 0u;
    }
    if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        w.e[0] = s ? pw.c[2] : pw.c[0];
        w.e[1] = s ? pw.c[3] : pw.c[1];
    } else {
        w.e[0] = w.e[1] = 0u;
    }
    return w;
}

// Can the weight of this draw be non-zero?  For the running-coupling kinds: "the reference's weight is not NaN", a
// function of the random words alone.  Every quantity of the test is affine in the words (log sampling), so it is made
// on float(word) with the offsets and the factor 2^-32 log2(eps) < 0 folded into the thresholds (FastParams::wT*):
//   CRIT:      1 + Lambda(z_c theta_c) > 0                                       <=>  f1 < wT1
//   CRIT_SUB:  theta_c above both Landau thresholds (a constant)                 <=>  f1 < wT1
//              C < 1 for log(2 - 2C) in b_bar(C), C = csub / theta_c^beta        <=>  lc > wT2
//              1 + Lambda(C, alpha(theta_c)) > 0, which after multiplying by alpha's positive denominator is linear
//              in the logs: 1 + ca (max(ln P_T theta_c, 0) - ln M_Z + ln C) > 0; with max(A, 0) + B = max(A + B, B):
//              log2 C > c2  or  log2 theta_c + log2 C > c3                       <=>  lc < wT4  or  f1 + lc < wT3
//   with lc = f2 + 2 beta f3 - beta f1  (log2 C = sL2 lc + const).  Three conversions, two FMAs, one add, four compares.
template <int KIND, bool FIXED>
__device__ __forceinline__ bool active_from_words(const FastParams& P, const uint32_t* r) {
    if (!FIXED && KIND == JMC_KIND_CRIT) return (__uint2float_rn(r[1]) < P.wT1) & !P.crit_always_nan;
    if (!FIXED && KIND == JMC_KIND_CRIT_SUB) {
        const float f1 = __uint2float_rn(r[1]), f2 = __uint2float_rn(r[2]), f3 = __uint2float_rn(r[3]);
        const float lc = fmaf(-P.beta, f1, fmaf(2.0f * P.beta, f3, f2));
        const float x = lc + f1;
        return (f1 < P.wT1) & (lc > P.wT2) & ((x < P.wT3) | (lc < P.wT4));
    }
    return true;
}

// Stage 1 ("light"): the sample from its words, log2(observable), and whether the weight can be non-zero.
// FIXED selects the coupling at compile time so that the other branch costs nothing.
template <int KIND, bool FIXED, bool OBS_MLL>
__device__ __forceinline__ bool light(const FastParams& P, const Words& w, Draw& d, float& l2obs) {
    const uint32_t* r = w.r;
    if (KIND == JMC_KIND_RADIATOR) {
        // ungroomedSampler (samplers.py:204-226), C_ungroomed (observables.py:82-109)
        d.l2zs = l2_sample(r[0], P.sL2, -1.0f);
        d.l2ts = l2_sample(r[1], P.sL2, P.l2R);          // log2(theta_em), theta_scale folded into l2R
        d.z = ex2(d.l2zs);
        l2obs = fmaf(P.beta, d.l2ts, d.l2zs);
        if (OBS_MLL) l2obs += lg2(1.0f - d.z);
        return true;
    }
    // critical emission: z = z_c + delta, delta log-sampled (samplers.py:38-62)
    d.l2d = l2_sample(r[0], P.sL2, P.l2w_cz);
    d.l2tc = l2_sample(r[1], P.sL2, P.l2R);
    d.delta = ex2(d.l2d);
    d.z = P.zc + d.delta;
    // C_groomed (observables.py:15-79) = a b theta_c^beta / (1 - z_cut)^2 with a = zmin - f (z_cut - z_pre) =
    // delta + c0 and b = 1 - (1-f)(z_cut - z_pre) [- z at MLL] = bm0 - delta; the normalisation rides on `a`
    const float aK = fmaf(d.delta, P.aK, P.c0K);
    const float bb = OBS_MLL ? P.bm0 - d.delta : P.one_m;
    if (KIND == JMC_KIND_CRIT) {
        l2obs = fmaf(P.beta, d.l2tc, lg2(aK * bb));
        return active_from_words<KIND, FIXED>(P, r);
    }
    // subsequent emission from the ungroomed sampler
    d.l2zs = l2_sample(r[2], P.sL2, -1.0f);
    d.l2ts = l2_sample(r[3], P.sL2, P.l2R);
    if (KIND == JMC_KIND_CRIT_SUB) {
        // observable of twoEmissionWeight's integrand: max(C_groomed, csub), csub = z_s th_s^b th_s^b (weights.py:74)
        const float l2cc = fmaf(P.beta, d.l2tc, lg2(aK * bb));
        const float l2cs = fmaf(2.0f * P.beta, d.l2ts, d.l2zs);
        l2obs = fmaxf(l2cc, l2cs);
        d.lcl2 = fmaf(-P.beta, d.l2tc, l2cs);                        // log2(csub / theta_c^beta)
        if (FIXED) return d.lcl2 < -1.0f;                            // subPDF_fc mask: C < theta_c^beta / 2
        return active_from_words<KIND, FIXED>(P, r);
    }
    if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        // all-emission integrand with the zero bin
        // (examples/tests/comparison_tests/comparison_utils_probabilistic.py:384-474)
        d.l2zp = l2_sample(w.e[0], P.sL2, P.l2zc);
        d.u5 = __uint2float_rn(w.e[1]) * 2.3283064365386963e-10f;
        const float E2 = -P.g * d.l2tc;                              // E / ln2, E = 2 C_R a/pi ln(R0/theta_c)
        d.cdf = ex2(-E2 * P.ln_zc_eps);                              // exp(-E ln(z_cut/eps))
        const bool zero_bin = d.u5 < d.cdf;
        const float zp = zero_bin ? 0.0f : ex2(d.l2zp);
        // zmin - f (z_cut - z_pre) = delta + (1-f) z_cut + f z_pre, summed without passing through
        // z_cut - z_pre, which would swallow a z_pre below the FP32 spacing of z_cut
        const float a = d.delta + P.omf_zc + P.f * zp;
        const float omf_zcl = (1.0f - P.f) * (P.zc - zp);
        const float b = OBS_MLL ? (1.0f - d.z) - omf_zcl : 1.0f - omf_zcl;
        const float l2cc = lg2(a * b) + fmaf(P.beta, d.l2tc, P.l2_onorm);
        const float l2cs = fmaf(P.beta, d.l2tc, d.l2zs);            // c_sub = z_s theta_c^beta
        l2obs = fmaxf(l2cc, l2cs);
        return true;
    }
    return false;
}

// Stage 2 ("heavy"): weight * jacobian of a sample whose weight can be non-zero.
template <int KIND, bool FIXED>
__device__ __forceinline__ float heavy(const FastParams& P, const Draw& d) {
    const float inv_pi = (float)(1.0 / H_PI);
    if (KIND == JMC_KIND_RADIATOR) {
        // radiatorWeight (weights.py:15-26) * z theta: theta cancels, z multiplies the splitting function
        float zp;
        if (P.split_acc == JMC_ACC_LL) zp = P.twoCR;
        else if (P.split_acc == JMC_ACC_LL_NONSING) zp = z_times_p(d.z, P.jet);
        else zp = z_times_p(d.z, P.jet) + d.z * p_of(1.0f - d.z, P.jet);
        const float alpha = FIXED ? (float)H_ALPHA : alpha_from_lnmu(P, fmaf(JMC_LN2F, d.l2zs + d.l2ts, P.lnPT));
        return zp * alpha * inv_pi;
    }
    const float l2z = lg2(d.z);
    const float Lc = JMC_LN2F * d.l2tc;
    if (KIND == JMC_KIND_CRIT) {
        // criticalEmissionWeight (weights.py:31-41) * (z - z_c) theta
        if (FIXED) return P.g * ex2(d.l2d - l2z + P.A * d.l2tc);
        float R;
        if (!crit_rad_rc(P, Lc, R)) return fnan();
        const float alpha = alpha_from_lnmu(P, P.lnPT + JMC_LN2F * l2z + Lc);
        return p_bar(d.z, P.jet) * d.delta * alpha * inv_pi * ex2(-R * JMC_INV_LN2F);
    }
    if (KIND == JMC_KIND_CRIT_SUB) {
        // twoEmissionWeight (weights.py:63-99) * jacobians
        const float lc = JMC_LN2F * d.lcl2;
        if (FIXED) {
            // critPDF_fc * subPDF_fc(csub, maxRadius=theta_c) * th_s^b th_c^b * (z - z_c) th_c z_s th_s
            const float x = JMC_LN2F + lc;                           // ln(2 csub / theta_c^beta) < 0
            const float e2 = d.l2d - l2z + (P.A + P.beta) * d.l2tc + (1.0f - P.beta) * d.l2ts
                             - P.c1 * x * x * JMC_INV_LN2F;
            return P.g * 2.0f * P.c1 * (-x) * ex2(e2);
        }
        float Rc, Rs, rp;
        if (!sub_rc(P, lc, Lc, Rs, rp) || !crit_rad_rc(P, Lc, Rc)) return fnan();
        const float alpha = alpha_from_lnmu(P, P.lnPT + JMC_LN2F * l2z + Lc);
        const float lin = p_bar(d.z, P.jet) * d.delta * alpha * inv_pi * rp;
        return lin * ex2(P.beta * d.l2tc + (1.0f - P.beta) * d.l2ts - (Rc + Rs) * JMC_INV_LN2F);
    }
    if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        // critPDF_fc * subPDF_fc(z_s) * precriticalEmissionWeight_zerobin * jacobians
        // (comparison_utils_probabilistic.py:82-98, 432-464)
        const float E = -P.g * Lc;
        const bool zero_bin = d.u5 < d.cdf;
        const float x = JMC_LN2F * (1.0f + d.l2zs);                  // ln(2 z_s) < 0
        const float pre = zero_bin ? d.cdf : E * ex2(-E * (P.l2zc - d.l2zp));
        const float e2 = d.l2d - l2z + P.A * d.l2tc + d.l2ts - P.c1 * x * x * JMC_INV_LN2F;
        return P.g * 2.0f * P.c1 * (-x) * pre * ex2(e2);
    }
    return 0.0f;
}

// ---- binning ---------------------------------------------------------------
// s_edges: edges in the compare domain (log2 or linear), FP32.  Returns -1 for a dropped sample.  (The generic
// form; np.logspace bins take the affine shortcut in fast_body.)
__device__ __forceinline__ int bin_fast(const FastParams& P, float l2obs, const float* __restrict__ s_edges) {
    const float v = P.dom_log ? l2obs : ex2(l2obs);
    const int ne = P.n_edges;
    const bool in_range = (v >= P.e_first) & (v <= P.e_last);       // NaN fails both
    int idx;
    if (P.uniform) {
        // affine guess, then settle against the stored edges (FP32 rounding of the map can be off by one)
        idx = __float2int_rd(fmaf(v, P.inv_d, P.mb0_inv_d));
        idx = max(0, min(idx, ne - 2));
        const float e_lo = s_edges[idx], e_hi = s_edges[idx + 1];
        idx += (int)((v >= e_hi) & (idx < ne - 2)) - (int)(v < e_lo);
    } else {
        int lo = 0, hi = ne;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_edges[mid] <= v) lo = mid + 1; else hi = mid;
        }
        idx = min(lo - 1, ne - 2);
    }
    return in_range ? idx : -1;
}

// HOT: histogram only (no min/max tracking) into np.logspace-like bins with nan_to_num weights -- the form every
// large job has; the generic instantiation reads those choices from FastParams at run time.
// COUNT_ALL: the per-bin count is np.histogram's (every sample inside the binned range).  With COUNT_ALL = false the
// count is taken over the samples that carry a non-zero weight only.  The reference uses the count for one thing: to
// zero the density of empty bins (integrator.py:101,123-124), and a bin without a weighted sample has sum w = 0
// anyway, so density, error and integral are the same -- but the kinds whose weight is NaN -> 0 for most samples (the
// running-coupling Landau region: 99.7 % of CRIT_SUB samples) then need neither observable nor bin nor count atomic
// for those samples: the random words decide (active_from_words) and the rest is evaluated only for the survivors.
// SLIM: the 1024-thread instantiations (block histograms above 72 KB) have 64 registers per thread
template <int KIND, bool FIXED, bool OBS_MLL, bool HOT, bool COUNT_ALL, bool SLIM = false>
__device__ __forceinline__ void
fast_body(const FastParams& P_in, uint64_t n, uint64_t seed, uint64_t first, const float* __restrict__ g_edges,
          double* __restrict__ g_sum_w, double* __restrict__ g_sum_w2, unsigned long long* __restrict__ g_count,
          int* g_poison, double* __restrict__ d_minmax, double* s_raw) {
    __shared__ int s_poison;
    // (The per-sample parameters are re-read from the constant bank once per trip.  Pinning them in registers
    // through opaque inline asm does not survive ptxas, which rematerialises parameter loads; measured: no effect.)
    const FastParams& P = P_in;
    const int ne = P.n_edges, nb = ne > 0 ? ne - 1 : 0;
    const int sh = P.spread_shift;
    const int n_slots = nb << sh;
    acc_w_t* s_sum_w = (acc_w_t*)s_raw;                           // slot [bin][lane & (R-1)]
    acc_w2_t* s_sum_w2 = (acc_w2_t*)(s_sum_w + n_slots);
    unsigned int* s_count = (unsigned int*)(s_sum_w2 + n_slots);  // the same slots + 32 dummy slots for dropped samples
    float* s_edges = (float*)(s_count + n_slots + 32);
    const bool do_hist = HOT || g_edges != nullptr;
    const bool do_minmax = !HOT && d_minmax != nullptr;
    if (threadIdx.x == 0) s_poison = INT_MAX;
    for (int k = threadIdx.x; k < n_slots; k += blockDim.x) { s_sum_w[k] = 0; s_sum_w2[k] = 0; s_count[k] = 0u; }
    if (threadIdx.x < 32) s_count[n_slots + threadIdx.x] = 0u;
    for (int k = threadIdx.x; k < ne; k += blockDim.x) s_edges[k] = g_edges[k];
    __syncthreads();
    float mn = INFINITY, mx = -INFINITY;
    // Kinds whose weight can be NaN by construction (running-coupling Landau region) run the hot instantiation
    // only with nan_to_num set; for the others the flag is looked at only if a NaN ever shows up.
    constexpr bool NAN_REGION = !FIXED && KIND != JMC_KIND_RADIATOR;
    const bool nan_to_num = (HOT && NAN_REGION) || P.nan_to_num;
    // Warp-level compaction of the rare samples whose weight is not NaN -> 0.  In the running-coupling
    // critical x subsequent integrand 0.3 % of the samples are active, i.e. 9 % of the warps would run the
    // ~330-instruction radiator code with one or two live lanes.  Instead the words of the active samples are
    // appended to a per-warp shared-memory queue (ballot + popc, no atomics) and evaluated 32 at a time with every
    // lane busy.  (Also the running-coupling critical integrand: 35 % of its samples are above the Landau scale.)
    constexpr bool COMPACT = (KIND == JMC_KIND_CRIT_SUB || KIND == JMC_KIND_CRIT) && !FIXED;
    constexpr bool LAZY = COMPACT && HOT && !COUNT_ALL;     // observable and bin only for the samples in the queue
    constexpr int QW = KIND == JMC_KIND_CRIT ? 2 : 4;       // words per queued sample
    const unsigned lane = threadIdx.x & 31u;

    // weight * jacobian of one sample into the block histogram (bq: bin, or -1 / -2 for a sample below /
    // above the binned range, which matters only for NumPy's NaN poisoning).
    // Back ends (JMC_ACC_MODE).  Shared-memory floating-point adds are compare-and-swap loops on sm_100 (LDS, add,
    // ATOMS.CAST.SPIN, branch; only 32-bit integer adds are native), and a loop runs one round per lane of the warp
    // on the same address -- with 99 log bins that is 3-6 rounds for most warps.  Measured on B200 (C1 / C3 fc / C4,
    // 1e9 samples, profiles/r02b_shootout.log):
    //  0: two FP64 loops on [bin]                                              2.64e11 / 1.78e11 / 1.43e11
    //  3 (default): R = 2^sh lane-interleaved copies of every bin, slot [bin][lane & (R-1)]: lanes of a warp that hit
    //     the same bin no longer share an address, so a loop almost never repeats; the copies are summed when the
    //     block flushes                                                         2.97e11 / 1.87e11 / 1.54e11
    //  2: __match_any_sync on the bin, peers reduced by shuffles, one loop pair per distinct bin per warp: the
    //     match + shuffle rounds cost more than the retries they save          2.15e11 / 1.38e11 / 1.20e11
    //  (dropped: both sums as two floats in one slot updated by an explicit 64-bit atomicCAS -- ATOMS.CAS.64 returns
    //   the old value and is served lane by lane, ~64 cycles per warp: 0.98e11 / 1.49e11 / 1.38e11)
    //  4 / 5: mode 3 with FP32 slots for both sums / for sum w^2 only (summed in FP64 at the flush)
#if JMC_ACC_MODE == 9
    float dbg_w = 0.0f, dbg_w2 = 0.0f;
#endif
    auto accumulate = [&](float w, int bq) {
        if (bq >= 0 && fabsf(w) > 0.0f) {                // the common case first: one compare rejects 0 and NaN
            if (LAZY) atomicAdd(&s_count[(bq << sh) | (int)(lane & ((1u << sh) - 1u))], 1u);
#if JMC_ACC_MODE == 9       // measurement only: no shared-memory atomics at all (upper bound of any back end)
            dbg_w += w + (float)bq;
            dbg_w2 += w * w;
#elif JMC_ACC_MODE == 2
            const unsigned mask = __activemask();
            const unsigned peers = __match_any_sync(mask, bq);
            const int leader = __ffs(peers) - 1;
            unsigned others = peers & ~(1u << leader);
            float aw = w, aw2 = w * w;
            const float mw = aw, mw2 = aw2;
            while (__any_sync(mask, others != 0u)) {
                const int src = others ? __ffs(others) - 1 : (int)lane;
                const float vw = __shfl_sync(mask, mw, src), vw2 = __shfl_sync(mask, mw2, src);
                if (others) { aw += vw; aw2 += vw2; }
                others &= others - 1u;
            }
            if ((int)lane == leader) {
                atomicAdd(&s_sum_w[bq], (double)aw);
                atomicAdd(&s_sum_w2[bq], (double)aw2);
            }
#else
            const int slot = (bq << sh) | (int)(lane & ((1u << sh) - 1u));
            const acc_w_t wd = (acc_w_t)w;
            atomicAdd(&s_sum_w[slot], wd);
            atomicAdd(&s_sum_w2[slot], (acc_w2_t)(wd * wd));
#endif
        } else if (w != w && !nan_to_num) {              // NaN: 0 under nan_to_num, else NumPy poisons bins >= bq
            const int pb = bq >= 0 ? bq : (bq == -1 ? 0 : INT_MAX);
            if (pb != INT_MAX) atomicMin(&s_poison, pb);
        }
    };

    // Bin of an observable: `ok` = inside the binned range, `b` = its bin (meaningful only when ok).
    // HOT (np.logspace bins): the bin is the floor of an affine map of log2(obs), and the range test is made on the
    // mapped value -- FFMA, one compare, F2I.  The host anchors the map at the first edge and lowers its slope by
    // single ulps (make_fast_params) until the first and the last edge themselves (the extreme samples, when the edges
    // come from setBins) map inside [0, nb) whatever the rounding of the FMA.
    auto locate = [&](float l2obs, bool& ok, int& b) {
        if (HOT) {
            const float f = fmaf(l2obs, P.inv_d, P.mb0_inv_d);
            // 0 <= f < nb in ONE unsigned compare of the bit patterns: non-negative floats order like their bits,
            // and a set sign bit or a NaN pattern is a huge unsigned number
            ok = (unsigned)__float_as_int(f) < (unsigned)__float_as_int(P.nb_f);
            b = __float2int_rd(f);
        } else {
            b = bin_fast(P, l2obs, s_edges);
            ok = b >= 0;
        }
    };
    // bin for a weight, or -1 / -2 for a sample below / above the binned range (NaN poisoning only)
    auto weight_bin = [&](float l2obs) {
        bool ok;
        int b;
        locate(l2obs, ok, b);
        if (ok) return b;
        if (HOT) return -1;
        const float v = P.dom_log ? l2obs : ex2(l2obs);
        return v < P.e_first ? -1 : -2;
    };
    const int dummy_slot = n_slots + (int)lane;
    const int lane_copy = (int)(lane & ((1u << sh) - 1u));
    // Every sample: min/max and the count histogram (native 32-bit shared atomics on the lane-interleaved copy of the
    // bin: with 16-32 copies the lanes of a warp fall on distinct banks).  Dropped samples count into a per-lane
    // dummy slot behind the bins: no branch around the atomic, no same-address serialisation among the lanes that miss.
    // `valid` is a compile-time true in the main loop and a run-time flag only in the ragged last trips.
    auto tally = [&](float l2obs, bool valid) {
        if (do_minmax && valid) {
            mn = fminf(mn, l2obs);
            mx = fmaxf(mx, l2obs);
        }
        if (!do_hist) return;
        bool ok;
        int b;
        locate(l2obs, ok, b);
        atomicAdd(&s_count[(ok && valid) ? ((b << sh) | lane_copy) : dummy_slot], 1u);
    };
    // a sample whose weight can be non-zero: everything from its words
    auto weigh_words = [&](const Words& w) {
        Draw d;
        float o;
        light<KIND, FIXED, OBS_MLL>(P, w, d, o);
        accumulate(heavy<KIND, FIXED>(P, d), weight_bin(o));
    };

    // per-warp queue of active samples: QW words per slot, 64 slots (at most 31 left over + 32 appended)
    uint32_t* q = (uint32_t*)(s_edges + ((ne + 3) & ~3)) + (threadIdx.x >> 5) * (kQueueFields * kQueueSlots);
    int qn = 0;                                          // warp-uniform queue length
    auto drain = [&](int slot) {
        Words w;
#pragma unroll
        for (int k = 0; k < 4; ++k) w.r[k] = k < QW ? q[k * kQueueSlots + slot] : 0u;
        w.e[0] = w.e[1] = 0u;
        weigh_words(w);
    };
    // warp-collective: append the lanes whose sample is active, evaluate 32 queued samples when there are that many
    auto enqueue = [&](const Words& w, bool active) {
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (!m) return;
        if (active) {
            const int slot = qn + __popc(m & ((1u << lane) - 1u));
#pragma unroll
            for (int k = 0; k < QW; ++k) q[k * kQueueSlots + slot] = w.r[k];
        }
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            qn -= 32;
            drain(qn + (int)lane);
            __syncwarp();
        }
    };
    // One sample with every check (ragged trips, the half pairs at the ends of the index range, and the generic
    // per-sample order of the kinds that are not compacted).  Warp-collective when COMPACT.
    auto one = [&](const Words& w, bool valid) {
        if (LAZY) {
            enqueue(w, valid && active_from_words<KIND, FIXED>(P, w.r));
            return;
        }
        Draw d;
        float o;
        const bool active = light<KIND, FIXED, OBS_MLL>(P, w, d, o) && valid;
        tally(o, valid);
        if (!do_hist) return;
        if (!FIXED && !nan_to_num && !active && valid) accumulate(fnan(), weight_bin(o));   // inactive = NaN weight
        if (COMPACT) enqueue(w, active);
        else if (active) accumulate(heavy<KIND, FIXED>(P, d), weight_bin(o));
    };

    // ---- the index range in pairs.  Interior pairs [pa, pb) have both samples inside [first, first + n); an odd
    // `first` / an odd end leave one half pair each, handled by warp 0 of block 0 after the loop.  The host never
    // lets a launch cross a multiple of 2^32 in the sample index, so the high word of the pair counter is
    // launch-uniform and the low word is plain 32-bit arithmetic (no wrap inside the launch).
    const uint64_t end = first + n;
    const uint64_t pa = (first + 1) >> 1, pb = end >> 1;
    const uint32_t np = pb > pa ? (uint32_t)(pb - pa) : 0u;
    const uint32_t p_hi = (uint32_t)(first >> 33);

    // V pairs per trip: their Philox chains are independent, so the scheduler interleaves them (the generator
    // is a 10-deep dependent chain of IMAD.WIDE -> LOP3), and the warp-uniform work of a trip (constant loads, loop
    // control, and for the compacted kinds ONE vote on "any lane has an active sample") is paid once per 2V samples.
    PhiloxKeys keys(seed);
#if JMC_FAST_PIN_KEYS
    keys.pin();
#endif
    // (measured with chunked work, profiles/r02q_pairs.log: the compacted 4-word kinds gain 3-6 % from two pairs per
    // trip -- one vote per four samples --, the others lose 1 %)
    constexpr int V = PairWords<KIND>::FOUR ? ((COMPACT && !SLIM) ? JMC_FAST_PAIRS_COMPACT : JMC_FAST_PAIRS) : JMC_FAST_PAIRS_2W;
    // Work is handed out, not pre-assigned.  A UNIT is 32 V consecutive pairs (one trip of one warp: lane l takes pairs
    // l, l + 32, ...); warps take JMC_FAST_CHUNK units at a time from one global counter.  With a fixed 1/#warps share
    // per warp the hardware scheduler lets some warps of an SM run ahead, they exit early, and the SM finishes its
    // work with ever fewer warps to hide latencies (ncu: 23 of the 32 resident warps active on average over the
    // kernel).  The next chunk is requested before the current one is processed, so the atomic's latency is hidden.
    // g_poison[1] is the counter, g_poison[2] counts finished blocks: the last one re-arms both for the next launch.
    const uint32_t n_units = np / (32u * (uint32_t)V);
    // Large jobs only (>= 16 units per warp): a small job is launch- and latency-bound, there every warp takes its
    // fixed share (unit w, w + #warps, ...) without touching the counter.
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const bool dynamic = n_units >= 16u * n_warps;
    const uint32_t chunk_units = dynamic ? min((uint32_t)JMC_FAST_CHUNK, n_units / (4u * n_warps)) : 1u;
    const uint32_t n_chunks = (n_units + chunk_units - 1u) / chunk_units;
    unsigned int* g_next = (unsigned int*)g_poison + 1;
    unsigned int fetched = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);        // static: the warp's number
    if (dynamic && lane == 0) fetched = atomicAdd(g_next, 1u);
    while (true) {
        const uint32_t chunk = __shfl_sync(0xffffffffu, fetched, 0);
        if (chunk >= n_chunks) break;
        if (!dynamic) fetched += n_warps;
        else if (lane == 0) fetched = atomicAdd(g_next, 1u);
        const uint32_t u_end = min(n_units, (chunk + 1u) * chunk_units);
        uint32_t idx = (uint32_t)pa + chunk * chunk_units * (32u * (uint32_t)V) + lane;
        for (uint32_t u = chunk * chunk_units; u < u_end; ++u, idx += 32u * (uint32_t)V) {
            PairWords<KIND> pw[V];
#pragma unroll
            for (int v = 0; v < V; ++v) draw_pair<KIND>(idx + (uint32_t)v * 32u, p_hi, keys, pw[v]);
            if (COMPACT) {
                bool act[2 * V];
                bool any = false;
#pragma unroll
                for (int k = 0; k < 2 * V; ++k) {
                    const Words w = sample_words<KIND>(pw[k >> 1], k & 1);
                    if (LAZY) {
                        act[k] = active_from_words<KIND, FIXED>(P, w.r);
                    } else {
                        Draw d;
                        float o;
                        act[k] = light<KIND, FIXED, OBS_MLL>(P, w, d, o);
                        tally(o, true);
                        if (do_hist && !nan_to_num && !act[k]) accumulate(fnan(), weight_bin(o));
                    }
                    any |= act[k];
                }
                if (do_hist && __any_sync(0xffffffffu, any)) {
#pragma unroll
                    for (int k = 0; k < 2 * V; ++k) enqueue(sample_words<KIND>(pw[k >> 1], k & 1), act[k]);
                }
            } else {
                // all arithmetic of the trip first (independent chains the scheduler can interleave), then its
                // shared-memory atomics back to back: an atomic in the middle would order the samples' instruction
                // streams
                float o[2 * V], wt[2 * V];
#pragma unroll
                for (int k = 0; k < 2 * V; ++k) {
                    Draw d;
                    const bool active = light<KIND, FIXED, OBS_MLL>(P, sample_words<KIND>(pw[k >> 1], k & 1), d, o[k]);
                    wt[k] = 0.0f;
                    if (do_hist) {
                        if (active) wt[k] = heavy<KIND, FIXED>(P, d);
                        else if (!FIXED && !nan_to_num) wt[k] = fnan();
                    }
                }
#pragma unroll
                for (int k = 0; k < 2 * V; ++k) {
                    tally(o[k], true);
                    if (do_hist) accumulate(wt[k], weight_bin(o[k]));
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // the pairs behind the last whole unit (fewer than 32 V), with every check
        const uint32_t done = n_units * 32u * (uint32_t)V;
#pragma unroll 1
        for (int v = 0; v < V; ++v) {
            const uint32_t k = done + (uint32_t)v * 32u + lane;
            if (!__any_sync(0xffffffffu, k < np)) break;
            PairWords<KIND> pw;
            draw_pair<KIND>((uint32_t)pa + k, p_hi, keys, pw);
            one(sample_words<KIND>(pw, 0), k < np);
            one(sample_words<KIND>(pw, 1), k < np);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        // half pairs: lane 0 takes sample `first` when it is odd, lane 1 takes sample end-1 when `end` is odd
        const bool front = lane == 0 && (first & 1ull), back = lane == 1 && (end & 1ull);
        if (__any_sync(0xffffffffu, front || back)) {
            PairWords<KIND> pw;
            draw_pair<KIND>((uint32_t)((back ? end : first) >> 1), p_hi, keys, pw);
            one(sample_words<KIND>(pw, back ? 0 : 1), front || back);
        }
    }
    if (COMPACT && do_hist) {
        __syncwarp();
        if ((int)lane < qn) drain((int)lane);
    }
#if JMC_ACC_MODE == 9
    if (dbg_w == 1.2345f) atomicAdd(&s_sum_w[0], (acc_w_t)dbg_w2);
#endif
    __syncthreads();
    if (do_hist) {
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            unsigned int c = 0u;
            for (int r = 0; r < (1 << sh); ++r) c += s_count[(k << sh) + r];
            if (c) {
                atomicAdd(&g_count[k], (unsigned long long)c);
                double sw = 0.0, sw2 = 0.0;
                for (int r = 0; r < (1 << sh); ++r) {
                    sw += (double)s_sum_w[(k << sh) + r];
                    sw2 += (double)s_sum_w2[(k << sh) + r];
                }
                if (sw != 0.0) atomicAdd(&g_sum_w[k], sw);
                if (sw2 != 0.0) atomicAdd(&g_sum_w2[k], sw2);
            }
        }
        if (threadIdx.x == 0 && s_poison != INT_MAX) atomicMin(g_poison, s_poison);
    }
    // the last block to get here has seen every other block leave the loop: re-arm the work counter
    if (dynamic && threadIdx.x == 0) {
        unsigned int* g_done = (unsigned int*)g_poison + 2;
        if (atomicAdd(g_done, 1u) == gridDim.x - 1u) {
            *g_next = 0u;
            *g_done = 0u;
        }
    }
    if (do_minmax) {
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

// one job: parameters in the constant bank
// BIG: 1024-thread blocks, one per SM, for histograms that need most of an SM's shared memory (the reference's
// default of 5000 radiator bins is 130 KB per block): 32 warps then share one histogram instead of 8.
template <int KIND, bool FIXED, bool OBS_MLL, bool HOT, bool BIG, bool COUNT_ALL>
__global__ void __launch_bounds__(BIG ? kFastThreadsBig : kFastThreads, BIG ? 1 : JMC_FAST_MIN_BLOCKS)
run_fast_kernel(FastParams P, uint64_t n, uint64_t seed, uint64_t first, const float* __restrict__ g_edges,
                double* __restrict__ g_sum_w, double* __restrict__ g_sum_w2,
                unsigned long long* __restrict__ g_count, int* g_poison, double* __restrict__ d_minmax) {
    extern __shared__ __align__(16) double s_dyn[];
    fast_body<KIND, FIXED, OBS_MLL, HOT, COUNT_ALL, BIG>(P, n, seed, first, g_edges, g_sum_w, g_sum_w2, g_count, g_poison,
                                                    d_minmax, s_dyn);
}

// A grid of parameter points in one launch (BASELINE config 5: gen_crit_sub_num_rad's theta_crit loop,
// generation.py:303-359, z_cut / beta sweeps): blockIdx.y selects the point, its parameters are staged in shared
// memory, and every point regenerates the same sample stream -- the reference re-weights one stored sample set
// per grid point, here the "stored set" is the counter RNG.
template <int KIND, bool FIXED, bool OBS_MLL, bool HOT, bool BIG>
__global__ void __launch_bounds__(BIG ? kFastThreadsBig : kFastThreads, BIG ? 1 : JMC_FAST_MIN_BLOCKS)
run_fast_batch_kernel(const FastParams* __restrict__ params, uint64_t n, uint64_t seed, uint64_t first,
                      const float* __restrict__ edges_all, int n_edges, double* __restrict__ sum_w_all,
                      double* __restrict__ sum_w2_all, unsigned long long* __restrict__ count_all, int* poison_all,
                      double* __restrict__ minmax_all) {
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ FastParams sP;
    const int point = blockIdx.y;
    const int words = (int)(sizeof(FastParams) / sizeof(int));
    for (int k = threadIdx.x; k < words; k += blockDim.x) ((int*)&sP)[k] = ((const int*)&params[point])[k];
    __syncthreads();
    const size_t nb = n_edges > 0 ? (size_t)n_edges - 1 : 0;
    fast_body<KIND, FIXED, OBS_MLL, HOT, true, BIG>(
        sP, n, seed, first, edges_all ? edges_all + (size_t)point * n_edges : nullptr,
        sum_w_all ? sum_w_all + point * nb : nullptr, sum_w2_all ? sum_w2_all + point * nb : nullptr,
        count_all ? count_all + point * nb : nullptr, poison_all + 4 * point,
        minmax_all ? minmax_all + 2 * (size_t)point : nullptr, s_dyn);
}

// one block; leaves the flag re-armed ("no poison") for the next job that uses the same scratch
__global__ void poison_fast_kernel(int* g_poison, int nb, double* g_sum_w, double* g_sum_w2) {
    const int pb = *g_poison;
    __syncthreads();
    if (threadIdx.x == 0) *g_poison = 0x7f7f7f7f;
    if (pb < 0 || pb >= nb) return;
    for (int k = pb + threadIdx.x; k < nb; k += blockDim.x) {
        g_sum_w[k] = __longlong_as_double(0x7ff8000000000000LL);
        g_sum_w2[k] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

// samples on request: per-sample samples / observable / weight of the FP32 path
template <int KIND, bool FIXED>
__global__ void __launch_bounds__(kFastThreads)
dump_fast_kernel(FastParams P, uint64_t n, uint64_t seed, uint64_t first, double* __restrict__ samples,
                 double* __restrict__ obs, double* __restrict__ weight) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Draw d;
        d.l2d = d.l2tc = d.l2zs = d.l2ts = d.l2zp = -INFINITY;
        d.delta = d.z = d.cdf = 0.0f;
        d.u5 = 2.0f;
        float l2obs;
        const PhiloxKeys keys(seed);
        const uint64_t gi = first + i, pair = gi >> 1;
        PairWords<KIND> pw;
        draw_pair<KIND>((uint32_t)pair, (uint32_t)(pair >> 32), keys, pw);
        const Words wd = sample_words<KIND>(pw, (int)(gi & 1ull));
        const bool active = P.obs_mll ? light<KIND, FIXED, true>(P, wd, d, l2obs)
                                      : light<KIND, FIXED, false>(P, wd, d, l2obs);
        float w = active ? heavy<KIND, FIXED>(P, d) : (FIXED ? 0.0f : fnan());
        if (w != w && P.nan_to_num) w = 0.0f;
        if (samples) {
            double* o = samples + 8 * i;
            const bool has_crit = KIND != JMC_KIND_RADIATOR;
            const bool zero_bin = KIND == JMC_KIND_PRE_CRIT_SUB && d.u5 < d.cdf;
            o[0] = has_crit ? d.z : 0.0;
            o[1] = has_crit ? ex2(d.l2tc) : 0.0;
            o[2] = KIND == JMC_KIND_RADIATOR ? d.z : ex2(d.l2zs);
            o[3] = ex2(d.l2ts);
            o[4] = KIND == JMC_KIND_PRE_CRIT_SUB ? (zero_bin ? 0.0 : ex2(d.l2zp)) : 0.0;
            o[5] = d.u5;
            o[6] = has_crit ? d.delta : 0.0;
            o[7] = 0.0;
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
    P.split_acc = g->split_acc; P.nan_to_num = g->nan_to_num; P.count_weighted = g->counts == JMC_COUNT_WEIGHTED;
    P.L2 = (float)std::log2(g->epsilon);
    P.sL2 = (float)(std::log2(g->epsilon) * 2.3283064365386963e-10);
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
    {
        const double K = 1.0 / ((1.0 - g->z_cut) * (1.0 - g->z_cut));
        P.aK = (float)K;
        P.c0K = (float)((g->z_cut - g->f_soft * zcl) * K);
        P.bm0 = (float)(1.0 - g->z_cut - (1.0 - g->f_soft) * zcl);
        P.omb = (float)(1.0 - g->beta);
        const double ca = 2.0 * H_AMZ * H_BETA0;
        const double c2 = (std::log(H_MZ) - 1.0 / ca) / ln2;
        P.act_c2 = (float)c2;
        P.act_c3 = (float)(c2 - std::log2(H_PT));
    }
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
    {
        // CRIT_SUB, running coupling: the weight is NaN unless 1 + Lambda(MU_NP, alpha(theta_c)) > 0, i.e.
        // 1 + ca (ln(P_T theta_c) - ln M_Z + ln MU_NP) > 0 (theta_c above the freezing scale there), and unless
        // the critical radiator's 1 + Lambda(z_c theta_c) > 0.
        const double ca = 2.0 * H_AMZ * H_BETA0;
        const double ln_th_sub = -1.0 / ca + std::log(H_MZ) - std::log(H_MU) - std::log(H_PT);
        const double k = 2.0 * H_ALPHA * H_BETA0;
        const double ln_th_crit = g->z_cut > 0 ? -1.0 / k - std::log(g->z_cut) : -INFINITY;
        P.l2tc_min = (float)(std::fmax(ln_th_sub, ln_th_crit) / ln2);
    }
    {
        // the activity tests on float(word): l2x = sL2 * float(word) + offset with sL2 < 0, so "l2x > t" is
        // "float(word) < (t - offset) / sL2" (active_from_words)
        const double sL2 = std::log2(g->epsilon) * 2.3283064365386963e-10, l2R = std::log2(g->radius);
        const double k0 = -1.0 + g->beta * l2R;             // log2 C = sL2 (f2 + 2 beta f3 - beta f1) + k0
        if (g->kind == JMC_KIND_CRIT) {
            const double k = 2.0 * H_ALPHA * H_BETA0;
            P.wT1 = (float)(((-1.0 / k - std::log(g->z_cut)) / ln2 - l2R) / sL2);
        } else {
            P.wT1 = (float)(((double)P.l2tc_min - l2R) / sL2);
        }
        P.wT2 = (float)((0.0 - k0) / sL2);
        P.wT4 = (float)(((double)P.act_c2 - k0) / sL2);
        P.wT3 = (float)(((double)P.act_c3 - l2R - k0) / sL2);
    }
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
        for (int k = 0; k < n_edges; ++k) (*edges_f)[k] = (float)dom[k];
        P.e_first = (*edges_f)[0];
        P.e_last = (*edges_f)[n_edges - 1];
        P.nb_f = (float)(n_edges - 1);
        // The affine bin map f(x) = fmaf(x, inv_d, mb0_inv_d), anchored at the first edge: mb0_inv_d is the smallest
        // float that maps the FP32 first edge to f >= 0, and the slope inv_d is lowered ulp by ulp (each costs < 1e-5
        // of a bin at the far end) until the FP32 last edge maps below nb.  An observable equal to an end edge (the
        // extreme sample, when the edges come from setBins) is then binned, as np.histogram's closed last bin does,
        // whatever the rounding of the FMA.
        P.inv_d = (float)(1.0 / d);
        for (int it = 0; it < 4096; ++it) {
            P.mb0_inv_d = (float)(-(double)P.e_first * (double)P.inv_d);
            while (std::fmaf(P.e_first, P.inv_d, P.mb0_inv_d) < 0.0f) P.mb0_inv_d = std::nextafterf(P.mb0_inv_d, INFINITY);
            if (std::fmaf(P.e_last, P.inv_d, P.mb0_inv_d) < P.nb_f) break;
            P.inv_d = std::nextafterf(P.inv_d, 0.0f);
        }
        for (int k = 1; k < n_edges; ++k)
            if (!((*edges_f)[k] > (*edges_f)[k - 1]))
                return set_error(JMC_ELIMIT, "jmc_run: bin edges %d and %d coincide in FP32; use JMC_F64 for bins this fine",
                                 k - 1, k);
    }
    *out = P;
    return JMC_OK;
}

static size_t fast_smem_bytes(int n_edges, int threads = kFastThreads, int shift = 0) {
    const size_t nb = n_edges > 0 ? (size_t)n_edges - 1 : 0;
    const size_t ne_pad = ((size_t)(n_edges > 0 ? n_edges : 0) + 3) & ~(size_t)3;
    const size_t acc = (sizeof(acc_w_t) + sizeof(acc_w2_t)) * (nb << shift);
    return acc + sizeof(unsigned int) * ((nb << shift) + 32) + sizeof(float) * ne_pad
           + sizeof(float) * (threads / 32) * kQueueFields * kQueueSlots;
}
// lane-interleaved accumulator copies: as many (up to one per lane) as keep the block under `budget` bytes
static int fast_spread_shift(int n_edges, int threads, size_t budget) {
    if (JMC_ACC_MODE == 0 || JMC_ACC_MODE == 2) return 0;
    static const int max_shift = [] {                    // experiment switch (scripts/acc_shootout.sh)
        const char* e = getenv("JMC_FAST_SPREAD_SHIFT");
        return e ? atoi(e) : 5;
    }();
    int shift = max_shift;
    while (shift > 0 && fast_smem_bytes(n_edges, threads, shift) > budget) --shift;
    return shift;
}
static constexpr size_t kSpreadBudget = 56 * 1024, kSpreadBudgetBig = 200 * 1024;

// samples from global index `first` up to the next multiple of 2^32
static uint64_t fast_span(uint64_t first) { return 0x100000000ull - (first & 0xffffffffull); }

template <int KIND, bool FIXED, bool OBS_MLL, bool HOT, bool BIG, bool COUNT_ALL = true>
static int launch_fast_t(const FastParams& P, uint64_t n, uint64_t seed, uint64_t first, const float* d_edges_f,
                       double* d_sum_w, double* d_sum_w2, uint64_t* d_count, int* d_poison, double* d_minmax,
                       cudaStream_t st) {
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    constexpr int threads = BIG ? kFastThreadsBig : kFastThreads;
    const int shift = fast_spread_shift(d_edges_f ? P.n_edges : 0, threads, BIG ? kSpreadBudgetBig : kSpreadBudget);
    const size_t smem = fast_smem_bytes(d_edges_f ? P.n_edges : 0, threads, shift);
    if (smem > (size_t)di.max_smem_optin)
        return set_error(JMC_ELIMIT, "jmc_run: %d edges need %zu B of shared memory (limit %d B)", P.n_edges, smem,
                         di.max_smem_optin);
    auto kern = run_fast_kernel<KIND, FIXED, OBS_MLL, HOT, BIG, COUNT_ALL>;
    // attribute + occupancy query once per (kernel, device, shared-memory size): small jobs are launch-bound
    static thread_local struct { int device; size_t smem; int per_sm; } cfg = {-1, 0, 0};
    if (cfg.device != di.device || cfg.smem != smem) {
        JMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        JMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        cfg = {di.device, smem, per_sm < 1 ? 1 : per_sm};
    }
    // persistent: at most one wave of resident blocks; small jobs get fewer blocks (>= 8 sample pairs per thread),
    // so that zeroing and flushing the block histograms does not dominate them
    const uint64_t want = (n / 16 + threads) / threads;
    static const uint64_t waves = [] {                   // experiment switch: blocks per resident slot
        const char* e = getenv("JMC_FAST_WAVES");
        return e ? (uint64_t)std::max(1, atoi(e)) : (uint64_t)1;
    }();
    const uint64_t cap = (uint64_t)di.sm_count * cfg.per_sm * waves;
    const int grid = (int)(want < cap ? want : cap);
    static const uint64_t small_n = [] {                 // experiment switch: jobs below this take the checked loop only
        const char* e = getenv("JMC_FAST_SMALL_N");
        return e ? (uint64_t)atof(e) : (uint64_t)0;
    }();
    FastParams Q = P;
    Q.spread_shift = shift;
    Q.checked_only = n < small_n;
    if (!d_edges_f) Q.n_edges = 0;
    kern<<<grid, threads, smem, st>>>(Q, n, seed, first, d_edges_f, d_sum_w, d_sum_w2, (unsigned long long*)d_count,
                                      d_poison, d_minmax);
    JMC_LAUNCHED("run_fast_kernel");
    return JMC_OK;
}

template <int KIND, bool FIXED>
static int launch_fast(const FastParams& P, uint64_t n, uint64_t seed, uint64_t first, const float* d_edges_f,
                       double* d_sum_w, double* d_sum_w2, uint64_t* d_count, int* d_poison, double* d_minmax,
                       cudaStream_t st) {
    const bool hot = d_edges_f != nullptr && d_minmax == nullptr && P.dom_log && P.uniform
                     && (P.nan_to_num || FIXED || KIND == JMC_KIND_RADIATOR);
    const bool big = hot && fast_smem_bytes(P.n_edges) > kBigSmemThreshold;
#define JMC_GO(M, H, B) launch_fast_t<KIND, FIXED, M, H, B>(P, n, seed, first, d_edges_f, d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, st)
    // counts over the weighted samples only (jmc_integrand.counts): the kinds with a Landau region skip observable,
    // bin and count atomic of the samples whose weight is NaN -> 0
    constexpr bool CAN_LAZY = (KIND == JMC_KIND_CRIT_SUB || KIND == JMC_KIND_CRIT) && !FIXED;
    if (CAN_LAZY && hot && !big && P.count_weighted) {
        if (P.obs_mll)
            return launch_fast_t<KIND, FIXED, true, true, false, !CAN_LAZY>(P, n, seed, first, d_edges_f, d_sum_w, d_sum_w2,
                                                                            d_count, d_poison, d_minmax, st);
        return launch_fast_t<KIND, FIXED, false, true, false, !CAN_LAZY>(P, n, seed, first, d_edges_f, d_sum_w, d_sum_w2,
                                                                         d_count, d_poison, d_minmax, st);
    }
    if (P.obs_mll) return hot ? (big ? JMC_GO(true, true, true) : JMC_GO(true, true, false)) : JMC_GO(true, false, false);
    return hot ? (big ? JMC_GO(false, true, true) : JMC_GO(false, true, false)) : JMC_GO(false, false, false);
#undef JMC_GO
}

// ---- a job prepared on the host: kernel parameters + FP32 edges in the compare domain ------------------------
struct FastPlan {
    jmc_integrand g;
    FastParams P;
    std::vector<float> edges_f;
    int n_edges;
};

int fast_plan_make(const jmc_integrand* g, const double* h_edges, int n_edges, FastPlan** out) {
    if (h_edges) JMC_REQUIRE(n_edges >= 2, "jmc_run: need at least 2 edges");
    FastPlan* p = new FastPlan;
    p->g = *g;
    p->n_edges = h_edges ? n_edges : 0;
    int rc = make_fast_params(g, h_edges, p->n_edges, &p->P, &p->edges_f);
    if (rc) { delete p; return rc; }
    *out = p;
    return JMC_OK;
}
void fast_plan_free(FastPlan* p) { delete p; }
const float* fast_plan_edges(const FastPlan* p) { return p->edges_f.data(); }

// d_edges_f: the plan's FP32 edges on the device (NULL for a min/max-only pass); d_poison: FOUR ints -- [0] preset to
// 0x7f7f7f7f ("no poison"), [1] (work counter) and [2] (finished blocks) preset to 0 -- all left in that state.
// Asynchronous on the stream, no host synchronisation.
int fast_plan_run(const FastPlan* plan, const float* d_edges_f, int* d_poison, uint64_t n, uint64_t seed,
                  uint64_t first_index, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax,
                  void* stream, bool defer_poison) {
    cudaStream_t st = (cudaStream_t)stream;
    const jmc_integrand* g = &plan->g;
    const FastParams& P = plan->P;
    if (d_edges_f) JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_run: histogram outputs missing");
    else JMC_REQUIRE(d_minmax != nullptr, "jmc_run: neither edges nor minmax requested");
    int rc = JMC_OK;
#define JMC_RUNF(K)                                                                                                   \
    rc = g->fixed_coupling                                                                                            \
             ? launch_fast<K, true>(P, m, seed, first_index, d_edges_f, d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, st) \
             : launch_fast<K, false>(P, m, seed, first_index, d_edges_f, d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, st)
    // one launch per stretch of the sample stream that stays inside one multiple of 2^32 (fast_body's 32-bit index)
    while (n) {
        const uint64_t m = std::min<uint64_t>(n, fast_span(first_index));
        switch (g->kind) {
            case JMC_KIND_RADIATOR: JMC_RUNF(JMC_KIND_RADIATOR); break;
            case JMC_KIND_CRIT: JMC_RUNF(JMC_KIND_CRIT); break;
            case JMC_KIND_CRIT_SUB: JMC_RUNF(JMC_KIND_CRIT_SUB); break;
            case JMC_KIND_PRE_CRIT_SUB:
                rc = launch_fast<JMC_KIND_PRE_CRIT_SUB, true>(P, m, seed, first_index, d_edges_f, d_sum_w, d_sum_w2,
                                                              d_count, d_poison, d_minmax, st);
                break;
            default: return set_error(JMC_EINVAL, "jmc_run: kind %d has no FP32 form", g->kind);
        }
        if (rc) return rc;
        first_index += m;
        n -= m;
    }
#undef JMC_RUNF
    if (d_edges_f && !g->nan_to_num && !defer_poison) {
        poison_fast_kernel<<<1, 256, 0, st>>>(d_poison, plan->n_edges - 1, d_sum_w, d_sum_w2);
        JMC_LAUNCHED("poison_fast_kernel");
    }
    return JMC_OK;
}

// jmc_run with DEVICE edges: the FP32 kernel bins against FP32 edges in its compare domain (log2 or linear), derived
// on the host from the caller's FP64 edges -- one synchronous 8*n_edges byte read-back and one upload per call.
// (jmc_handle_run takes host edges, caches the derived forms and never synchronises.)
int run_fast(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, const double* d_edges,
             int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (d_edges) {
        JMC_REQUIRE(n_edges >= 2, "jmc_run: need at least 2 edges");
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_run: histogram outputs missing");
    } else {
        JMC_REQUIRE(d_minmax != nullptr, "jmc_run: neither edges nor minmax requested");
    }
    std::vector<double> h_edges;
    if (d_edges) {
        h_edges.resize(n_edges);
        JMC_CUDA(cudaMemcpyAsync(h_edges.data(), d_edges, sizeof(double) * n_edges, cudaMemcpyDeviceToHost, st));
        JMC_CUDA(cudaStreamSynchronize(st));
    }
    FastPlan* plan = nullptr;
    int rc = fast_plan_make(g, d_edges ? h_edges.data() : nullptr, n_edges, &plan);
    if (rc) return rc;
    if (n == 0) { fast_plan_free(plan); return JMC_OK; }
    char* scr;
    rc = scratch(256 + sizeof(float) * (size_t)(n_edges > 0 ? n_edges : 1), (void**)&scr, (void*)st);
    if (rc) { fast_plan_free(plan); return rc; }
    int* d_poison = (int*)scr;
    float* d_edges_f = nullptr;
    cudaError_t e = cudaMemsetAsync(d_poison, 0x7f, sizeof(int), st);          // [0] poison flag: "none"
    if (e == cudaSuccess) e = cudaMemsetAsync(d_poison + 1, 0, 3 * sizeof(int), st);    // [1] work counter, [2] finished blocks
    if (e == cudaSuccess && d_edges) {
        d_edges_f = (float*)(scr + 256);
        e = cudaMemcpyAsync(d_edges_f, plan->edges_f.data(), sizeof(float) * n_edges, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // the plan's host edges die at return
    }
    if (e != cudaSuccess) { fast_plan_free(plan); return cuda_fail(e, "jmc_run: staging the FP32 edges"); }
    rc = fast_plan_run(plan, d_edges_f, d_poison, n, seed, first_index, d_sum_w, d_sum_w2, d_count, d_minmax, stream);
    fast_plan_free(plan);
    return rc;
}

__global__ void poison_batch_kernel(const int* poison_all, int nb, double* sum_w_all, double* sum_w2_all) {
    const int point = blockIdx.x;
    const int pb = poison_all[4 * point];
    if (pb < 0 || pb >= nb) return;
    for (int k = pb + threadIdx.x; k < nb; k += blockDim.x) {
        sum_w_all[(size_t)point * nb + k] = __longlong_as_double(0x7ff8000000000000LL);
        sum_w2_all[(size_t)point * nb + k] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

template <int KIND, bool FIXED, bool OBS_MLL, bool HOT, bool BIG>
static int launch_fast_batch_t(const FastParams* d_params, int n_points, int n_edges, uint64_t n, uint64_t seed,
                               uint64_t first, const float* d_edges_f, double* d_sum_w, double* d_sum_w2,
                               uint64_t* d_count, int* d_poison, double* d_minmax, int shift, cudaStream_t st) {
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    constexpr int threads = BIG ? kFastThreadsBig : kFastThreads;
    const size_t smem = fast_smem_bytes(d_edges_f ? n_edges : 0, threads, d_edges_f ? shift : 0);
    if (smem > (size_t)di.max_smem_optin)
        return set_error(JMC_ELIMIT, "jmc_run_batch: %d edges need %zu B of shared memory (limit %d B)", n_edges, smem,
                         di.max_smem_optin);
    auto kern = run_fast_batch_kernel<KIND, FIXED, OBS_MLL, HOT, BIG>;
    JMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    JMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)di.sm_count * per_sm;
    // enough blocks for ~8 full waves (the tail wave then costs little), but at least ~128 samples per thread so
    // that zeroing and flushing the block histogram stays negligible
    uint64_t per_point = (8 * cap + (uint64_t)n_points - 1) / (uint64_t)n_points;
    const uint64_t most = n / ((uint64_t)threads * 128);
    if (per_point > most) per_point = most;
    if (per_point < 1) per_point = 1;
    dim3 grid((unsigned)per_point, (unsigned)n_points);
    kern<<<grid, threads, smem, st>>>(d_params, n, seed, first, d_edges_f, n_edges, d_sum_w, d_sum_w2,
                                      (unsigned long long*)d_count, d_poison, d_minmax);
    JMC_LAUNCHED("run_fast_batch_kernel");
    return JMC_OK;
}

int run_fast_batch(const jmc_integrand* g, int n_points, uint64_t n, uint64_t seed, uint64_t first_index,
                   const double* h_edges, int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count,
                   double* d_minmax, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    JMC_REQUIRE(g != nullptr && n_points >= 1 && n_points <= 65535, "jmc_run_batch: need 1..65535 parameter points");
    if (h_edges) {
        JMC_REQUIRE(n_edges >= 2, "jmc_run_batch: need at least 2 edges");
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_run_batch: histogram outputs missing");
    } else {
        JMC_REQUIRE(d_minmax != nullptr, "jmc_run_batch: neither edges nor minmax requested");
        n_edges = 0;
    }
    std::vector<FastParams> params(n_points);
    // FP32 edges of every point, staged in pinned memory (100 MB for the reference's 5000 x 5000 grid)
    const size_t n_edges_all = (size_t)n_points * (size_t)n_edges;
    float* edges_all = nullptr;
    if (n_edges_all) {
        int prc = pinned_scratch(sizeof(float) * n_edges_all, (void**)&edges_all);
        if (prc) return prc;
    }
    for (int i = 0; i < n_points; ++i)
        JMC_REQUIRE(g[i].kind == g[0].kind && g[i].fixed_coupling == g[0].fixed_coupling && g[i].obs_acc == g[0].obs_acc,
                    "jmc_run_batch: all points must share kind, coupling and observable accuracy (point %d differs)", i);
    // Per point: kernel parameters and the FP32 compare-domain edges (a log2 per edge: 25 M of them for the
    // reference's 5000 x 5000 grid -- 0.2 s on one core, which was two thirds of that histogram pass).  Points are
    // independent: split over the host's cores.  A failing point is redone on the calling thread for its message.
    std::atomic<int> first_bad(n_points);
    auto plan_points = [&](int lo, int hi) {
        std::vector<float> one;
        for (int i = lo; i < hi && first_bad.load(std::memory_order_relaxed) > i; ++i) {
            const int rc = make_fast_params(&g[i], h_edges ? h_edges + (size_t)i * n_edges : nullptr, n_edges, &params[i], &one);
            if (rc) {
                int seen = first_bad.load();
                while (i < seen && !first_bad.compare_exchange_weak(seen, i)) {}
                return;
            }
            if (h_edges) std::memcpy(&edges_all[(size_t)i * n_edges], one.data(), sizeof(float) * n_edges);
            if (!h_edges) params[i].n_edges = 0;
        }
    };
    const size_t work = (size_t)n_points * (size_t)(n_edges > 0 ? n_edges : 1);
    unsigned workers = work < (1u << 18) ? 1u : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (workers > (unsigned)n_points) workers = (unsigned)n_points;
    if (workers <= 1) {
        plan_points(0, n_points);
    } else {
        std::vector<std::thread> pool;
        const int per = (n_points + (int)workers - 1) / (int)workers;
        for (unsigned t = 0; t < workers; ++t) {
            const int lo = (int)t * per, hi = std::min(n_points, lo + per);
            if (lo < hi) pool.emplace_back(plan_points, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    if (first_bad.load() < n_points) {
        std::vector<float> one;
        const int i = first_bad.load();
        const int rc = make_fast_params(&g[i], h_edges ? h_edges + (size_t)i * n_edges : nullptr, n_edges, &params[i], &one);
        return rc ? rc : set_error(JMC_EINVAL, "jmc_run_batch: point %d could not be planned", i);
    }
    bool hot = h_edges != nullptr && d_minmax == nullptr;
    for (int i = 0; i < n_points; ++i)
        hot = hot && params[i].dom_log && params[i].uniform
              && (params[i].nan_to_num || g[0].fixed_coupling || g[0].kind == JMC_KIND_RADIATOR);
    if (n == 0) return JMC_OK;
    const bool big = hot && fast_smem_bytes(n_edges) > kBigSmemThreshold;
    const int shift = h_edges ? fast_spread_shift(n_edges, big ? kFastThreadsBig : kFastThreads,
                                                  big ? kSpreadBudgetBig : kSpreadBudget) : 0;
    for (int i = 0; i < n_points; ++i) params[i].spread_shift = shift;
    const size_t bytes_params = sizeof(FastParams) * (size_t)n_points;
    const size_t off_edges = (bytes_params + 255) & ~(size_t)255;
    const size_t bytes_edges = sizeof(float) * n_edges_all;
    const size_t off_poison = (off_edges + bytes_edges + 255) & ~(size_t)255;
    char* scr;
    int rc = scratch(off_poison + 4 * sizeof(int) * (size_t)n_points, (void**)&scr, (void*)st);
    if (rc) return rc;
    FastParams* d_params = (FastParams*)scr;
    float* d_edges_f = h_edges ? (float*)(scr + off_edges) : nullptr;
    int* d_poison = (int*)(scr + off_poison);
    JMC_CUDA(cudaMemcpyAsync(d_params, params.data(), bytes_params, cudaMemcpyHostToDevice, st));
    if (h_edges) JMC_CUDA(cudaMemcpyAsync(d_edges_f, edges_all, bytes_edges, cudaMemcpyHostToDevice, st));
    // per point: [poison flag = "none", work counter = 0, finished blocks = 0, unused]
    JMC_CUDA(cudaMemsetAsync(d_poison, 0, 4 * sizeof(int) * (size_t)n_points, st));
    JMC_CUDA(cudaMemset2DAsync(d_poison, 4 * sizeof(int), 0x7f, sizeof(int), (size_t)n_points, st));
    JMC_CUDA(cudaStreamSynchronize(st));           // the staging vectors die at return
    const bool mll = g[0].obs_acc == JMC_ACC_MLL, fixed = g[0].fixed_coupling != 0;
#define JMC_B(K, F, M, H)                                                                                              \
    rc = (H && big) ? launch_fast_batch_t<K, F, M, H, H>(d_params, n_points, n_edges, n, seed, first_index, d_edges_f, \
                                                         d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, shift, st)      \
                    : launch_fast_batch_t<K, F, M, H, false>(d_params, n_points, n_edges, n, seed, first_index,         \
                                                             d_edges_f, d_sum_w, d_sum_w2, d_count, d_poison, d_minmax, \
                                                             shift, st)
#define JMC_BK(K)                                                              \
    do {                                                                       \
        if (fixed) { if (mll) { if (hot) JMC_B(K, true, true, true); else JMC_B(K, true, true, false); }     \
                     else { if (hot) JMC_B(K, true, false, true); else JMC_B(K, true, false, false); } }      \
        else { if (mll) { if (hot) JMC_B(K, false, true, true); else JMC_B(K, false, true, false); }         \
               else { if (hot) JMC_B(K, false, false, true); else JMC_B(K, false, false, false); } }          \
    } while (0)
    const uint64_t n_all = n;
    while (n_all && n) {
        const uint64_t n_left = n;
        n = std::min<uint64_t>(n_left, fast_span(first_index));
        switch (g[0].kind) {
            case JMC_KIND_RADIATOR: JMC_BK(JMC_KIND_RADIATOR); break;
            case JMC_KIND_CRIT: JMC_BK(JMC_KIND_CRIT); break;
            case JMC_KIND_CRIT_SUB: JMC_BK(JMC_KIND_CRIT_SUB); break;
            default: return set_error(JMC_EINVAL, "jmc_run_batch: kind %d is not batched", g[0].kind);
        }
        if (rc) return rc;
        first_index += n;
        n = n_left - n;
    }
#undef JMC_BK
#undef JMC_B
    bool any_nan_semantics = false;
    for (int i = 0; i < n_points; ++i) any_nan_semantics = any_nan_semantics || !g[i].nan_to_num;
    if (h_edges && any_nan_semantics) {
        poison_batch_kernel<<<n_points, 128, 0, st>>>(d_poison, n_edges - 1, d_sum_w, d_sum_w2);
        JMC_LAUNCHED("poison_batch_kernel");
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
#define JMC_DUMPF(K)                                                                                              \
    do {                                                                                                          \
        if (g->fixed_coupling)                                                                                    \
            dump_fast_kernel<K, true><<<grid, kFastThreads, 0, st>>>(P, n, seed, first_index, d_samples, d_obs, d_weight); \
        else                                                                                                      \
            dump_fast_kernel<K, false><<<grid, kFastThreads, 0, st>>>(P, n, seed, first_index, d_samples, d_obs, d_weight); \
    } while (0)
    switch (g->kind) {
        case JMC_KIND_RADIATOR: JMC_DUMPF(JMC_KIND_RADIATOR); break;
        case JMC_KIND_CRIT: JMC_DUMPF(JMC_KIND_CRIT); break;
        case JMC_KIND_CRIT_SUB: JMC_DUMPF(JMC_KIND_CRIT_SUB); break;
        case JMC_KIND_PRE_CRIT_SUB:
            dump_fast_kernel<JMC_KIND_PRE_CRIT_SUB, true><<<grid, kFastThreads, 0, st>>>(P, n, seed, first_index, d_samples,
                                                                                         d_obs, d_weight);
            break;
        default: return set_error(JMC_EINVAL, "jmc_dump: kind %d has no FP32 form", g->kind);
    }
#undef JMC_DUMPF
    JMC_LAUNCHED("dump_fast_kernel");
    return JMC_OK;
}

}  // namespace jmc
