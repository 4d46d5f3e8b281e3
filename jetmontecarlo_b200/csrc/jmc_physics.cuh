// FP64 device evaluators of the jet observables, splitting functions,
// radiators and phase-space weights, written in NumPy operation order so that
// the replay path reproduces the reference to rounding: every region of the
// running-coupling radiators is evaluated and then masked (bool * value), which
// is what makes the reference return NaN below the Landau scale (log of a
// negative number times a zero mask is NaN) -- parity means reproducing that.
//
// This header must be compiled with --fmad=false (see csrc/build.py): FMA
// contraction would change the rounding of the observables and with it the
// bin a sample lands in.
//
// ref (paths relative to /root/reference/jetmontecarlo):
//   analytics/qcd_utils.py, analytics/radiators/fixedcoupling.py,
//   analytics/radiators/running_coupling.py, numerics/observables.py,
//   numerics/weights.py
#pragma once
#include <cmath>
#include <cstdint>
#include "../../include/jmc_b200.h"

namespace jmc {

// ---- constants                                   ref: analytics/qcd_utils.py:33-66,195-206
// Literals are the FP64 values NumPy produces (repr round-trips exactly).
#define JMC_CF 1.3333333333333333        /* 8./6. */
#define JMC_CA 3.0
#define JMC_TF 0.5
#define JMC_NF 5.0
#define JMC_BETA0 0.6100939485189322      /* (33 - 2 N_F)/(12 pi) */
#define JMC_MZ 91.19
#define JMC_PT 3000.0
#define JMC_R0 1.0
#define JMC_ALPHA_MZ 0.118
#define JMC_ALPHA_FIXED 0.0785101418528313 /* alpha1loop(P_T * R0) */
#define JMC_MU_NP 0.0003333333333333333    /* 1/(P_T R0) */
#define JMC_PI 3.141592653589793
#define JMC_TINY 1e-100
#define JMC_ALPHA_VETO 1.5

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// np.maximum / np.minimum propagate NaN (fmax/fmin do not)
__device__ __forceinline__ double np_max(double a, double b) {
    return (a != a || b != b) ? qnan() : fmax(a, b);
}
__device__ __forceinline__ double np_min(double a, double b) {
    return (a != a || b != b) ? qnan() : fmin(a, b);
}
__device__ __forceinline__ double sq(double x) { return x * x; }
// ndarray ** python-float: NumPy's scalar-exponent fast paths, else pow
__device__ __forceinline__ double np_pow(double x, double y) {
    if (y == 2.0) return x * x;
    if (y == 1.0) return x;
    if (y == 0.5) return sqrt(x);
    if (y == -1.0) return 1.0 / x;
    if (y == 0.0) return 1.0;
    return pow(x, y);
}
__device__ __forceinline__ double msk(bool c) { return c ? 1.0 : 0.0; }

__device__ __forceinline__ double casimir(int jet) { return jet == JMC_QUARK ? JMC_CF : JMC_CA; }

// ref: analytics/qcd_utils.py:102-105
__device__ __forceinline__ double alpha_one_loop(double mu) {
    return JMC_ALPHA_MZ / (1.0 + 2.0 * JMC_ALPHA_MZ * JMC_BETA0 * log(mu / JMC_MZ));
}
// ref: analytics/qcd_utils.py:115-119
__device__ __forceinline__ double alpha_running(double z, double theta) {
    return alpha_one_loop(np_max(JMC_PT * z * theta, 1.0));
}

// ---- splitting functions                          ref: analytics/qcd_utils.py:124-190
__device__ __forceinline__ double p_quark(double z) { return JMC_CF * ((1.0 + sq(1.0 - z)) / z); }
__device__ __forceinline__ double p_gluon(double z) {
    return JMC_CA * (2.0 * (1.0 - z) / z + z * (1.0 - z)
                     + (JMC_TF * JMC_NF) * (sq(z) + sq(1.0 - z)) / JMC_CA);
}
__device__ __forceinline__ double splitting(double z, int jet, int acc) {
    if (acc == JMC_ACC_LL) return 2.0 * casimir(jet) / z;
    if (jet == JMC_QUARK) {
        if (acc == JMC_ACC_LL_NONSING) return p_quark(z);
        return p_quark(z) + p_quark(1.0 - z);
    }
    if (acc == JMC_ACC_LL_NONSING) return p_gluon(z);
    return p_gluon(z) + p_gluon(1.0 - z);
}

// ref: analytics/qcd_utils.py:223-236
__device__ __forceinline__ double b_bar(double zc, int jet) {
    if (jet == JMC_QUARK) return JMC_CF * (-3.0 + 6.0 * zc + 4.0 * log(2.0 - 2.0 * zc)) / 2.0;
    return ((-1.0 + 2.0 * zc) * (-4.0 * JMC_NF * JMC_TF * (1.0 + (-1.0 + zc) * zc)
                                 + JMC_CA * (11.0 + 2.0 * (-1.0 + zc) * zc)) / 6.0
            + 2.0 * JMC_CA * log(2.0 - 2.0 * zc));
}
// W(x) = x ln x, Lambda(x, alpha) = 2 alpha beta0 ln x      ref: qcd_utils.py:211-221
__device__ __forceinline__ double xlogx(double x) { return x * log(x); }
__device__ __forceinline__ double lam(double x, double alpha) { return 2.0 * alpha * JMC_BETA0 * log(x); }
// (coeff*piece)*alpha/(pi*2*alpha*beta0)          ref: e.g. running_coupling.py:24-29
__device__ __forceinline__ double colour_norm(double piece, double coeff, double alpha) {
    return (coeff * piece) * alpha / (JMC_PI * 2.0 * alpha * JMC_BETA0);
}
__device__ __forceinline__ double two_cr(int jet) { return jet == JMC_QUARK ? 2 * JMC_CF : 2 * JMC_CA; }

// ---- observables                                  ref: numerics/observables.py:15-114
__device__ __forceinline__ double ecf_groomed(double z_crit, double theta_crit, double z_cut, double beta,
                                              double z_pre, double f, int acc) {
    const double zc_left = z_cut - z_pre;
    const double zmin = np_min(z_crit, 1.0 - z_crit);
    if (acc == JMC_ACC_LL)
        return ((zmin - f * zc_left) * (1.0 - (1.0 - f) * zc_left) * np_pow(theta_crit, beta))
               / sq(1.0 - z_cut);
    return ((zmin - f * zc_left) * (1.0 - zmin - (1.0 - f) * zc_left) * np_pow(theta_crit, beta))
           / sq(1.0 - z_cut);
}
__device__ __forceinline__ double ecf_ungroomed(double z, double theta, double beta, int acc) {
    if (acc == JMC_ACC_LL) return z * np_pow(theta, beta);
    return z * (1.0 - z) * np_pow(theta, beta);
}

// ---- fixed-coupling analytics                     ref: analytics/radiators/fixedcoupling.py
// :59-72
__device__ __forceinline__ double pdf_crit_fc(double z, double theta, double z_c, int jet) {
    theta = theta + JMC_TINY;
    z_c = z_c + JMC_TINY;
    const double cr = casimir(jet);
    return (2.0 * cr * JMC_ALPHA_FIXED / (JMC_PI * theta * z)
            * exp(log(1.0 / (2.0 * z_c)) * log(theta / JMC_R0) * 2.0 * cr * JMC_ALPHA_FIXED / JMC_PI)
            * msk(z_c < z) * msk(z < 1.0 / 2.0));
}
// :7-19
__device__ __forceinline__ double rad_crit_fc(double theta, double z_c, int jet) {
    theta = theta + JMC_TINY;
    z_c = z_c + JMC_TINY;
    return (-2.0 * casimir(jet) * log(JMC_R0 / theta) * log(2.0 * z_c) * JMC_ALPHA_FIXED / JMC_PI);
}
// :77-90  (max_radius ** beta with the array fast path when it is per-sample)
__device__ __forceinline__ double rad_sub_fc(double C, double beta, int jet, double max_radius) {
    C = C + JMC_TINY;
    const double rb = np_pow(max_radius, beta);
    const double rad = casimir(jet) * JMC_ALPHA_FIXED / (beta * JMC_PI) * sq(log(2.0 * C / rb));
    return rad * msk(C < rb / 2.0);
}
// :92-108
__device__ __forceinline__ double radprime_sub_fc(double C, double beta, int jet, double max_radius) {
    C = C + JMC_TINY;
    const double rb = np_pow(max_radius, beta);
    const double d = 2.0 * casimir(jet) * JMC_ALPHA_FIXED / (beta * JMC_PI) * log(2.0 * C / rb) / C;
    return d * msk(C < rb / 2.0);
}
// :162-176 (the +1e-100 is applied again inside both callees, as in the reference)
__device__ __forceinline__ double pdf_sub_fc(double C, double beta, int jet, double max_radius) {
    C = C + JMC_TINY;
    const double sud = exp(-rad_sub_fc(C, beta, jet, max_radius));
    return sud * (-radprime_sub_fc(C, beta, jet, max_radius));
}
// :181-192
__device__ __forceinline__ double rad_pre_fc(double z_pre, double theta_crit, double z_cut, int jet,
                                             double max_radius) {
    const double rad = 2 * casimir(jet) * JMC_ALPHA_FIXED / JMC_PI * log(z_pre / z_cut)
                       * log(theta_crit / max_radius);
    return rad * msk(0.0 < z_pre) * msk(z_pre < z_cut) * msk(0.0 < theta_crit)
           * msk(theta_crit < max_radius);
}

// ---- running-coupling analytics                   ref: analytics/radiators/running_coupling.py
// subsequent emissions, C space.  :11-29
__device__ __forceinline__ double sub_sc1(double C, double beta, int jet, double alpha) {
    const double sc = beta / (2.0 * alpha * JMC_BETA0 * (beta - 1.0))
                      * (xlogx(1.0 + lam(0.5, alpha)) * (1.0 - 1.0 / beta)
                         - xlogx(1.0 + lam(C / pow(2.0, beta - 1.0), alpha) / beta)
                         + xlogx(1.0 + lam(C, alpha)) / beta);
    return colour_norm(sc, two_cr(jet), alpha);
}
// :31-56
__device__ __forceinline__ double sub_sc2(double C, double beta, int jet, double alpha) {
    const double pert = beta / (2.0 * alpha * JMC_BETA0 * (beta - 1.0))
                        * (xlogx(1.0 + lam(0.5, alpha) * (1.0 - 1.0 / beta) + lam(JMC_MU_NP, alpha) / beta)
                           - xlogx(1.0 + lam(0.5, alpha) * (1.0 - 1.0 / beta) + lam(C, alpha) / beta)
                           + lam(C / JMC_MU_NP, alpha) / beta * (1 + log(1 + lam(JMC_MU_NP, alpha))));
    const double nonpert = 2.0 * alpha * JMC_BETA0 * sq(log(JMC_MU_NP / C)) / (2.0 * (beta - 1.0));
    return colour_norm(pert + nonpert, two_cr(jet), alpha);
}
// :58-72
__device__ __forceinline__ double sub_sc3(double C, double beta, int jet, double alpha) {
    const double sc = (sq(log(2.0 * C)) / beta - beta * sq(log(2.0 * JMC_MU_NP))) * 2.0 * alpha * JMC_BETA0 / 2.0;
    return colour_norm(sc, two_cr(jet), alpha);
}
// :77-95
__device__ __forceinline__ double sub_hc1(double C, double beta, int jet, double alpha) {
    const double hc = log(1 + lam(0.5, alpha)) - log(1 + lam(C / pow(2.0, beta - 1.0), alpha) / beta);
    return colour_norm(hc, b_bar(C, jet), alpha);
}
// :97-117
__device__ __forceinline__ double sub_hc2(double C, double beta, int jet, double alpha) {
    const double hc = 2.0 * alpha * JMC_BETA0 * log(pow(JMC_MU_NP, beta) * pow(2.0, beta - 1) / C)
                      / ((1.0 + lam(JMC_MU_NP, alpha)) * beta);
    return colour_norm(hc, b_bar(C, jet), alpha);
}
// :119-148, every region evaluated and masked as the reference writes it
__device__ __forceinline__ double rad_sub_rc_all_regions(double C, double beta, int jet, double max_radius) {
    C = C / np_pow(max_radius, beta);
    const double alpha = alpha_one_loop(np_max(JMC_PT * max_radius, 1.0));
    const double knee = pow(JMC_MU_NP, beta) * pow(2.0, beta - 1.0);
    const double soft =
        msk(C > JMC_MU_NP) * sub_sc1(C, beta, jet, alpha)
        + msk((knee < C) && (C < JMC_MU_NP)) * (sub_sc1(JMC_MU_NP, beta, jet, alpha) + sub_sc2(C, beta, jet, alpha))
        + msk(C < knee) * (sub_sc1(JMC_MU_NP, beta, jet, alpha) + sub_sc2(knee, beta, jet, alpha)
                           + sub_sc3(C, beta, jet, alpha));
    const double hard = msk(C > knee) * sub_hc1(C, beta, jet, alpha)
                        + msk(C < knee) * (sub_hc1(knee, beta, jet, alpha) + sub_hc2(C, beta, jet, alpha));
    return (soft + hard) * msk(C < 1.0 / 2.0);
}

// The same value from the ACTIVE region only.  The reference's sum  m1*X1 + m2*X2 + m3*X3  of masked regions equals the
// one region whose mask is 1 exactly (0 * finite = +-0, and x + 0 = x) -- unless some region's expression is NaN, which
// happens iff one of its logarithm arguments is negative.  Those arguments are, for every region,
// 1 + Lambda(MU_NP), 1 + Lambda(C) and 2 - 2C (b_bar) or larger (for beta > 1; derivation in DESIGN.md 4.2), so: test
// the three, take the all-regions form when one fails (it yields the reference's NaN), else evaluate one region
// (5 logarithms instead of ~25).  ACTIVE = false is the all-regions form, instantiated by the host for parameters
// outside that derivation (beta <= 1, z_c >= 1/2): a launch-uniform choice, made where the kernel is picked.
template <bool ACTIVE = true>
__device__ __forceinline__ double rad_sub_rc(double C_in, double beta, int jet, double max_radius) {
    if (!ACTIVE) return rad_sub_rc_all_regions(C_in, beta, jet, max_radius);
    const double C = C_in / np_pow(max_radius, beta);
    const double alpha = alpha_one_loop(np_max(JMC_PT * max_radius, 1.0));
    if (!(1.0 + lam(JMC_MU_NP, alpha) > 0.0) || !(1.0 + lam(C, alpha) > 0.0) || !(2.0 - 2.0 * C > 0.0))
        return qnan();                                    // a negative logarithm argument in some region: NaN
    const double knee = pow(JMC_MU_NP, beta) * pow(2.0, beta - 1.0);
    double soft = 0.0, hard = 0.0;
    if (C > JMC_MU_NP) soft = sub_sc1(C, beta, jet, alpha);
    else if ((knee < C) && (C < JMC_MU_NP)) soft = sub_sc1(JMC_MU_NP, beta, jet, alpha) + sub_sc2(C, beta, jet, alpha);
    else if (C < knee) soft = sub_sc1(JMC_MU_NP, beta, jet, alpha) + sub_sc2(knee, beta, jet, alpha) + sub_sc3(C, beta, jet, alpha);
    if (C > knee) hard = sub_hc1(C, beta, jet, alpha);
    else if (C < knee) hard = sub_hc1(knee, beta, jet, alpha) + sub_hc2(C, beta, jet, alpha);
    return (soft + hard) * msk(C < 1.0 / 2.0);
}

// critical emissions, theta space.  :157-175
__device__ __forceinline__ double crit_sc1(double theta, double z_c, int jet, double alpha) {
    const double sc = (xlogx(1.0 + lam(0.5, alpha)) - xlogx(1.0 + lam(z_c, alpha))
                       - xlogx(1.0 + lam(theta / 2.0, alpha)) + xlogx(1.0 + lam(z_c * theta, alpha)))
                      / (2.0 * alpha * JMC_BETA0);
    return colour_norm(sc, two_cr(jet), alpha);
}
// :177-199
__device__ __forceinline__ double crit_sc2(double theta, double z_c, int jet, double alpha) {
    const double pert = (xlogx(1.0 + lam(JMC_MU_NP / (2.0 * z_c), alpha)) - xlogx(1 + lam(theta / 2.0, alpha))
                         + lam(theta * z_c / JMC_MU_NP, alpha) * (1.0 + log(1 + lam(JMC_MU_NP, alpha))))
                        / (2.0 * alpha * JMC_BETA0);
    const double nonpert = (2.0 * alpha * JMC_BETA0) * sq(log(theta * z_c / JMC_MU_NP))
                           / (2.0 * (1.0 + lam(JMC_MU_NP, alpha)));
    return colour_norm(pert + nonpert, two_cr(jet), alpha);
}
// :201-219
__device__ __forceinline__ double crit_sc3(double theta, double z_c, int jet, double alpha) {
    const double sc = -(2.0 * alpha * JMC_BETA0) * log(2.0 * z_c) * log(2.0 * JMC_MU_NP / theta)
                      / (1.0 + lam(JMC_MU_NP, alpha));
    return colour_norm(sc, two_cr(jet), alpha);
}
// :224-240
__device__ __forceinline__ double crit_hc1(double theta, double z_c, int jet, double alpha) {
    const double hc = -log(1.0 + lam(theta, alpha));
    return colour_norm(hc, b_bar(z_c, jet), alpha);
}
// :242-258
__device__ __forceinline__ double crit_hc2(double theta, double z_c, int jet, double alpha) {
    const double hc = lam(JMC_MU_NP * 2.0 / theta, alpha) / (1.0 + lam(JMC_MU_NP, alpha));
    return colour_norm(hc, b_bar(z_c, jet), alpha);
}
// :263-289.  The third soft region calls critSC2/critSC3 with the DEFAULT alpha
// (:276-277); kept.
__device__ __forceinline__ double rad_crit_rc_all_regions(double theta, double z_c, int jet) {
    const double alpha = JMC_ALPHA_FIXED;
    const double soft =
        msk(theta > JMC_MU_NP / z_c) * crit_sc1(theta, z_c, jet, alpha)
        + msk((2.0 * JMC_MU_NP < theta) && (theta < JMC_MU_NP / z_c))
              * (crit_sc1(JMC_MU_NP / z_c, z_c, jet, alpha) + crit_sc2(theta, z_c, jet, alpha))
        + msk(theta < 2.0 * JMC_MU_NP)
              * (crit_sc1(JMC_MU_NP / z_c, z_c, jet, alpha) + crit_sc2(2.0 * JMC_MU_NP, z_c, jet, JMC_ALPHA_FIXED)
                 + crit_sc3(theta, z_c, jet, JMC_ALPHA_FIXED));
    const double hard = msk(theta > 2.0 * JMC_MU_NP) * crit_hc1(theta, z_c, jet, alpha)
                        + msk(theta < 2.0 * JMC_MU_NP)
                              * (crit_hc1(2.0 * JMC_MU_NP, z_c, jet, alpha) + crit_hc2(theta, z_c, jet, alpha));
    return soft + hard;
}
// The active region only (see rad_sub_rc).  The smallest logarithm arguments of all regions are 1 + Lambda(z_c theta)
// (z_c < 1/2) and the constants 1 + Lambda(z_c), 1 + Lambda(MU_NP).
// The values the lower regions carry over from the region boundaries depend on (z_c, jet) only: a kernel computes them
// once per thread (CritConsts) instead of once per sample and region.
struct CritConsts {
    double s1, s2, h1;     // critSC1(MU_NP / z_c), critSC2(2 MU_NP), critHC1(2 MU_NP)
};
__device__ __forceinline__ CritConsts make_crit_consts(double z_c, int jet) {
    CritConsts c;
    c.s1 = crit_sc1(JMC_MU_NP / z_c, z_c, jet, JMC_ALPHA_FIXED);
    c.s2 = crit_sc2(2.0 * JMC_MU_NP, z_c, jet, JMC_ALPHA_FIXED);
    c.h1 = crit_hc1(2.0 * JMC_MU_NP, z_c, jet, JMC_ALPHA_FIXED);
    return c;
}
template <bool ACTIVE = true>
__device__ __forceinline__ double rad_crit_rc(double theta, double z_c, int jet, const CritConsts* cc = nullptr) {
    if (!ACTIVE) return rad_crit_rc_all_regions(theta, z_c, jet);
    const double alpha = JMC_ALPHA_FIXED;
    if (!(1.0 + lam(z_c * theta, alpha) > 0.0) || !(1.0 + lam(z_c, alpha) > 0.0) || !(1.0 + lam(JMC_MU_NP, alpha) > 0.0))
        return qnan();
    const CritConsts local = cc ? *cc : make_crit_consts(z_c, jet);
    double soft = 0.0, hard = 0.0;
    if (theta > JMC_MU_NP / z_c) soft = crit_sc1(theta, z_c, jet, alpha);
    else if ((2.0 * JMC_MU_NP < theta) && (theta < JMC_MU_NP / z_c)) soft = local.s1 + crit_sc2(theta, z_c, jet, alpha);
    else if (theta < 2.0 * JMC_MU_NP) soft = local.s1 + local.s2 + crit_sc3(theta, z_c, jet, JMC_ALPHA_FIXED);
    if (theta > 2.0 * JMC_MU_NP) hard = crit_hc1(theta, z_c, jet, alpha);
    else if (theta < 2.0 * JMC_MU_NP) hard = local.h1 + crit_hc2(theta, z_c, jet, alpha);
    return soft + hard;
}
// :326-347
__device__ __forceinline__ double radprime_crit_rc(double theta, double z_c, int jet) {
    const double alpha = JMC_ALPHA_FIXED;
    const double m = np_max(z_c, np_min(JMC_MU_NP / theta, 0.5));
    const double sc = log((1 + 2.0 * alpha * JMC_BETA0 * log(theta / 2.0))
                          / (1 + 2.0 * alpha * JMC_BETA0 * log(m * theta)))
                          / (2.0 * alpha * JMC_BETA0)
                      + log(m / z_c) / (1 + 2.0 * alpha * JMC_BETA0 * log(JMC_MU_NP));
    const double hc = 1.0 / (1.0 + 2.0 * alpha * JMC_BETA0 * log(np_max(theta / 2.0, JMC_MU_NP)));
    return alpha * (two_cr(jet) * sc + b_bar(z_c, jet) * hc) / (JMC_PI * theta);
}
// :350-380
__device__ __forceinline__ double radprime_sub_rc(double c, double beta, int jet, double max_radius) {
    const double rb = np_pow(max_radius, beta);
    c = c / rb;
    const double jac = 1.0 / rb;
    const double alpha = alpha_one_loop(np_max(JMC_PT * max_radius, 1.0));
    const double m = np_max(c, np_min(pow(pow(JMC_MU_NP, beta) / c, 1.0 / (beta - 1.0)), 0.5));
    const double sc = log((1 + 2.0 * alpha * JMC_BETA0 * log(c / pow(2.0, beta - 1.0)) / beta)
                          / (1 + 2.0 * alpha * JMC_BETA0 * log(c * np_pow(m, beta - 1.0)) / beta))
                          * beta / ((beta - 1.0) * 2.0 * alpha * JMC_BETA0)
                      + log(m / c) / (1 + 2.0 * alpha * JMC_BETA0 * log(JMC_MU_NP));
    const double hc = 1.0 / (1.0 + 2.0 * alpha * JMC_BETA0
                                       * log(np_max(pow(2.0, -(beta - 1.0) / beta) * np_pow(c, 1 / beta), JMC_MU_NP)));
    return jac * alpha * (two_cr(jet) * sc + b_bar(c, jet) * hc) / (JMC_PI * beta * c);
}
// :386-404
template <bool ACTIVE = true>
__device__ __forceinline__ double pdf_crit_rc(double z, double theta, double z_c, int jet, const CritConsts* cc = nullptr) {
    return splitting(z, jet, JMC_ACC_MLL) * alpha_running(z, theta) * exp(-1.0 * rad_crit_rc<ACTIVE>(theta, z_c, jet, cc))
           / (JMC_PI * theta);
}
// :409-422
template <bool ACTIVE = true>
__device__ __forceinline__ double pdf_sub_rc(double c, double beta, int jet, double max_radius) {
    return radprime_sub_rc(c, beta, jet, max_radius) * exp(-rad_sub_rc<ACTIVE>(c, beta, jet, max_radius));
}

// ---- phase-space weights                          ref: numerics/weights.py
// :15-26
__device__ __forceinline__ double weight_radiator(double z, double theta, int jet, int fixed_coupling, int acc) {
    const double alpha = fixed_coupling ? JMC_ALPHA_FIXED : alpha_running(z, theta);
    return splitting(z, jet, acc) * alpha / (theta * JMC_PI);
}
// :31-41
template <bool ACTIVE = true>
__device__ __forceinline__ double weight_critical(double z, double theta, double z_c, int jet, int fixed_coupling,
                                                  const CritConsts* cc = nullptr) {
    return fixed_coupling ? pdf_crit_fc(z, theta, z_c, jet) : pdf_crit_rc<ACTIVE>(z, theta, z_c, jet, cc);
}
// :43-58 (fixed coupling only; the host wrapper raises TypeError otherwise)
__device__ __forceinline__ double weight_precritical(double z_pre, double theta_crit, double z_c, int jet) {
    const double cr = casimir(jet);
    return (2.0 * cr * JMC_ALPHA_FIXED / JMC_PI * log(JMC_R0 / theta_crit) / z_pre
            * exp(-2.0 * cr * JMC_ALPHA_FIXED / JMC_PI * log(JMC_R0 / theta_crit) * log(z_c / z_pre)))
           * msk(z_pre > 0.0) * msk(z_pre < z_c);
}
// np.nan_to_num defaults: NaN -> 0, +-inf -> +-DBL_MAX
__device__ __forceinline__ double nan_to_num(double x) {
    if (x != x) return 0.0;
    if (isinf(x)) return x > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return x;
}
// ref: /root/reference/examples/tests/comparison_tests/comparison_utils_probabilistic.py:82-98
__device__ __forceinline__ double weight_precritical_zerobin(double z_pre, double theta_crit, double z_c, int jet,
                                                             double epsilon) {
    const double cr = casimir(jet);
    const double zero = exp(-2.0 * cr * JMC_ALPHA_FIXED / JMC_PI * log(JMC_R0 / theta_crit) * log(z_c / epsilon))
                        * msk(z_pre == 0.0);
    const double w = weight_precritical(z_pre, theta_crit, z_c, jet) * msk(z_pre > 0.0);
    return zero + nan_to_num(w);
}
// :63-99 -- csub = z_s th_s^b th_s^b exactly as the reference writes it (:74)
__device__ __forceinline__ double csub_verbatim(double z_s, double th_s, double beta) {
    return z_s * np_pow(th_s, beta) * np_pow(th_s, beta);
}
template <bool ACTIVE = true>
__device__ __forceinline__ double weight_two_emission(double z_c_em, double th_c, double z_s, double th_s,
                                                      double z_c, double beta, int jet, int fixed_coupling,
                                                      const CritConsts* cc = nullptr) {
    const double csub = csub_verbatim(z_s, th_s, beta);
    double pc, ps;
    if (fixed_coupling) {
        pc = pdf_crit_fc(z_c_em, th_c, z_c, jet);
        ps = pdf_sub_fc(csub, beta, jet, th_c);
    } else {
        pc = pdf_crit_rc<ACTIVE>(z_c_em, th_c, z_c, jet, cc);
        ps = pdf_sub_rc<ACTIVE>(csub, beta, jet, th_c);
    }
    return pc * ps * (np_pow(th_s, beta) * np_pow(th_c, beta));
}

// ---- samplers                 ref: utils/montecarlo_utils.py:12-55, numerics/radiators/samplers.py
__device__ __forceinline__ double lin_sample(double lo, double hi, double u) { return lo + u * (hi - lo); }
// log(eps) and log(hi-lo) are per-launch constants computed on the device once
__device__ __forceinline__ double log_sample(double lo, double log_eps, double log_width, double u) {
    return lo + exp(log_eps * u + log_width);
}

// ---- np.histogram bin index               ref: numpy.histogram @ montecarlo/integrator.py:97-101
// idx = searchsorted(edges, x, 'right') - 1; x == edges[-1] -> last bin; out of
// range or NaN -> -1.
__device__ __forceinline__ int bin_index(double x, const double* __restrict__ edges, int n_edges) {
    if (!(x >= edges[0]) || !(x <= edges[n_edges - 1])) return -1;   // also drops NaN
    int lo = 0, hi = n_edges;                                        // first edge > x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] <= x) lo = mid + 1; else hi = mid;
    }
    int idx = lo - 1;
    if (idx > n_edges - 2) idx = n_edges - 2;                        // x == last edge
    return idx;
}

// Same result, fewer instructions, for edges that are (nearly) equally spaced in x or in log(x): an FP32 affine
// guess of the bin, then a walk against the stored FP64 edges -- which decides, so the result is the exact
// searchsorted one whatever the guess.  mode: 0 binary search, 1 linear guess, 2 log2 guess.
struct BinGuess {
    int mode;              // 0 none (binary search), 1 linear, 2 log-uniform, 3 log-uniform behind a zero-bin edge
    float b0, inv_d;       // guess = floor((g(x) - b0) * inv_d), g = identity or log2
    float eps;             // > 0: the guess is CERTAIN when the fractional part of the mapped value lies in
                           // (eps, 1 - eps) -- eps bounds every rounding error of the FP32 map (derive_bin_guess)
};
// MODE >= 0: the guess's mode as a compile-time constant (a streaming kernel branches on it once per block, not once
// per sample); MODE < 0: read it from G
template <int MODE = -1>
__device__ __forceinline__ int bin_index_guided(double x, const double* __restrict__ edges, int n_edges,
                                                const BinGuess& G) {
    const int mode = MODE < 0 ? G.mode : MODE;
    if (mode == 0) return bin_index(x, edges, n_edges);
    if (!(x >= edges[0]) || !(x <= edges[n_edges - 1])) return -1;
    // mode 3: a zero bin [edges[0], edges[1]) in front of log-uniform edges (np.insert(np.logspace(...), 0, 1e-100):
    // vals_to_pdf, utils/hist_utils.py:128-131; the shower histograms); the guess describes edges[1:]
    const int off = mode == 3 ? 1 : 0;
    if (off) {
        if (x < edges[1]) return 0;
        edges += 1;
        n_edges -= 1;
    }
    const float xf = (float)x;
    const float g = mode == 1 ? xf : __log2f(xf);
    const float f = (g - G.b0) * G.inv_d;
    const float fl = floorf(f);
    int idx = (int)fl;
    // Away from the edges the FP32 map cannot be off by a bin: no look at the stored edges at all (two shared-memory
    // loads with scattered addresses less per sample).  Within eps of an edge the exact walk below decides.
    if (G.eps > 0.0f && f - fl > G.eps && f - fl < 1.0f - G.eps && idx >= 0 && idx <= n_edges - 2) return idx + off;
    idx = max(0, min(idx, n_edges - 2));
    while (idx > 0 && x < edges[idx]) --idx;
    while (idx < n_edges - 2 && x >= edges[idx + 1]) ++idx;
    return idx + off;
}

}  // namespace jmc
