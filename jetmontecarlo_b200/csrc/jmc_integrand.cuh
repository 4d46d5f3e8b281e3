// FP64 (NumPy-order) evaluation of one sample of an integrand: samples ->
// jacobian, observable, weight*jacobian.  Shared by the replay kernel
// (samples read from device arrays) and the exact mode of the fused kernel
// (samples drawn from Philox in registers).
#pragma once
#include "jmc_common.cuh"
#include "jmc_philox.cuh"
#include "jmc_physics.cuh"

namespace jmc {

struct SampleSet {
    double z_c, th_c;   // critical emission
    double z_s, th_s;   // subsequent / ungroomed emission
    double z_pre;       // pre-critical emission (or the simple sampler's x)
    double zb_u;        // zero-bin uniform
};

// ---- the sampler formulas, from uniforms ----------------------------------
// ref: numerics/radiators/samplers.py:38-62 (critical; z is drawn first)
__device__ __forceinline__ void sample_critical(const RunParams& p, double u0, double u1, double& z, double& th) {
    if (p.g.method == JMC_LIN) {
        z = lin_sample(p.g.z_cut, 0.5, u0);
        th = lin_sample(0.0, p.g.radius, u1);
    } else {
        z = log_sample(p.g.z_cut, p.log_eps, p.lw_crit_z, u0);
        th = log_sample(0.0, p.log_eps, p.lw_radius, u1);
    }
}
// ref: numerics/radiators/samplers.py:204-226 (ungroomed)
__device__ __forceinline__ void sample_ungroomed(const RunParams& p, double u0, double u1, double& z, double& th) {
    if (p.g.method == JMC_LIN) {
        z = lin_sample(0.0, 0.5, u0);
        th = lin_sample(0.0, p.g.radius, u1);
    } else {
        z = log_sample(0.0, p.log_eps, p.lw_half, u0);
        th = log_sample(0.0, p.log_eps, p.lw_radius, u1);
    }
}
// ref: numerics/radiators/samplers.py:140-158 (precritical)
__device__ __forceinline__ double sample_precritical(const RunParams& p, double u) {
    return p.g.method == JMC_LIN ? lin_sample(0.0, p.g.z_cut, u) : log_sample(0.0, p.log_eps, p.lw_pre, u);
}
// ref: montecarlo/sampler.py:124-132 (simple)
__device__ __forceinline__ double sample_simple(const RunParams& p, double u) {
    return p.g.method == JMC_LIN ? lin_sample(p.g.lo, p.g.hi, u) : log_sample(p.g.lo, p.log_eps, p.lw_simple, u);
}

// Draw the samples of integrand KIND for global sample index `idx`.
// Uniform slots: 0,1 critical (z, theta) | 2,3 subsequent (z, theta) | 4 pre | 5 zero-bin;
// kinds with a single sampler use slots 0,1.
template <int KIND>
__device__ __forceinline__ void draw_exact(const RunParams& p, uint64_t idx, uint64_t seed, SampleSet& s) {
    double u0, u1;
    philox_uniform53x2(idx, 0u, seed, u0, u1);
    if (KIND == JMC_KIND_RADIATOR) {
        sample_ungroomed(p, u0, u1, s.z_s, s.th_s);
    } else if (KIND == JMC_KIND_SIMPLE) {
        s.z_pre = sample_simple(p, u0);
    } else {
        sample_critical(p, u0, u1, s.z_c, s.th_c);
    }
    if (KIND == JMC_KIND_CRIT_SUB || KIND == JMC_KIND_PRE_CRIT_SUB) {
        philox_uniform53x2(idx, 1u, seed, u0, u1);
        sample_ungroomed(p, u0, u1, s.z_s, s.th_s);
    }
    if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        philox_uniform53x2(idx, 2u, seed, u0, u1);
        s.z_pre = sample_precritical(p, u0);
        s.zb_u = u1;
    }
}

// ---- "the reference's weight is certainly NaN" ------------------------------
// The running-coupling radiators evaluate every region and mask afterwards, so a single log of a negative
// number anywhere makes the weight NaN.  The tests below recompute a few of those log ARGUMENTS exactly as the
// full evaluation does (same FP64 expressions) and report NaN only when one of them is negative -- so skipping
// the full evaluation changes nothing in the result; it only saves the ~60 transcendental calls for the 65 %
// (critical) to 99.7 % (critical x subsequent) of samples below the Landau scale.
// ref: analytics/radiators/running_coupling.py:157-175 (critSC1), :31-56 (subSC2), :11-29 (subSC1), qcd_utils.py:223-236
__device__ __forceinline__ bool crit_rc_surely_nan(double theta, double z_c) {
    return 1.0 + lam(z_c * theta, JMC_ALPHA_FIXED) < 0.0;            // W(1 + Lambda(z_c theta)) in critSC1
}
__device__ __forceinline__ bool sub_rc_surely_nan(double csub, double beta, double max_radius) {
    const double C = csub / np_pow(max_radius, beta);
    const double alpha = alpha_one_loop(np_max(JMC_PT * max_radius, 1.0));
    if (1.0 + lam(JMC_MU_NP, alpha) < 0.0) return true;              // log(1 + Lambda(MU_NP)) in subSC2
    if (1.0 + lam(C, alpha) < 0.0) return true;                      // W(1 + Lambda(C)) in subSC1
    return 2.0 - 2.0 * C < 0.0;                                      // log(2 - 2C) in b_bar(C) of subHC1
}

// ---- the observable alone ---------------------------------------------------
// (the same expressions as in integrand_eval, for passes that need nothing else)
template <int KIND>
__device__ __forceinline__ double integrand_observable(const RunParams& p, const SampleSet& s) {
    const jmc_integrand& g = p.g;
    if (KIND == JMC_KIND_RADIATOR) return ecf_ungroomed(s.z_s, s.th_s * g.theta_scale, g.beta, g.obs_acc);
    if (KIND == JMC_KIND_CRIT) return ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, g.z_pre, g.f_soft, g.obs_acc);
    if (KIND == JMC_KIND_CRIT_SUB)
        return np_max(ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, g.z_pre, g.f_soft, g.obs_acc),
                      csub_verbatim(s.z_s, s.th_s, g.beta));
    if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        const double cr = casimir(g.jet);
        const double cdf_eps = exp(-2.0 * cr * JMC_ALPHA_FIXED / JMC_PI * log(JMC_R0 / s.th_c) * log(g.z_cut / g.epsilon));
        const double zp = (g.method == JMC_LOG && s.zb_u < cdf_eps) ? 0.0 : s.z_pre;
        return np_max(ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, zp, g.f_soft, g.obs_acc), s.z_s * np_pow(s.th_c, g.beta));
    }
    return s.z_pre;
}

// ---- running-coupling CRIT / CRIT_SUB in two steps ---------------------------
// light: observable and "is the weight certainly NaN"; heavy: the weight times the jacobians of a sample that
// passed -- the same expressions and operation order as integrand_eval, so that a kernel may evaluate the heavy step
// for a different lane than the one that drew the sample (run_exact_kernel's warp compaction).
template <int KIND>
__device__ __forceinline__ bool integrand_light_rc(const RunParams& p, const SampleSet& s, double& obs) {
    const jmc_integrand& g = p.g;
    if (KIND == JMC_KIND_CRIT) {
        obs = ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, g.z_pre, g.f_soft, g.obs_acc);
        return !crit_rc_surely_nan(s.th_c, g.z_cut);
    }
    const double csub = csub_verbatim(s.z_s, s.th_s, g.beta);
    obs = np_max(ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, g.z_pre, g.f_soft, g.obs_acc), csub);
    return !(crit_rc_surely_nan(s.th_c, g.z_cut) || sub_rc_surely_nan(csub, g.beta, s.th_c));
}
template <int KIND, bool ACTIVE>
__device__ __forceinline__ double integrand_heavy_rc(const RunParams& p, double z_c, double th_c, double z_s, double th_s,
                                                     const CritConsts* cc) {
    const jmc_integrand& g = p.g;
    const bool is_log = g.method == JMC_LOG;
    if (KIND == JMC_KIND_CRIT) {
        double w = weight_critical<ACTIVE>(z_c, th_c, g.z_cut, g.jet, 0, cc);
        if (g.nan_to_num) w = nan_to_num(w);
        return w * (is_log ? (z_c - g.z_cut) * th_c : 1.0);
    }
    double w = weight_two_emission<ACTIVE>(z_c, th_c, z_s, th_s, g.z_cut, g.beta, g.jet, 0, cc);
    if (g.nan_to_num) w = nan_to_num(w);
    const double jc = is_log ? (z_c - g.z_cut) * th_c : 1.0;
    const double js = is_log ? z_s * th_s : 1.0;
    return w * jc * js;
}

// ---- one sample of the integrand -------------------------------------------
// jac: product of the sampler jacobians (incl. theta_scale for log RADIATOR);
// obs: the binned observable; wj: weight * jac, nan_to_num applied to the
// weight first when g.nan_to_num is set.
template <int KIND, bool ACTIVE = true>
__device__ __forceinline__ void integrand_eval(const RunParams& p, const SampleSet& s, double& jac, double& obs,
                                               double& wj, const CritConsts* cc = nullptr, bool need_w = true) {
    const jmc_integrand& g = p.g;
    const bool is_log = g.method == JMC_LOG;
    if (!need_w) {          // a pass that only looks at the observables (setBins' min/max): no weight, no jacobian
        jac = wj = 0.0;
        obs = integrand_observable<KIND>(p, s);
        return;
    }
    if (KIND == JMC_KIND_RADIATOR) {
        // ref: numerics/radiators/generation.py:303-335 (theta_scale = theta_crit), weights.py:15-26
        const double th_em = s.th_s * g.theta_scale;
        obs = ecf_ungroomed(s.z_s, th_em, g.beta, g.obs_acc);
        double w = weight_radiator(s.z_s, th_em, g.jet, g.fixed_coupling, g.split_acc);
        if (g.nan_to_num) w = nan_to_num(w);
        jac = is_log ? (s.z_s * s.th_s) * g.theta_scale : 1.0;
        wj = w * jac;
    } else if (KIND == JMC_KIND_CRIT) {
        // ref: examples/tests/oneEmission_tests/test_fixedcoupcritsudakov.py:124-220
        obs = ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, g.z_pre, g.f_soft, g.obs_acc);
        double w = (!g.fixed_coupling && crit_rc_surely_nan(s.th_c, g.z_cut))
                       ? qnan() : weight_critical<ACTIVE>(s.z_c, s.th_c, g.z_cut, g.jet, g.fixed_coupling, cc);
        if (g.nan_to_num) w = nan_to_num(w);
        jac = is_log ? (s.z_c - g.z_cut) * s.th_c : 1.0;
        wj = w * jac;
    } else if (KIND == JMC_KIND_CRIT_SUB) {
        // ref: numerics/weights.py:63-99; examples/sudakov_comparisons/sudakov_utils.py:525-538
        const double csub = csub_verbatim(s.z_s, s.th_s, g.beta);
        const double ccrit = ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, g.z_pre, g.f_soft, g.obs_acc);
        obs = np_max(ccrit, csub);
        double w = (!g.fixed_coupling && (crit_rc_surely_nan(s.th_c, g.z_cut) || sub_rc_surely_nan(csub, g.beta, s.th_c)))
                       ? qnan()
                       : weight_two_emission<ACTIVE>(s.z_c, s.th_c, s.z_s, s.th_s, g.z_cut, g.beta, g.jet, g.fixed_coupling, cc);
        if (g.nan_to_num) w = nan_to_num(w);
        const double jc = is_log ? (s.z_c - g.z_cut) * s.th_c : 1.0;
        const double js = is_log ? s.z_s * s.th_s : 1.0;
        jac = jc * js;
        wj = w * jc * js;
    } else if (KIND == JMC_KIND_PRE_CRIT_SUB) {
        // ref: examples/tests/comparison_tests/comparison_utils_probabilistic.py:384-474
        const double cr = casimir(g.jet);
        const double cdf_eps = exp(-2.0 * cr * JMC_ALPHA_FIXED / JMC_PI * log(JMC_R0 / s.th_c)
                                   * log(g.z_cut / g.epsilon));
        const double zp = (is_log && s.zb_u < cdf_eps) ? 0.0 : s.z_pre;
        const double thb = np_pow(s.th_c, g.beta);
        const double c_sub = s.z_s * thb;
        obs = np_max(ecf_groomed(s.z_c, s.th_c, g.z_cut, g.beta, zp, g.f_soft, g.obs_acc), c_sub);
        double w = weight_critical(s.z_c, s.th_c, g.z_cut, g.jet, 1)
                   * pdf_sub_fc(c_sub / thb, g.beta, g.jet, 1.0)
                   * weight_precritical_zerobin(zp, s.th_c, g.z_cut, g.jet, g.epsilon);
        if (g.nan_to_num) w = nan_to_num(w);
        const double jc = is_log ? (s.z_c - g.z_cut) * s.th_c : 1.0;
        const double js = is_log ? s.z_s * s.th_s : 1.0;
        const double jp = is_log ? s.z_pre : 1.0;
        jac = (zp == 0.0) ? jc * js : jc * jp * js;
        wj = w * jac;
    } else {  // JMC_KIND_SIMPLE
        obs = s.z_pre;
        jac = is_log ? s.z_pre : 1.0;
        wj = jac;
    }
}

}  // namespace jmc
