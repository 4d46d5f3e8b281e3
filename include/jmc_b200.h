/* jmc_b200.h -- C ABI of libjmc_b200.so: the B200 (sm_100a) implementation of the
 * JetMonteCarlo Monte Carlo integration hot path
 *
 *     sample emission phase space -> jacobian x weight -> weighted histogram
 *     -> density +- error -> cumulative integral +- error.
 *
 * The reference (samcaf/JetMonteCarlo) is pure Python/NumPy and has NO FFI or
 * plugin layer on this path; every entry point below therefore cites the Python
 * interface it replaces (paths relative to the reference's jetmontecarlo/
 * package) and INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary;
 *  - pointers named d_* are DEVICE pointers owned by the caller, h_* are HOST
 *    pointers; `stream` is a cudaStream_t passed as void* (NULL = legacy
 *    default stream).  Device entry points are asynchronous on `stream`;
 *    host entry points (suffix _host) synchronise before returning;
 *  - every function returns 0 on success or a negative JMC_E* code; the
 *    message of the last failure on the calling thread is jmc_last_error();
 *  - histogram outputs ACCUMULATE (+=) into the caller's buffers so that a job
 *    can be split over several launches; the caller zeroes them first.
 *  - all arithmetic of the replay path is IEEE FP64 in NumPy operation order.
 */
#ifndef JMC_B200_H
#define JMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMC_ABI_VERSION 2

/* error codes */
#define JMC_OK 0
#define JMC_EINVAL (-1)   /* bad argument (message says which)            */
#define JMC_ECUDA (-2)    /* CUDA runtime error (message has the string)  */
#define JMC_ENODEV (-3)   /* no usable sm_100 device                      */
#define JMC_ELIMIT (-4)   /* size beyond a documented limit               */
#define JMC_EDOMAIN (-5)  /* reference would raise (e.g. veto prob > 1)   */

/* enums (kept as int32 in the structs) */
enum { JMC_LIN = 0, JMC_LOG = 1 };                         /* sampleMethod 'lin' / 'log'        */
enum { JMC_QUARK = 0, JMC_GLUON = 1 };                     /* jet_type                          */
enum { JMC_ACC_LL = 0, JMC_ACC_LL_NONSING = 1, JMC_ACC_MLL = 2 };
enum { JMC_BC_FIRST = 0, JMC_BC_LAST_PLUS = 1, JMC_BC_LAST_MINUS = 2 };
enum { JMC_F64 = 0, JMC_F32 = 1 };                         /* arithmetic of the fused path      */
/* What d_count of the fused path holds.  JMC_COUNT_ALL: np.histogram's count (every sample inside the binned range),
 * as integrator.setDensity computes it (montecarlo/integrator.py:101).  JMC_COUNT_WEIGHTED: only the samples whose
 * weight is non-zero.  The reference uses the count solely to zero the density of empty bins (:123-124), and a bin
 * without a weighted sample has sum w = 0, so density, error and integral are identical; the running-coupling
 * integrands (weight NaN -> 0 below the Landau scale for up to 99.7 % of the samples) then skip observable, bin and
 * count for those samples.  Honoured by JMC_F32 runs with nan_to_num set; everything else counts all. */
enum { JMC_COUNT_ALL = 0, JMC_COUNT_WEIGHTED = 1 };
/* recursive safe subtraction (getECFs_rss, utils/partonshower_utils.py:501-674) */
enum { JMC_RSS_PRECRIT = 1, JMC_RSS_CRIT = 2, JMC_RSS_SUB = 4, JMC_RSS_FEW_PRES = 8 };

/* Samplers.  ref: montecarlo/sampler.py:114-140 (simple),
 * numerics/radiators/samplers.py:23-119 (critical), :125-180 (precritical),
 * :189-302 (ungroomed). */
enum { JMC_SAMPLER_SIMPLE = 0, JMC_SAMPLER_CRITICAL = 1,
       JMC_SAMPLER_PRECRITICAL = 2, JMC_SAMPLER_UNGROOMED = 3 };

typedef struct jmc_sampler {
    int32_t kind;        /* JMC_SAMPLER_*                                  */
    int32_t method;      /* JMC_LIN / JMC_LOG                              */
    double  epsilon;     /* log-sampling cutoff                            */
    double  zc;          /* critical / precritical: z_cut                  */
    double  radius;      /* critical / ungroomed: jet radius               */
    double  lo, hi;      /* simple: bounds                                 */
} jmc_sampler;

/* Array functions of the replay path (jmc_eval_fn).  Arguments a0..a3 are
 * per-sample arrays (stride in doubles) or NULL = use the scalar in
 * jmc_fn_params.  ref lines are in the table of DESIGN.md and next to each
 * device function in csrc/jmc_physics.cuh. */
enum {
    JMC_FN_ALPHA_S = 0,          /* a0=z a1=theta            qcd_utils.py:115-119              */
    JMC_FN_SPLITTING = 1,        /* a0=z                     qcd_utils.py:160-190              */
    JMC_FN_RADIATOR_WEIGHT = 2,  /* a0=z a1=theta            numerics/weights.py:15-26         */
    JMC_FN_CRIT_PDF_FC = 3,      /* a0=z a1=theta            radiators/fixedcoupling.py:59-72  */
    JMC_FN_CRIT_RAD_FC = 4,      /* a0=theta                 radiators/fixedcoupling.py:7-19   */
    JMC_FN_SUB_RAD_FC = 5,       /* a0=C [a1=maxRadius]      radiators/fixedcoupling.py:77-90  */
    JMC_FN_SUB_RADPRIME_FC = 6,  /* a0=C [a1=maxRadius]      radiators/fixedcoupling.py:92-108 */
    JMC_FN_SUB_PDF_FC = 7,       /* a0=C [a1=maxRadius]      radiators/fixedcoupling.py:162-176*/
    JMC_FN_PRE_RAD_FC = 8,       /* a0=z_pre a1=theta_crit a2=maxRadius  radiators/fixedcoupling.py:181-192*/
    JMC_FN_CRIT_RAD_RC = 9,      /* a0=theta                 radiators/running_coupling.py:263-289 */
    JMC_FN_CRIT_RADPRIME_RC = 10,/* a0=theta                 radiators/running_coupling.py:326-347 */
    JMC_FN_CRIT_PDF_RC = 11,     /* a0=z a1=theta            radiators/running_coupling.py:386-404 */
    JMC_FN_SUB_RAD_RC = 12,      /* a0=C [a1=maxRadius]      radiators/running_coupling.py:119-148 */
    JMC_FN_SUB_RADPRIME_RC = 13, /* a0=C [a1=maxRadius]      radiators/running_coupling.py:350-380 */
    JMC_FN_SUB_PDF_RC = 14,      /* a0=C [a1=maxRadius]      radiators/running_coupling.py:409-422 */
    JMC_FN_PRE_WEIGHT = 15,      /* a0=z_pre a1=theta_crit   numerics/weights.py:43-58         */
    JMC_FN_C_GROOMED = 16,       /* a0=z a1=theta [a2=z_pre] [a3=f]  numerics/observables.py:15-79 */
    JMC_FN_C_UNGROOMED = 17,     /* a0=z a1=theta            numerics/observables.py:82-109    */
    JMC_FN_TWO_EMISSION = 18,    /* a0=z_c a1=th_c a2=z_s a3=th_s    numerics/weights.py:63-99 */
    JMC_FN_PRE_WEIGHT_ZEROBIN = 19, /* a0=z_pre a1=theta_crit  examples/tests/comparison_tests/
                                       comparison_utils_probabilistic.py:82-98 */
    JMC_FN_ALPHA_SPLITTING = 20, /* a0=z a1=theta: alpha * splittingFn(z)   numerics/splitting.py:34-39,52-58 */
    JMC_FN_MUL = 21,             /* a0 * a1 (weights * jacobians, e.g. montecarlo/integrator.py:447)        */
    JMC_FN_MASK_FINITE = 22,     /* a0=w a1=v: w if both finite else 0 (make_finite of utils/hist_utils.py:103-108) */
    JMC_FN_EWOC_WEIGHT = 23,     /* a0=z1 a1=z2 a2=w, s3=w1, beta=w2: z1**w1 * z2**w2 * w   montecarlo/ewoc.py:23-27 */
    /* the pieces of the running-coupling radiators (soft-collinear regions 1-3, hard-collinear 1-2), a0 = C or theta,
     * a1 = alpha (NULL: s1); the reference's analytic MLL curves are assembled from them
     * (analytics/radiators/running_coupling.py:11-117 and :157-258) */
    JMC_FN_SUB_SC1 = 24, JMC_FN_SUB_SC2 = 25, JMC_FN_SUB_SC3 = 26, JMC_FN_SUB_HC1 = 27, JMC_FN_SUB_HC2 = 28,
    JMC_FN_CRIT_SC1 = 29, JMC_FN_CRIT_SC2 = 30, JMC_FN_CRIT_SC3 = 31, JMC_FN_CRIT_HC1 = 32, JMC_FN_CRIT_HC2 = 33,
    JMC_FN_COUNT = 34
};

typedef struct jmc_fn_params {
    int32_t jet;             /* JMC_QUARK / JMC_GLUON                          */
    int32_t acc;             /* JMC_ACC_* (splitting / observable accuracy)    */
    int32_t fixed_coupling;  /* 1 = alpha_fixed, 0 = running                   */
    int32_t reserved;
    double  z_cut;           /* z_c / z_cut                                    */
    double  beta;            /* angular exponent                               */
    double  max_radius;      /* maxRadius scalar (when a1 == NULL)             */
    double  z_pre;           /* C_groomed: scalar z_pre (when a2 == NULL)      */
    double  f_soft;          /* C_groomed: scalar f (when a3 == NULL)          */
    double  epsilon;         /* zero-bin weight cutoff                         */
    double  s0, s1, s2, s3;  /* scalar stand-ins for NULL a0..a3               */
} jmc_fn_params;

/* Integrands = (samplers, observable, weight) combinations of the hot path. */
enum {
    /* ungroomedSampler; obs C_ungroomed(z, theta*theta_scale, beta, obs_acc);
     * w = radiatorWeight(z, theta*theta_scale, jet, fc, split_acc) * jac
     *     [* theta_scale for log sampling].
     * ref: examples/tests/radiator_tests/test_runningcoupsubradiators.py:125-201,
     *      numerics/radiators/generation.py:303-335 (theta_scale = theta_crit). */
    JMC_KIND_RADIATOR = 0,
    /* criticalSampler; obs C_groomed(z, theta, z_cut, beta, z_pre, f, obs_acc);
     * w = criticalEmissionWeight * jac.
     * ref: examples/tests/oneEmission_tests/test_fixedcoupcritsudakov.py:124-220. */
    JMC_KIND_CRIT = 1,
    /* criticalSampler x ungroomedSampler; csub = z_s th_s^b th_s^b;
     * obs = max(C_groomed(z_c, th_c, z_cut, beta, z_pre, f, obs_acc), csub);
     * w = twoEmissionWeight * jac_c * jac_s.
     * ref: numerics/weights.py:63-99,
     *      examples/sudakov_comparisons/sudakov_utils.py:481-486,525-538. */
    JMC_KIND_CRIT_SUB = 2,
    /* critical x ungroomed x precritical with the zero-bin draw (fc only).
     * ref: examples/tests/comparison_tests/comparison_utils_probabilistic.py:384-474. */
    JMC_KIND_PRE_CRIT_SUB = 3,
    /* simpleSampler; obs = sample; w = jac (the caller's weight function runs
     * on the host).  ref: montecarlo/integrator.py:390-457. */
    JMC_KIND_SIMPLE = 4,
    /* Angularity shower with the MLL veto algorithm; one "sample" = one jet,
     * weight 1.  obs = ungroomed, Soft Drop or RSS-groomed ECF of the emission chain.
     * ref: utils/partonshower_utils.py:130-328, 445-499, 501-674, 676-790. */
    JMC_KIND_SHOWER = 5,
    JMC_KIND_COUNT = 6
};

typedef struct jmc_integrand {
    int32_t kind;            /* JMC_KIND_*                                     */
    int32_t method;          /* JMC_LIN / JMC_LOG (all samplers of the kind)   */
    int32_t jet;             /* JMC_QUARK / JMC_GLUON                          */
    int32_t fixed_coupling;  /* 1 / 0                                          */
    int32_t obs_acc;         /* JMC_ACC_LL / JMC_ACC_MLL                       */
    int32_t split_acc;       /* JMC_ACC_* (radiatorWeight; shower: LL or MLL)  */
    int32_t nan_to_num;      /* 1: weights pass through np.nan_to_num          */
    int32_t shower_groomer;  /* SHOWER: 0 ungroomed, 1 Soft Drop, 2 recursive safe subtraction (f_soft, z_cut) */
    int32_t shower_n_emissions; /* SHOWER: 1, 2, ... or 0 = 'all'              */
    int32_t counts;          /* JMC_COUNT_ALL / JMC_COUNT_WEIGHTED (fused FP32 path) */
    double  epsilon;         /* log-sampling cutoff                            */
    double  z_cut;           /* grooming parameter                             */
    double  radius;          /* jet radius                                     */
    double  beta;            /* angular exponent                               */
    double  f_soft;          /* C_groomed f                                    */
    double  z_pre;           /* C_groomed scalar z_pre (kinds without a pre sampler) */
    double  theta_scale;     /* RADIATOR: theta_crit rescaling (1 = none)      */
    double  lo, hi;          /* SIMPLE: bounds                                 */
    double  beta_sd;         /* SHOWER: Soft Drop angular exponent             */
    double  shower_cutoff;   /* SHOWER: continue while z*theta > cutoff        */
    int32_t shower_rss_flags;/* SHOWER, groomer 2: JMC_RSS_* (emission types kept: the reference matches its
                                emission_type string by substring, partonshower_utils.py:588-640; few_pres) */
    int32_t reserved;
} jmc_integrand;

/* ---- library ---------------------------------------------------------- */
int         jmc_abi_version(void);
const char* jmc_last_error(void);
/* Number of CUDA devices visible; selects `device` for the calling thread. */
int         jmc_device_count(void);
int         jmc_set_device(int device);
/* 148 on B200; used to size persistent grids. */
int         jmc_sm_count(void);
/* phase-space area of a sampler (setArea).  ref: samplers.py:30-36,132-138,196-202;
 * sampler.py:118-122 */
int         jmc_sampler_area(const jmc_sampler* s, double* area);
/* number of uniforms one sample of the integrand consumes */
int         jmc_integrand_uniforms(const jmc_integrand* g);
/* product of the sampler areas of the integrand */
int         jmc_integrand_area(const jmc_integrand* g, double* area);

/* ---- sampling on request (replaces sampler.generateSamples,
 *      montecarlo/sampler.py:30-34 + the addSample of each sampler) -------- */
/* d_samples: [n] (1-D samplers) or [n,2] row-major (z, theta); d_jac: [n].
 * Sample i uses Philox-4x32-10 counter (first_index+i, slot), key = seed. */
int jmc_sample(const jmc_sampler* s, uint64_t n, uint64_t seed, uint64_t first_index,
               double* d_samples, double* d_jac, void* stream);
/* raw 53-bit uniforms [n, n_uniform] of the same stream (tests, zero-bin draw) */
int jmc_uniforms(uint64_t n, uint64_t seed, uint64_t first_index, int n_uniform,
                 double* d_out, void* stream);

/* Inverse-transform sampling from a tabulated cdf: out[i] = interp1d(cdf_vals, x_vals, fill_value='extrapolate')(u_i)
 * with u_i the first uniform of stream index first_index + i.  The table (n_table >= 2 points) must be sorted by
 * cdf value (scipy's interp1d sorts it; the host wrapper does the same).
 * ref: utils/montecarlo_utils.py:377-399 (inverse_transform_samples) */
int jmc_inverse_transform(const double* d_cdf_vals, const double* d_x_vals, int n_table, uint64_t n, uint64_t seed,
                          uint64_t first_index, double* d_out, void* stream);

/* Per-event inverse-transform sampling from a CONDITIONAL cdf exp(-R(x, cond_i)): one sample per event i from the
 * cdf of x on [d_dom_lo[i], d_dom_hi[i]] given cond_i, with the uniform d_uniforms[i].  R is a numerical radiator
 * tabulated on the rectangular grid (d_grid_x [nx], d_grid_y [ny], d_table [nx, ny] row-major) and interpolated as
 * get_2d_interpolation does (utils/interpolation.py:248-273): interpolation 0 = 'RectangularGrid' (bilinear,
 * linearly extrapolated), 1 = 'Nearest' grid node.  Per event this is samples_from_cdf(cdf_i, 1, domain_i,
 * catch_turnpoint, force_monotone) on its tabulated route (utils/montecarlo_utils.py:142-318: probe grid, zero bin,
 * overflow bin, negligible head, turning point, forced monotonicity, then interp1d inversion :377-399); one event per
 * thread block.  Synchronises: an event whose cdf none of the options resolves is reported as JMC_EDOMAIN.
 * ref: examples/sudakov_comparisons/sudakov_utils.py:134-238 (generate_precritical / generate_subsequent_samples) */
int jmc_conditional_inverse_transform(const double* d_grid_x, int nx, const double* d_grid_y, int ny,
                                      const double* d_table, int interpolation, const double* d_cond,
                                      const double* d_dom_lo, const double* d_dom_hi, const double* d_uniforms,
                                      uint64_t n, int catch_turnpoint, int force_monotone, double* d_samples,
                                      void* stream);

/* generate_precritical_samples / generate_subsequent_samples of the production workflow, complete in one kernel
 * (examples/sudakov_comparisons/sudakov_utils.py:134-182, 185-238): one sample per critical angle from the forced-
 * monotone cdf exp(-R(x, theta_crit_i)) of the tabulated radiator, on [0, param] (JMC_SUDAKOV_PRECRITICAL, param =
 * z_cut) or on [0, theta_crit_i^param / 2] (JMC_SUDAKOV_SUBSEQUENT, param = beta; events whose upper end is below
 * 1e-10 take the underflow value 1e-100).  Uniform of event i: d_uniforms[i], or (d_uniforms NULL) the first uniform
 * of Philox stream (seed, i) -- what jmc_uniforms(n, seed, 0, 1, ...) returns.  An infinite sample is returned as 0
 * with weight 0 (:117-120), every other weight is 1. */
enum { JMC_SUDAKOV_PRECRITICAL = 1, JMC_SUDAKOV_SUBSEQUENT = 2 };
int jmc_sudakov_generate(int emission, const double* d_grid_x, int nx, const double* d_grid_y, int ny,
                         const double* d_table, int interpolation, const double* d_theta_crit, double param,
                         const double* d_uniforms, uint64_t seed, uint64_t n, double* d_samples, double* d_weights,
                         void* stream);

/* ---- replay path: arrays in, arrays out -------------------------------- */
int jmc_eval_fn(int fn, const jmc_fn_params* p, uint64_t n,
                const double* d_a0, int64_t stride0, const double* d_a1, int64_t stride1,
                const double* d_a2, int64_t stride2, const double* d_a3, int64_t stride3,
                double* d_out, void* stream);

/* Whole-integrand replay on given samples: jac, obs, weight (= pdf weight *
 * jac, nan_to_num applied if set) and bin index (-1 = dropped) per sample.
 * d_crit/d_sub: [n,2]; d_pre: [n]; d_zb_u: [n] zero-bin uniforms; unused ones
 * NULL.  For RADIATOR/SIMPLE the samples go in d_sub / d_pre respectively.
 * Any output may be NULL.  d_edges may be NULL when d_bin is NULL. */
int jmc_eval_integrand(const jmc_integrand* g, uint64_t n,
                       const double* d_crit, const double* d_sub, const double* d_pre,
                       const double* d_zb_u, const double* d_edges, int n_edges,
                       double* d_jac, double* d_obs, double* d_weight, int32_t* d_bin,
                       void* stream);

/* np.histogram(obs, edges) bin index per sample (-1 dropped).
 * ref: numpy.histogram as called at montecarlo/integrator.py:97-101 */
int jmc_bin_index(const double* d_obs, uint64_t n, const double* d_edges, int n_edges,
                  int32_t* d_bin, void* stream);

/* The three histograms of integrator.setDensity (montecarlo/integrator.py:97-101):
 * sum w, sum w^2, count per bin; n_edges-1 bins.  NaN weights poison their bin
 * and every higher bin, as NumPy's cumsum-based weighted histogram does.
 * d_weights may be NULL (count only). */
int jmc_hist(const double* d_obs, const double* d_weights, uint64_t n,
             const double* d_edges, int n_edges,
             double* d_sum_w, double* d_sum_w2, uint64_t* d_count, void* stream);

/* The three 2-D histograms of integrator_2d.setDensity (np.histogram2d, montecarlo/integrator.py:539-545):
 * row-major [nx_edges-1, ny_edges-1] sum w, sum w^2, count; half-open bins, last bin closed, outliers and NaN
 * dropped (histogramdd semantics: a NaN weight stays in its own bin).  d_weights may be NULL. */
int jmc_hist2d(const double* d_x, const double* d_y, const double* d_weights, uint64_t n,
               const double* d_edges_x, int nx_edges, const double* d_edges_y, int ny_edges,
               double* d_sum_w, double* d_sum_w2, uint64_t* d_count, void* stream);

/* integrator_2d: density, error (both stored transposed, [ny-1, nx-1], as the reference keeps them after
 * its .T at :577-578) and the double cumulative sum with boundary condition (multidim_cumsum, :463-467,
 * integrate :584-648).  d_integral / d_integral_err may both be NULL (density only). */
int jmc_finalize2d(const double* d_sum_w, const double* d_sum_w2, const uint64_t* d_count,
                   const double* d_edges_x, int nx_edges, const double* d_edges_y, int ny_edges,
                   double area, uint64_t n_total, int bc_kind, double bc_value,
                   double* d_density, double* d_density_err, double* d_integral, double* d_integral_err,
                   void* stream);

/* min / max of an array with NumPy NaN propagation -> d_minmax[0..1]
 * (integrator.setBins, montecarlo/integrator.py:42-46) */
int jmc_minmax(const double* d_x, uint64_t n, double* d_minmax, void* stream);

/* ---- fused path: generate -> weight -> bin, nothing per-sample in HBM ---- */
/* Samples first_index .. first_index+n-1 of stream `seed`.  Histograms
 * accumulate into d_sum_w / d_sum_w2 / d_count (n_edges-1 each).  If d_minmax
 * is not NULL, min and max of the observable are also merged into
 * d_minmax[0..1] (caller initialises to +inf, -inf); d_edges may then be NULL
 * for a min/max-only pass (data-dependent bins, integrator.setBins).
 * precision: JMC_F64 = NumPy-order FP64 (checkable sample by sample against
 * the oracle), JMC_F32 = log-space FP32 fast path (statistical parity). */
int jmc_run(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index,
            int precision, const double* d_edges, int n_edges,
            double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax,
            void* stream);

/* A grid of parameter points in ONE launch (BASELINE config 5; the theta_crit loop of gen_crit_sub_num_rad,
 * numerics/radiators/generation.py:303-359, and z_cut / beta sweeps).  g: n_points integrands sharing kind,
 * coupling and observable accuracy; every point sees the same sample stream (seed, first_index), as the reference
 * re-weights one stored sample set per grid point.  h_edges: HOST array [n_points, n_edges] (each point has its
 * own bins) or NULL for a min/max-only pass; d_sum_w, d_sum_w2, d_count: DEVICE [n_points, n_edges-1],
 * accumulated into; d_minmax: DEVICE [n_points, 2] or NULL (caller initialises to +inf, -inf). */
int jmc_run_batch(const jmc_integrand* g, int n_points, uint64_t n, uint64_t seed, uint64_t first_index,
                  int precision, const double* h_edges, int n_edges, double* d_sum_w, double* d_sum_w2,
                  uint64_t* d_count, double* d_minmax, void* stream);

/* jmc_finalize for n_jobs histograms at once: every array has a leading n_jobs dimension; d_area: [n_jobs]. */
int jmc_finalize_batch(int n_jobs, const double* d_sum_w, const double* d_sum_w2, const uint64_t* d_count,
                       const double* d_edges, int n_edges, const double* d_area, uint64_t n_total,
                       int bc_kind, double bc_value, double* d_density, double* d_density_err,
                       double* d_integral, double* d_integral_err, void* stream);

/* Samples on request: the per-sample view of the same streams jmc_run bins, written to HBM only when a
 * caller asks for it (sampler.getSamples() on a fused job; parity tests).  d_samples: [n,8] = z_crit,
 * theta_crit, z_sub, theta_sub, z_pre, zero-bin uniform, delta = z_crit - z_cut (in FP32 mode the sampled
 * quantity itself, free of the reference's cancellation), jacobian (FP64 mode); unused slots 0.  In FP32
 * mode z_pre is 0 for a zero-binned sample.  d_obs, d_weight (= weight * jacobian): [n].  Any output may
 * be NULL. */
int jmc_dump(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
             double* d_samples, double* d_obs, double* d_weight, void* stream);

/* Per-jet view of the shower streams: ECF observable, number of accepted emissions and number of veto trials of
 * jets first_index .. first_index+n-1 (outputs indexed 0..n-1; any may be NULL).
 * ref: utils/partonshower_utils.py:130-328 (gen_jets), :445-499, :676-790 (getECFs_*) */
int jmc_shower_dump(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                    double* d_obs, int32_t* d_n_emissions, int32_t* d_n_trials, void* stream);

/* The shower's event record on request: d_chains [n, max_emissions, 2] = (z, theta) of each accepted emission in
 * emission order (entries beyond a jet's length are left untouched), d_n_emissions [n] (may exceed max_emissions;
 * the record is then truncated).  With split_soft=False (partonshower_utils.py:321) this chain is the whole jet:
 * the reference's Parton tree and 3-vectors are functions of it.  ref: utils/partonshower_utils.py:224-328 */
int jmc_shower_chains(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                      int max_emissions, double* d_chains, int32_t* d_n_emissions, void* stream);

/* The ECF pickers on a stored event record: d_chains [n, max_emissions, 2] = (z, theta) in emission order, d_n_emissions
 * [n] (lengths; longer than max_emissions = truncated record) -> d_obs [n], with the groomer, accuracy, z_cut, f_soft,
 * n_emissions and RSS flags of `g`: ungroomed, Soft Drop or recursive safe subtraction, FP64 in the reference's
 * operation order (the same accumulator the shower kernel carries per jet).
 * ref: utils/partonshower_utils.py:445-499 (getECFs_ungroomed), :676-790 (getECFs_softdrop), :501-674 (getECFs_rss) */
int jmc_chain_ecf(const jmc_integrand* g, const int32_t* d_n_emissions, const double* d_chains, uint64_t n,
                  int max_emissions, double* d_obs, void* stream);

/* density, error, cumulative integral (+ error) from the histograms.
 * ref: montecarlo/integrator.py:113-124 (density), :129-202 (integrate).
 * density, density_err, integral_err: n_edges-1; integral: n_edges. */
int jmc_finalize(const double* d_sum_w, const double* d_sum_w2, const uint64_t* d_count,
                 const double* d_edges, int n_edges, double area, uint64_t n_total,
                 int bc_kind, double bc_value,
                 double* d_density, double* d_density_err,
                 double* d_integral, double* d_integral_err, void* stream);
/* d_integral / d_integral_err may both be NULL: density only (setDensity). */

/* integrator.integrate() alone: cumulative integral of a given density.
 * ref: montecarlo/integrator.py:129-202 */
int jmc_cumulative(const double* d_density, const double* d_density_err, const double* d_edges,
                   int n_edges, int bc_kind, double bc_value,
                   double* d_integral, double* d_integral_err, void* stream);

/* ---- host-buffer entry points (what a reference-side binding calls) ------ */
/* integrator.setDensity + integrate on host arrays: copies obs/weights to the
 * device, runs jmc_hist + jmc_finalize, copies the results back.  Any of the
 * h_sum_w / h_sum_w2 / h_count outputs may be NULL. */
int jmc_set_density_host(const double* h_obs, const double* h_weights, uint64_t n,
                         const double* h_edges, int n_edges, double area,
                         int bc_kind, double bc_value,
                         double* h_density, double* h_density_err,
                         double* h_integral, double* h_integral_err,
                         double* h_sum_w, double* h_sum_w2, uint64_t* h_count);

/* Whole job from a description: fused run of n samples + finalize.  n_total
 * is the normalisation count (= n on one GPU).  h_edges is read, everything
 * else is written. */
int jmc_integrate_host(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index,
                       int precision, const double* h_edges, int n_edges,
                       int bc_kind, double bc_value,
                       double* h_density, double* h_density_err,
                       double* h_integral, double* h_integral_err,
                       double* h_sum_w, double* h_sum_w2, uint64_t* h_count);

/* ---- generic Monte Carlo process + pairwise observables + energy-weighted correlators -------------------------
 * MonteCarloProcess (montecarlo/process.py:109-191): hit-or-miss sampling of a box of kinematic variables under a
 * phase-space constraint, weight = jacobian x differential cross section, area = box volume x acceptance.  The
 * constraint and the cross section are code, so a process is one of the kinds implemented on the device; the
 * reference has exactly one, e+ e- -> hadrons at LO (examples/processes/ee2hadrons_LO.py: variables 1 - x_1, 1 - x_2
 * in (0,1)^2, 0 < x_i < 1). */
enum { JMC_PROCESS_EE2HADRONS_LO = 0 };
/* pairwise observables of two partons with energy fractions (x_i, x_j), examples/processes/ee2hadrons_LO.py:95-150 */
enum { JMC_PAIR_COSTHETA = 0, JMC_PAIR_THETA = 1, JMC_PAIR_M2 = 2, JMC_PAIR_GECF = 3, JMC_PAIR_Z_I = 4, JMC_PAIR_Z_J = 5 };

typedef struct jmc_process {
    int32_t kind;            /* JMC_PROCESS_*                                   */
    int32_t method;          /* JMC_LIN / JMC_LOG                               */
    double  epsilon;         /* log-sampling cutoff                             */
    double  lo[4], hi[4];    /* kinematic_bounds (the first two are used)       */
    double  energy;          /* GeV                                             */
    double  prefactor;       /* of the differential cross section (analytics/ewocs/eec.py:9-19), from the host */
} jmc_process;

/* generateSamples: sample i is the first accepted trial of Philox stream first_index + i (trial t = block t).
 * d_samples [n,2] and d_weights [n] (= jacobian * differential cross section) may be NULL; *d_rejected (device
 * scalar) is incremented by the number of rejected trials (num_rejected_samples, process.py:142).  Synchronises
 * (a process whose box holds no accepted point is reported as JMC_EDOMAIN). */
int jmc_process_sample(const jmc_process* p, uint64_t n, uint64_t seed, uint64_t first_index,
                       double* d_samples, double* d_weights, uint64_t* d_rejected, void* stream);
/* ee2hLO_PairwiseObservable.update: d_obs [3n], pair-major ((1,2) of all events, then (2,3), then (3,1)). */
int jmc_process_pairs(const jmc_process* p, const double* d_samples, uint64_t n, int obs_id, double kappa,
                      double* d_obs, void* stream);
/* Fused Npoint_MonteCarloEWOC (montecarlo/ewoc.py:11-75): generate -> three pairs per event -> observable obs_id,
 * weight z_i^w1 z_j^w2 * event weight -> histograms over d_edges (accumulated; 3n entries), and / or min, max of the
 * observable in d_minmax (caller initialises to +inf, -inf; d_edges may then be NULL).  Nothing per event is stored. */
int jmc_process_ewoc(const jmc_process* p, uint64_t n, uint64_t seed, uint64_t first_index, int obs_id, double kappa,
                     double w1, double w2, const double* d_edges, int n_edges,
                     double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax,
                     uint64_t* d_rejected, void* stream);

/* ---- handle: cached job description, device workspace and pinned staging ----
 * The reference has no such object (it has no FFI); it exists so that a job issued from a host thread costs its
 * kernels and nothing else.  A handle belongs to the device that was current at jmc_create and to one host thread /
 * stream at a time.  It remembers the last (integrand, precision, edges) it saw together with everything derived
 * from them (FP32 edges in the kernel's compare domain, kernel parameters, both edge arrays on the device), so a
 * repeated or point-by-point job re-derives nothing, and it owns the device workspace and pinned staging of the
 * host-buffer entry point.  jmc_integrate_host above uses one implicit handle per (host thread, device). */
typedef struct jmc_handle jmc_handle;
int jmc_create(jmc_handle** out);
int jmc_destroy(jmc_handle* h);

/* jmc_run with HOST edges (what a caller of integrator.setBins has anyway): no read-back, no synchronisation,
 * asynchronous on `stream`.  Histograms accumulate into the caller's DEVICE buffers.  h_edges NULL + d_minmax: the
 * min/max-only pass. */
int jmc_handle_run(jmc_handle* h, const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index,
                   int precision, const double* h_edges, int n_edges,
                   double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, void* stream);

/* jmc_integrate_host on an explicit handle: at most one upload (edges, when they changed), the kernels, ONE
 * device-to-host copy of histograms + results, ONE synchronisation; no allocation after the first call. */
int jmc_handle_integrate_host(jmc_handle* h, const jmc_integrand* g, uint64_t n, uint64_t seed,
                              uint64_t first_index, int precision, const double* h_edges, int n_edges,
                              int bc_kind, double bc_value,
                              double* h_density, double* h_density_err,
                              double* h_integral, double* h_integral_err,
                              double* h_sum_w, double* h_sum_w2, uint64_t* h_count);

/* Device plumbing for hosts without an array library of their own: d_out[0..n) = value (FP64), and a byte fill
 * (cudaMemsetAsync).  Asynchronous on `stream`. */
int jmc_fill(double* d_out, uint64_t n, double value, void* stream);
int jmc_memset(void* d_ptr, int byte_value, uint64_t n_bytes, void* stream);
/* Buffers of the two collectives of a sharded job.  Histograms: d_buf [3, nb] FP64 = (sum w, sum w^2, count as FP64,
 * exact up to 2^53), unpack != 0 writes the reduced buffer back; one SUM all-reduce in between.  Min/max: d_buf
 * [n_pairs, 3] = (-min, max, has-NaN) with NaN -> -inf, so that ONE MAX all-reduce merges minima, maxima and the NaN
 * flags (a NaN on any rank makes the pair NaN, as np.min / np.max of the whole sample set would). */
int jmc_pack_histograms(double* d_sum_w, double* d_sum_w2, uint64_t* d_count, uint64_t nb, double* d_buf, int unpack,
                        void* stream);
int jmc_pack_minmax(double* d_minmax, uint64_t n_pairs, double* d_buf, int unpack, void* stream);
/* np.tile(d_in, reps) into d_out [reps * n] (the event weights of the three parton pairs, ewoc.py:11-57) */
int jmc_tile(const double* d_in, uint64_t n, int reps, double* d_out, void* stream);
/* number of finite entries (vals_to_pdf's "no finite values" test, utils/hist_utils.py:105-112); synchronises */
int jmc_count_finite(const double* d_in, uint64_t n, uint64_t* h_count, void* stream);

/* sampler.generateSamples into host arrays */
int jmc_sample_host(const jmc_sampler* s, uint64_t n, uint64_t seed, uint64_t first_index,
                    double* h_samples, double* h_jac);

/* array function on host arrays (same argument rules as jmc_eval_fn) */
int jmc_eval_fn_host(int fn, const jmc_fn_params* p, uint64_t n,
                     const double* h_a0, int64_t stride0, const double* h_a1, int64_t stride1,
                     const double* h_a2, int64_t stride2, const double* h_a3, int64_t stride3,
                     double* h_out);

/* number of kernels this library has launched on the calling process so far
 * (bench.py reports the difference over the timed region as gpu_launches) */
uint64_t jmc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JMC_B200_H */
