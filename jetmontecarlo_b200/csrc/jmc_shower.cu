// Angularity-ordered parton shower with the MLL veto (acceptance-rejection)
// algorithm, fused with the ECF observable and the histogram.
//
// What it replaces: gen_jets -> angularity_shower -> angularity_split
// (jetmontecarlo/utils/partonshower_utils.py:130-328) followed by
// getECFs_ungroomed (:445-499) or getECFs_softdrop (:676-790).  The reference
// builds a tree of Parton objects with 3-vectors per jet; with split_soft=False
// (:321) the geometry is redundant -- the opening angle of the daughters IS
// theta and the softer momentum fraction IS z -- so a jet is just its ordered
// chain of (z, theta), and the observables are running maxima / sums over
// that chain.  Nothing per jet is stored.
//
// Device design: one jet per lane, but lanes do not wait for each other: the
// kernels loop over veto TRIALS, and a lane whose chain ends starts its next
// jet (persistent lanes with refill).  Trial counts per jet are geometric
// (about 30 for quark, 57 for gluon jets at the 1 GeV cutoff), so lock-step
// jets would idle most lanes.
//
// Two kernels:
//   shower_kernel (FP64): the reference's operation sequence on 53-bit Philox
//         uniforms, jet by jet comparable with oracle/jmc_oracle.py:
//         shower_emissions;
//   shower_fast_kernel (FP32): log2(2 ang) carried instead of ang, one Philox
//         block (r1, r2, r3) per trial, MUFU transcendentals, division-free
//         veto, deferred observable work (see the kernel).
// compiled with --fmad=false (shared flag of the NumPy-order translation units)
#include <climits>
#include <cmath>
#include <type_traits>
#include <vector>
#include "jmc_common.cuh"
#include "jmc_philox.cuh"
#include "jmc_physics.cuh"

namespace jmc {

static constexpr int kShowerThreads = 128;
static constexpr int kMaxEmissions = 64;     // as the oracle; the reference recurses until the cutoff
static constexpr int kTopN = 4;              // n_emissions up to 4 (+ 'all')

// lanes of a warp that have finished their jet wait until this many are ready, then bin their observables and
// refill together (see shower_kernel)
#ifndef JMC_SHOWER_GROUP_EXACT
#define JMC_SHOWER_GROUP_EXACT 8
#endif

struct ShowerParams {
    int jet, mll_shower, obs_mll, groomer, n_emissions;
    int rss_pre, rss_crit, rss_sub, rss_few_pres;     // RSS: emission types kept, and the few-precriticals approximation
    double beta, radius, cutoff, z_cut, beta_sd;
    double f_soft;         // RSS: energy fraction groomed from the softer branch
    double alpha;          // alpha_fixed (LL) or the veto overestimate 1.5 (MLL)
    double kfac;           // pi beta / (C_R alpha)
    double cr;
    double ang_init;       // R^beta / 2
    float lf_init;         // ln(2 ang_init) = beta ln R
    BinGuess guess;        // affine first guess of the bin (exact walk decides); mode 0 = binary search
    int guess_offset;      // the guess describes edges[guess_offset:] (a leading zero-bin edge is skipped)
};

// sequential uniforms of one jet: block b of the jet's Philox stream yields uniforms 2b, 2b+1
struct UStream {
    uint64_t jet, seed;
    uint32_t blk;
    bool have;
    double buf;
    __device__ __forceinline__ void reset(uint64_t j, uint64_t s) { jet = j; seed = s; blk = 0u; have = false; }
    __device__ __forceinline__ double next() {
        if (have) { have = false; return buf; }
        double u0;
        philox_uniform53x2(jet, blk++, seed, u0, buf);
        have = true;
        return u0;
    }
};

// running top-N / sum of the chain's ECFs, in the order the reference adds them.  T = double in the FP64 mode
// (bit-for-bit the reference's sums); T = float in the FP32 mode, where the reference's 1e-50 place holders
// (below FP32 range) are carried as 0 and restored when the result is formed.
template <typename T>
struct EcfAcc {
    T top[kTopN];        // largest values seen, descending (ungroomed: all; soft drop: the ones after the first kept)
    T sum_all;           // running sum in chain order
    T first;             // soft drop: the critical emission's value; RSS: c_crit
    bool dropping;       // soft drop: still below the grooming condition; RSS: pre-critical emissions still allowed
    T this_zcut, max_zpre, sum_zpre, prod_scale;      // RSS state (partonshower_utils.py:501-674)
    static __device__ __forceinline__ T tiny() { return sizeof(T) == 8 ? (T)1e-50 : (T)0; }
    static __device__ __forceinline__ double wide(T v) { return (sizeof(T) == 4 && v == (T)0) ? 1e-50 : (double)v; }
    __device__ __forceinline__ void reset(int groomer, double z_cut = 0.0) {
#pragma unroll
        for (int k = 0; k < kTopN; ++k) top[k] = (T)-1;
        // getECFs_ungroomed starts its list with 1e-50 (:459); soft drop appends its two 1e-50 at the end;
        // RSS: c_subs = [1e-50], c_crit = 1e-50, z_pres = [1e-50] (:560-563)
        sum_all = groomer == 1 ? (T)0 : tiny();
        first = groomer == 2 ? tiny() : (T)-1;
        dropping = true;
        if (groomer != 1) insert(tiny());
        this_zcut = (T)z_cut;
        max_zpre = sum_zpre = tiny();
        prod_scale = (T)1;
    }
    // branch-free insertion into the descending top-N (a bubble of max/min pairs)
    __device__ __forceinline__ void insert(T c) {
#pragma unroll
        for (int k = 0; k < kTopN; ++k) {
            const T hi = c > top[k] ? c : top[k];
            c = c > top[k] ? top[k] : c;
            top[k] = hi;
        }
    }
    __device__ __forceinline__ void add(T c, bool passes_groomer, int groomer) {
        if (!groomer) {                        // ref: partonshower_utils.py:445-499
            insert(c);
            sum_all += c;
            return;
        }
        // ref: partonshower_utils.py:676-790: the first emission with z > z_cut theta^beta_sd is critical,
        // every later one is subsequent
        if (dropping) {
            if (!passes_groomer) return;
            dropping = false;
            first = c;
            sum_all = c;
            return;
        }
        insert(c);
        sum_all += c;
    }
    // Recursive safe subtraction: emissions softer than what is left of z_cut are pre-critical (they use z_cut up
    // and, for f != 1, rescale the hard branch), the first survivor is critical, later ones are subsequent.
    // thb = theta^beta.  Operation order of the reference (:575-640) -- in FP64 the result is the oracle's bit for bit.
    __device__ __forceinline__ void add_rss(T z, T thb, const ShowerParams& p) {
        const T f = (T)p.f_soft;
        T cutoff, scale;
        bool valid_pre;
        if (p.rss_few_pres) {
            cutoff = this_zcut - max_zpre;
            scale = (T)1 - ((T)1 - f) * max_zpre / f;
            valid_pre = z < this_zcut;
        } else {
            cutoff = this_zcut - sum_zpre;
            scale = prod_scale;
            valid_pre = z < cutoff;
        }
        if (p.rss_pre && dropping && valid_pre) {
            max_zpre = z > max_zpre ? z : max_zpre;
            sum_zpre += z;
            prod_scale *= (T)1 - ((T)1 - f) * z / f;
            return;
        }
        const bool is_crit = p.rss_crit && z >= cutoff && cutoff > (T)0;
        const bool is_sub = !is_crit && p.rss_sub && this_zcut == (T)0;
        if (!is_crit && !is_sub) return;
        const T zc = is_crit ? cutoff : (T)0;
        const T tail = p.obs_mll ? (T)1 - z - ((T)1 - f) * zc : (T)1 - ((T)1 - f) * zc;
        const T c = (z - f * zc) * thb * tail * (scale * scale);
        if (is_crit) {
            first = c;
            this_zcut = (T)0;
            dropping = false;
        } else {
            sum_all += c;
            insert(c);
        }
    }
    template <int G = -1>
    __device__ __forceinline__ double result(const ShowerParams& p) {
        const int groomer = G < 0 ? p.groomer : G;
        if (groomer == 2) {
            // c_list = c_subs + [c_crit]: the sum of all, or of the n largest (:655-668)
            if (p.n_emissions == 0) return (double)sum_all + wide(first) + (sizeof(T) == 4 ? 1e-50 : 0.0);
            insert(first);
            double s = 0.0;
            bool any = false;
#pragma unroll
            for (int k = 0; k < kTopN; ++k)
                if (k < p.n_emissions && top[k] >= (T)0) { s = any ? s + wide(top[k]) : wide(top[k]); any = true; }
            return s;
        }
        if (!groomer) {
            if (p.n_emissions == 0) return (double)sum_all + (sizeof(T) == 4 ? 1e-50 : 0.0);
            double s = 0.0;                    // sum(cs[:n]) over the descending list
#pragma unroll
            for (int k = 0; k < kTopN; ++k) if (k < p.n_emissions && top[k] >= (T)0) s += wide(top[k]);
            return s;
        }
        // c_list += [1e-50, 1e-50]
        if (dropping) {                        // nothing survived grooming: c_list = [1e-50, 1e-50]
            if (p.n_emissions == 0) return 1e-50 + 1e-50;
            if (p.n_emissions == 1) return 1e-50;
            double s = 1e-50;
            for (int k = 1; k < p.n_emissions && k < 2; ++k) s += 1e-50;
            return s;
        }
        if (p.n_emissions == 0) return ((double)sum_all + 1e-50) + 1e-50;
        if (p.n_emissions == 1) return (double)first;
        insert(tiny());
        insert(tiny());
        double s = (double)first;              // sum([c0] + rest[:n-1])
#pragma unroll
        for (int k = 0; k < kTopN; ++k) if (k < p.n_emissions - 1 && top[k] >= (T)0) s += wide(top[k]);
        return s;
    }
};

// One veto trial in the reference's FP64 operation order.  Returns true if the emission is accepted.
// ref: partonshower_utils.py:168-209
template <bool MLL, bool QUARK>
__device__ __forceinline__ bool trial_exact(const ShowerParams& p, UStream& us, double& ang, double& z, double& theta,
                                            int* error_flag) {
    const double r1 = us.next(), r2 = us.next();
    const double l = log(2.0 * ang);
    const double ang_f = exp(-sqrt(l * l - p.kfac * log(r1))) / 2.0;
    z = pow(2.0 * ang_f, r2) / 2.0;
    theta = pow(2.0 * ang_f, (1 - r2) / p.beta);
    ang = ang_f;
    if (!MLL) return true;
    const double cut = alpha_running(z, theta) * splitting(z, QUARK ? JMC_QUARK : JMC_GLUON, JMC_ACC_MLL)
                       / (2.0 * p.cr * p.alpha / z);
    if (cut > 1.0) *error_flag = 1;            // the reference raises ValueError here
    return us.next() < cut;
}

// warp-reduced (min, max) of the observables into the two global slots
__device__ __forceinline__ void block_minmax_to_global(double mn, double mx, double* __restrict__ d_minmax) {
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        unsigned long long* a0 = (unsigned long long*)&d_minmax[0];
        unsigned long long* a1 = (unsigned long long*)&d_minmax[1];
        unsigned long long old = *a0;
        while (mn < __longlong_as_double(old)) {
            const unsigned long long as = old;
            old = atomicCAS(a0, as, (unsigned long long)__double_as_longlong(mn));
            if (old == as) break;
        }
        old = *a1;
        while (mx > __longlong_as_double(old)) {
            const unsigned long long as = old;
            old = atomicCAS(a1, as, (unsigned long long)__double_as_longlong(mx));
            if (old == as) break;
        }
    }
}

// The FP64 shower: the reference's operation order, jet by jet comparable with the oracle.  One jet per lane with
// refill; launch-uniform choices are template parameters (shower accuracy, groomer, jet type, whether anything per
// jet is written out).
template <bool MLL, int GROOMER, bool QUARK, bool RECORD>
__global__ void __launch_bounds__(kShowerThreads)
shower_kernel(ShowerParams p, uint64_t n, uint64_t seed, uint64_t first, const double* __restrict__ edges,
              int n_edges, unsigned long long* __restrict__ g_count, double* __restrict__ d_minmax,
              double* __restrict__ dump_obs, int* __restrict__ dump_n, int* __restrict__ dump_trials,
              double* __restrict__ dump_chain, int chain_cap, int* error_flag) {
    extern __shared__ double s_raw[];
    const int nb = n_edges > 0 ? n_edges - 1 : 0;
    double* s_edges = s_raw;
    unsigned int* s_count = (unsigned int*)(s_edges + (n_edges > 0 ? n_edges : 0));
    for (int k = threadIdx.x; k < n_edges; k += blockDim.x) s_edges[k] = edges[k];
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s_count[k] = 0u;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mn = INFINITY, mx = -INFINITY;

    // per-lane jet state
    UStream us;
    EcfAcc<double> acc;
    double ang = 0.0, z = 0.0, theta = 0.0;
    int n_em = 0, n_trials = 0;
    bool in_veto_loop = false;
    bool alive = j < n;
    auto start_jet = [&]() {
        us.reset(first + j, seed);
        acc.reset(GROOMER, p.z_cut);
        ang = p.ang_init;                       // ang_init = R^beta / 2, z_init = 1/2, theta_init = R (:316-321)
        z = 0.5;
        theta = p.radius;
        n_em = 0;
        n_trials = 0;
        in_veto_loop = false;
    };
    if (alive) start_jet();

    while (true) {
        // does the chain continue?  (angularity_shower: recurse while z theta > cutoff, :262-279).  The cutoff is
        // tested on ACCEPTED emissions only: inside angularity_split the veto loop keeps trying without it (:188-209)
        bool finished = false;
        if (alive && !in_veto_loop) finished = !(z * theta > p.cutoff) || n_em >= kMaxEmissions;
        // Finishing a jet (observable, bin, histogram atomic, refill) is ~100 instructions that only the finished
        // lanes execute; with ~30 trials per jet some lane of a warp finishes in two iterations out of three, and
        // the other lanes would idle through it every time.  Finished lanes therefore wait until JMC_SHOWER_GROUP_EXACT
        // of them are ready (or until no lane of the warp has a trial left to do) and finish together.
        const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
        if (!alive_mask) break;
        const unsigned fin_mask = __ballot_sync(0xffffffffu, finished);
        const bool flush = __popc(fin_mask) >= JMC_SHOWER_GROUP_EXACT || fin_mask == alive_mask;
        if (!alive || (finished && !flush)) continue;
        if (!finished) {
            ++n_trials;
            const bool accepted = trial_exact<MLL, QUARK>(p, us, ang, z, theta, error_flag);
            in_veto_loop = !accepted;
            if (accepted) {
                if (RECORD && dump_chain && n_em < chain_cap) {        // event record on request: (z, theta) of emission n_em
                    double* c = dump_chain + ((size_t)j * chain_cap + n_em) * 2;
                    c[0] = z;
                    c[1] = theta;
                }
                ++n_em;
                if (n_em >= kMaxEmissions && error_flag) atomicOr(error_flag, 2);     // the record is cut here
                if (GROOMER == 2)
                    acc.add_rss(z, np_pow(theta, p.beta), p);
                else
                    acc.add(ecf_ungroomed(z, theta, p.beta, p.obs_mll ? JMC_ACC_MLL : JMC_ACC_LL),
                            z > p.z_cut * np_pow(theta, p.beta_sd), GROOMER);
            }
            continue;
        }
        // jet finished: observable -> histogram, then take the next jet
        const double obs = acc.template result<GROOMER>(p);
        if (RECORD && dump_obs) dump_obs[j] = obs;
        if (RECORD && dump_n) dump_n[j] = n_em;
        if (RECORD && dump_trials) dump_trials[j] = n_trials;
        if (n_edges > 0) {
            // the jets' bins: np.logspace edges, usually behind one zero-bin edge (plot_comparisons.py:296-298)
            int b;
            if (p.guess.mode == 0 || obs < s_edges[p.guess_offset]) {
                b = bin_index(obs, s_edges, n_edges);
            } else {
                b = bin_index_guided(obs, s_edges + p.guess_offset, n_edges - p.guess_offset, p.guess);
                if (b >= 0) b += p.guess_offset;
            }
            if (b >= 0) atomicAdd(&s_count[b], 1u);
        }
        mn = fmin(mn, obs);
        mx = fmax(mx, obs);
        j += stride;
        alive = j < n;
        if (alive) start_jet();
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
        if (s_count[k]) atomicAdd(&g_count[k], (unsigned long long)s_count[k]);
    if (d_minmax) block_minmax_to_global(mn, mx, d_minmax);
}

// ---- the FP32 shower ------------------------------------------------------------------------------------------------
// Same algorithm, organised around what the hardware charges for: issue slots of instructions that run with few
// lanes.  About one trial in ten is accepted and one in thirty ends a jet, so code that handles either "at once" runs
// after nearly every trial of a warp with 1-3 of 32 lanes.  Three stages are therefore decoupled:
//   * the PRODUCER of a lane draws veto trials -- one Philox block (r1, r2, r3) each -- and only carries
//     (Philox index of the jet, block counter, log2(2 ang), emission count).  Everything is in base-2 logarithms, the
//     MUFU units' native base; the veto test  r3 < alpha_s(kt) z P(z) / (2 C_R alpha_veto / z)  is multiplied through
//     by its two denominators (both positive), so a trial has no division: LG2, SQRT, EX2 and ~35 FP32 operations.
//     A lane whose jet ended takes the next jet index from the warp's range straight away; ranges come from one
//     global counter (JMC_SHOWER_CHUNK jets per atomic), so lanes and blocks run dry together instead of each owning
//     a fixed 1/grid share of geometrically distributed work;
//   * an accepted emission is an EVENT (log2 z, log2 theta, "last of its jet") pushed on the lane's 4-deep ring in
//     shared memory.  The ACCUMULATOR stage -- ECF top-N/sum, grooming -- pops one event per lane for all lanes
//     together, when JMC_SHOWER_GROUP lanes have events, when a lane's ring is full, or when the warp has no trials
//     left;
//   * a jet's last event turns the accumulators into the observable, parked in a register; the BINNING stage
//     (bin search, shared-memory counter, min/max) runs when JMC_SHOWER_BIN_GROUP lanes have one parked or a lane
//     needs its slot again.
// Jet j always uses Philox stream (first + j): the histogram does not depend on which lane ran which jet.
#ifndef JMC_SHOWER_CHUNK
#define JMC_SHOWER_CHUNK 128
#endif
#ifndef JMC_SHOWER_GROUP
#define JMC_SHOWER_GROUP 24
#endif
#ifndef JMC_SHOWER_BIN_GROUP
#define JMC_SHOWER_BIN_GROUP 24
#endif
#ifndef JMC_SHOWER_TRIALS
#define JMC_SHOWER_TRIALS 3
#endif
#ifndef JMC_SHOWER_SHARE
#define JMC_SHOWER_SHARE 1      // three trials per iteration from two Philox blocks (0: JMC_SHOWER_TRIALS trials, a block each)
#endif
#ifndef JMC_SHOWER_MIN_BLOCKS
#define JMC_SHOWER_MIN_BLOCKS 1
#endif
static constexpr int kRing = 4;
__device__ __forceinline__ float sh_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sh_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sh_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

static size_t shower_fast_smem(int n_edges) {
    const size_t nb = n_edges > 0 ? (size_t)n_edges - 1 : 0;
    const size_t bytes = sizeof(double) * ((size_t)(n_edges > 0 ? n_edges : 0) + (nb + 1) / 2);      // edges, counters
    return bytes + sizeof(float2) * kRing * kShowerThreads;
}

template <bool MLL, int GROOMER, bool QUARK, bool RECORD>
__global__ void __launch_bounds__(kShowerThreads, JMC_SHOWER_MIN_BLOCKS)
shower_fast_kernel(ShowerParams p, uint64_t n, uint64_t seed, uint64_t first, const double* __restrict__ edges,
                   int n_edges, unsigned long long* __restrict__ g_count, double* __restrict__ d_minmax,
                   double* __restrict__ dump_obs, int* __restrict__ dump_n, int* __restrict__ dump_trials,
                   double* __restrict__ dump_chain, int chain_cap, int* error_flag,
                   unsigned long long* __restrict__ next_jet) {
    extern __shared__ double s_raw[];
    const int nb = n_edges > 0 ? n_edges - 1 : 0;
    double* s_edges = s_raw;
    unsigned int* s_count = (unsigned int*)(s_edges + (n_edges > 0 ? n_edges : 0));
    float2* s_ring = (float2*)(s_raw + (n_edges > 0 ? n_edges : 0) + (nb + 1) / 2) + threadIdx.x;   // slot k: s_ring[k * 128]
    for (int k = threadIdx.x; k < n_edges; k += blockDim.x) s_edges[k] = edges[k];
    for (int k = threadIdx.x; k < nb; k += blockDim.x) s_count[k] = 0u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, lanes_below = (1u << lane) - 1u;
    double mn = INFINITY, mx = -INFINITY;

    // launch constants, base-2 logarithms
    constexpr float kLn2 = 0.69314718055994531f, kLog2e = 1.4426950408889634f, k2m24 = 5.9604644775390625e-08f;
    const float k2 = (float)(p.kfac / 0.69314718055994531);               // l'^2 = l^2 - (K / ln 2) log2 r1
    const float inv_beta = (float)(1.0 / p.beta), beta_f = (float)p.beta, beta_sd_f = (float)p.beta_sd;
    const float l2_init = p.lf_init * kLog2e;
    const float l2_cut = (float)(log(p.cutoff) * 1.4426950408889634);
    const float l2_zcut = p.z_cut > 0.0 ? (float)(log(p.z_cut) * 1.4426950408889634) : -INFINITY;
    const float l2_pt = (float)(log(JMC_PT) * 1.4426950408889634);
    // alpha_s(kt) = alpha_MZ / D,  D = 1 + 2 alpha_MZ beta_0 (max(ln(pT z theta), 0) - ln M_Z) = d0 + d1 max(log2 .., 0)
    const float d1 = (float)(2.0 * JMC_ALPHA_MZ * JMC_BETA0) * kLn2;
    const float d0 = (float)(1.0 - 2.0 * JMC_ALPHA_MZ * JMC_BETA0 * log(JMC_MZ));
    // accept  <=>  (i3 + 1/2) D (1 - z) < 2^24 alpha_MZ / (2 C_R alpha_veto) * [z P(z) + z P(1 - z)] (1 - z)
    const float veto_c = (float)(16777216.0 * JMC_ALPHA_MZ / (2.0 * p.cr * p.alpha));
    // the jet's first splitting has z = 1/2, theta = R: a shower whose cutoff is above R / 2 has no emissions at all
    const bool start_ok = -1.0f + l2_init * inv_beta > l2_cut;
    const PhiloxKeys keys(seed);

    // producer
    uint64_t jx = 0;                                  // Philox index of the lane's jet: first + j
    uint32_t blk = 0u;
    float l2 = 0.0f;
    int n_em = 0, n_trials = 0, err = 0;
    bool run = false;
    int q_n = 0, q_head = 0;                          // the lane's ring of events
    uint64_t ev_j = 0;                                // RECORD: jet of the (single) queued event
    // accumulator stage, and the parked observable (NaN = none)
    EcfAcc<float> acc;
    acc.reset(GROOMER, p.z_cut);
    double done_obs = qnan();
    // the warp's range of jet indices (warp-uniform): Philox indices w_base + [w_off, w_cnt)
    uint64_t w_base = 0;
    unsigned w_off = 0u, w_cnt = 0u;
    bool dry = false;                                 // the global counter has reached n
    unsigned want_mask = 0xffffffffu;                 // lanes that need a jet

    while (true) {
        // ---- hand out jets ----
        if (want_mask != 0u && (w_off < w_cnt || !dry)) {
            const unsigned need = __popc(want_mask), left = w_cnt - w_off;
            const unsigned r = __popc(want_mask & lanes_below);
            const bool mine = (want_mask >> lane) & 1u;
            if (need <= left) {
                if (mine) { jx = w_base + (w_off + r); run = true; }
                w_off += need;
            } else {                                  // one atomic for the warp; the old range's rest is used first
                uint64_t fresh = 0;
                if (!dry) {
                    if (lane == 0) fresh = atomicAdd(next_jet, (unsigned long long)JMC_SHOWER_CHUNK);
                    fresh = __shfl_sync(0xffffffffu, fresh, 0);
                }
                const unsigned cnt = (dry || fresh >= n) ? 0u
                                     : (n - fresh < (uint64_t)JMC_SHOWER_CHUNK ? (unsigned)(n - fresh) : (unsigned)JMC_SHOWER_CHUNK);
                if (mine) {
                    if (r < left) { jx = w_base + (w_off + r); run = true; }
                    else if (r - left < cnt) { jx = first + fresh + (r - left); run = true; }
                }
                w_base = first + fresh;
                w_off = need - left < cnt ? need - left : cnt;
                w_cnt = cnt;
                if (cnt < (unsigned)JMC_SHOWER_CHUNK) dry = true;
            }
            if (mine && run) { blk = 0u; l2 = l2_init; n_em = 0; n_trials = 0; }
        }
        const unsigned run_mask = __ballot_sync(0xffffffffu, run);
        const unsigned ev_mask = __ballot_sync(0xffffffffu, q_n != 0);
        const unsigned done_mask = __ballot_sync(0xffffffffu, done_obs == done_obs);
        if (!(run_mask | ev_mask | done_mask)) break;

        // ---- veto trials: JMC_SHOWER_TRIALS per iteration (a lane stops at its first accepted one), so that the
        // warp-level work around them -- hand-out, votes, queue decisions, the push -- is paid once per that many ----
        bool accepted = false, last = false;
        float lz = __int_as_float(0x7fc00000), lth = 0.0f;
        // one trial from two 24-bit uniforms (r1, r2) and the acceptance uniform i3 in units of 2^-24
        auto trial = [&](uint32_t w1, uint32_t w2, float i3) {
            if (RECORD) ++n_trials;
            const float i1 = (float)(w1 >> 8) + 0.5f, i2 = (float)(w2 >> 8) + 0.5f;
            const float lp = -sh_sqrt(fmaf(-k2, sh_lg2(i1 * k2m24), l2 * l2));
            const float a = (i2 * k2m24) * lp;                            // r2 l'
            lz = a - 1.0f;                                                // log2 z = r2 l' - 1
            lth = (lp - a) * inv_beta;                                    // log2 theta = (1 - r2) l' / beta
            l2 = lp;
            bool ok = true;
            if (MLL) {
                const float zf = sh_ex2(lz), omz = 1.0f - zf;
                const float D = fmaf(d1, fmaxf(l2_pt + lz + lth, 0.0f), d0);
                float num;                       // [z P(z) + z P(1 - z)] (1 - z)
                if (QUARK) {
                    const float o2 = omz * omz, z2 = zf * zf;
                    num = (float)JMC_CF * (fmaf(o2, omz, omz) + fmaf(z2, zf, zf));
                } else {
                    const float s2 = fmaf(zf, zf, omz * omz);
                    const float pz = (float)JMC_CA * (2.0f * omz + zf * zf * omz) + (float)(JMC_TF * JMC_NF) * zf * s2;
                    const float p1 = (float)JMC_CA * (2.0f * zf + omz * omz * zf) + (float)(JMC_TF * JMC_NF) * omz * s2;
                    num = fmaf(pz, omz, zf * p1);
                }
                const float rhs = veto_c * num, lhs = D * omz;
                if (rhs > 16777216.0f * lhs) err |= 1;    // the reference raises: acceptance probability above 1
                ok = i3 * lhs < rhs;
            }
            if (ok) {
                if (RECORD && dump_chain && n_em < chain_cap) {
                    double* c = dump_chain + ((size_t)(jx - first) * chain_cap + n_em) * 2;
                    c[0] = (double)sh_ex2(lz);
                    c[1] = (double)sh_ex2(lth);
                }
                ++n_em;
                // the chain continues while z theta > cutoff (angularity_shower, :262-279); tested on accepted
                // emissions only -- the veto loop keeps trying without it (:188-209)
                last = !(lz + lth > l2_cut);
                if (n_em >= kMaxEmissions) { last = true; err |= 2; }
            }
            return ok;
        };
        if (run) {
            if (!start_ok) {
                accepted = last = true;              // event without emission (log2 z = NaN)
            } else {
#if JMC_SHOWER_SHARE
                // Three trials from TWO Philox blocks: r1 and r2 take a word each (24 bits used), the acceptance
                // uniform 21 bits -- the top 21 of a third word for trials 1 and 2, the 11 + 10 bits those left over
                // for trial 3.  (The veto compares r3 with a probability of order 0.1: 2^-21 is 5e-6 of it.)
                uint32_t a[4], b[4];
                philox4x32_10(jx, blk, keys, a);
                if (!MLL) {                      // LL: no veto, the first trial is the emission -- one block is enough
                    ++blk;
                    accepted = trial(a[0], a[1], 0.0f);
                } else {
                philox4x32_10(jx, blk + 1u, keys, b);
                blk += 2u;
                accepted = trial(a[0], a[1], fmaf((float)(a[2] >> 11), 8.0f, 4.0f));
                if (!accepted) accepted = trial(a[3], b[0], fmaf((float)(b[1] >> 11), 8.0f, 4.0f));
                if (!accepted)
                    accepted = trial(b[2], b[3], fmaf((float)(((a[2] & 0x7ffu) << 10) | ((b[1] & 0x7ffu) >> 1)), 8.0f, 4.0f));
                }
#else
#pragma unroll
                for (int t = 0; t < JMC_SHOWER_TRIALS; ++t) {
                    if (!accepted) {
                        uint32_t r[4];
                        philox4x32_10(jx, blk++, keys, r);
                        accepted = trial(r[0], r[1], (float)(r[2] >> 8) + 0.5f);
                    }
                }
#endif
            }
        }
        // ---- accumulator stage: one event of every lane that has one ----
        const bool full = accepted && q_n == kRing;
        const bool drain = RECORD || __popc(ev_mask) >= JMC_SHOWER_GROUP || run_mask == 0u || __any_sync(0xffffffffu, full);
        if (drain) {
            const bool have = q_n != 0;
            float e_lz = 0.0f, e_lth = 0.0f;
            if (have) {
                const float2 e = s_ring[q_head * kShowerThreads];
                e_lz = e.x;
                e_lth = e.y;
                q_head = (q_head + 1) & (kRing - 1);
                --q_n;
            }
            const bool e_last = have && (__float_as_uint(e_lth) & 1u);
            // binning stage, before a lane overwrites its parked observable
            const bool parked = done_obs == done_obs;
            const bool bin_now = RECORD || __popc(done_mask) >= JMC_SHOWER_BIN_GROUP || (run_mask | ev_mask) == 0u
                                 || __any_sync(0xffffffffu, e_last && parked);
            if (bin_now && parked) {
                if (n_edges > 0) {
                    int b;
                    if (p.guess.mode == 0 || done_obs < s_edges[p.guess_offset]) {
                        b = bin_index(done_obs, s_edges, n_edges);
                    } else {
                        b = bin_index_guided(done_obs, s_edges + p.guess_offset, n_edges - p.guess_offset, p.guess);
                        if (b >= 0) b += p.guess_offset;
                    }
                    if (b >= 0) atomicAdd(&s_count[b], 1u);
                }
                mn = fmin(mn, done_obs);
                mx = fmax(mx, done_obs);
                done_obs = qnan();
            }
            if (have) {
                if (e_lz == e_lz) {
                    const float zf = sh_ex2(e_lz);
                    if (GROOMER == 2) {
                        acc.add_rss(zf, sh_ex2(beta_f * e_lth), p);
                    } else {
                        float c = sh_ex2(fmaf(beta_f, e_lth, e_lz));                  // C = z theta^beta [(1 - z)]
                        if (p.obs_mll) c *= 1.0f - zf;
                        acc.add(c, e_lz > fmaf(beta_sd_f, e_lth, l2_zcut), GROOMER);  // z > z_cut theta^beta_sd
                    }
                }
                if (e_last) {
                    done_obs = acc.template result<GROOMER>(p);
                    if (RECORD && dump_obs) dump_obs[ev_j] = done_obs;
                    acc.reset(GROOMER, p.z_cut);
                }
            }
        }
        // ---- push the new event; a lane whose jet ended asks for the next one ----
        if (accepted) {
            // "last" travels in the lowest mantissa bit of log2 theta
            s_ring[((q_head + q_n) & (kRing - 1)) * kShowerThreads] =
                make_float2(lz, __uint_as_float((__float_as_uint(lth) & ~1u) | (last ? 1u : 0u)));
            ++q_n;
            if (RECORD) ev_j = jx - first;
            if (last) {
                if (RECORD && dump_n) dump_n[jx - first] = n_em;
                if (RECORD && dump_trials) dump_trials[jx - first] = n_trials;
                run = false;
            }
        }
        want_mask = __ballot_sync(0xffffffffu, accepted && last);
    }
    if (err) atomicOr(error_flag, err);
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
        if (s_count[k]) atomicAdd(&g_count[k], (unsigned long long)s_count[k]);
    if (d_minmax) block_minmax_to_global(mn, mx, d_minmax);
}

// ECF of stored event records (replay form of the pickers): one jet per thread walks its (z, theta) chain through the
// same accumulator the shower kernel carries per lane, in FP64 and the reference's operation order.
// chains: [n, cap, 2]; a warp stages 32 jets x 8 emissions at a time through shared memory so that the reads of a
// record are 128-byte rows rather than one 16-byte element per lane.
__global__ void __launch_bounds__(kShowerThreads)
chain_ecf_kernel(ShowerParams p, const int* __restrict__ lens, const double* __restrict__ chains, uint64_t n, int cap,
                 double* __restrict__ out, int* __restrict__ truncated) {
    __shared__ double2 s_tile[kShowerThreads / 32][32][9];      // [warp][jet][emission], padded against bank conflicts
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t base = r * stride + (uint64_t)blockIdx.x * blockDim.x + warp * 32;   // first jet of this warp
        const uint64_t j = base + lane;
        const int stored = j < n ? lens[j] : 0;
        if (stored > cap) *truncated = 1;                     // a record longer than its slots (any writer wins)
        const int len = min(stored, cap);
        int longest = len;
        for (int o = 16; o > 0; o >>= 1) longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, o));
        EcfAcc<double> acc;
        acc.reset(p.groomer, p.z_cut);
        for (int e0 = 0; e0 < longest; e0 += 8) {
            // 8 lanes read the 8 emissions (128 B) of one jet: 4 jets per load instruction, 8 instructions per tile
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned jet = q * 4 + (lane >> 3), em = lane & 7u;
                const uint64_t gj = base + jet;
                double2 v = make_double2(0.0, 0.0);
                if (gj < n && e0 + (int)em < cap) v = *(const double2*)(chains + ((size_t)gj * cap + e0 + em) * 2);
                s_tile[warp][jet][em] = v;
            }
            __syncwarp();
            for (int e = 0; e < 8 && e0 + e < len; ++e) {
                const double2 zt = s_tile[warp][lane][e];
                if (p.groomer == 2) acc.add_rss(zt.x, np_pow(zt.y, p.beta), p);
                else acc.add(ecf_ungroomed(zt.x, zt.y, p.beta, p.obs_mll ? JMC_ACC_MLL : JMC_ACC_LL),
                             zt.x > p.z_cut * np_pow(zt.y, p.beta_sd), p.groomer);
            }
            __syncwarp();
        }
        if (j < n) out[j] = acc.result(p);
    }
}

// weight 1 per jet: sum w = sum w^2 = count
__global__ void shower_sums_kernel(const unsigned long long* __restrict__ before, const unsigned long long* __restrict__ after,
                                   int nb, double* __restrict__ sum_w, double* __restrict__ sum_w2) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nb; k += gridDim.x * blockDim.x) {
        const double d = (double)(after[k] - before[k]);
        sum_w[k] += d;
        sum_w2[k] += d;
    }
}

static int make_shower_params(const jmc_integrand* g, ShowerParams* out) {
    JMC_REQUIRE(g != nullptr, "integrand is NULL");
    JMC_REQUIRE(g->jet == JMC_QUARK || g->jet == JMC_GLUON, "jet type %d is not a supported jet type.", g->jet);
    JMC_REQUIRE(g->split_acc == JMC_ACC_LL || g->split_acc == JMC_ACC_MLL, "shower accuracy must be LL or MLL");
    JMC_REQUIRE(g->obs_acc == JMC_ACC_LL || g->obs_acc == JMC_ACC_MLL, "Accuracy must be LL or MLL");
    JMC_REQUIRE(g->radius > 0.0, "Jet radius must be greater than zero!");
    JMC_REQUIRE(g->beta > 0.0, "shower: beta must be positive");
    JMC_REQUIRE(g->shower_cutoff > 0.0, "shower: cutoff must be positive");
    JMC_REQUIRE(g->shower_n_emissions >= 0 && g->shower_n_emissions <= kTopN, "shower: n_emissions must be 1..%d or 0 ('all')", kTopN);
    if (g->shower_groomer) JMC_REQUIRE(g->z_cut > 0.0 && g->z_cut < 0.5, "zc=%g is an invalid grooming parameter.", g->z_cut);
    ShowerParams p;
    p.jet = g->jet;
    p.mll_shower = g->split_acc == JMC_ACC_MLL;
    p.obs_mll = g->obs_acc == JMC_ACC_MLL;
    p.groomer = g->shower_groomer;
    JMC_REQUIRE(p.groomer >= 0 && p.groomer <= 2, "shower: groomer must be 0 (none), 1 (Soft Drop) or 2 (RSS)");
    p.rss_pre = (g->shower_rss_flags & JMC_RSS_PRECRIT) != 0;
    p.rss_crit = (g->shower_rss_flags & JMC_RSS_CRIT) != 0;
    p.rss_sub = (g->shower_rss_flags & JMC_RSS_SUB) != 0;
    p.rss_few_pres = (g->shower_rss_flags & JMC_RSS_FEW_PRES) != 0;
    p.f_soft = g->f_soft;
    if (p.groomer == 2)
        JMC_REQUIRE(g->f_soft > 1.0 / 3.0, "f = %g is less than 1/3, and the grooming may groom away harder branches first.",
                    g->f_soft);
    p.n_emissions = g->shower_n_emissions;
    p.beta = g->beta; p.radius = g->radius; p.cutoff = g->shower_cutoff; p.z_cut = g->z_cut; p.beta_sd = g->beta_sd;
    p.cr = g->jet == JMC_QUARK ? JMC_CF : JMC_CA;
    p.alpha = p.mll_shower ? JMC_ALPHA_VETO : JMC_ALPHA_FIXED;
    p.kfac = JMC_PI * g->beta / (p.cr * p.alpha);
    p.ang_init = std::pow(g->radius, g->beta) / 2.0;
    p.lf_init = (float)(g->beta * std::log(g->radius));
    p.guess = BinGuess{0, 0.f, 0.f};
    p.guess_offset = 0;
    *out = p;
    return JMC_OK;
}

// log-uniform edges (possibly after a first zero-bin edge) get an affine first guess
static void shower_bin_guess(const std::vector<double>& e, ShowerParams* p) {
    const int ne = (int)e.size();
    for (int off = 0; off <= 1; ++off) {
        const int m = ne - off;
        if (m < 8 || !(e[off] > 1e-30 && e[ne - 1] < 1e30)) continue;
        const double b0 = std::log2(e[off]), d = (std::log2(e[ne - 1]) - b0) / (m - 1);
        bool ok = d > 0.0 && std::isfinite(d);
        for (int k = 0; ok && k < m; ++k) ok = std::fabs(std::log2(e[off + k]) - (b0 + k * d)) < 0.25 * d;
        if (ok) {
            p->guess = BinGuess{2, (float)b0, (float)(1.0 / d)};
            p->guess_offset = off;
            return;
        }
    }
}

static int shower_launch(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                         const double* d_edges, int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count,
                         double* d_minmax, double* d_obs, int* d_n_em, int* d_trials, void* stream,
                         double* d_chain = nullptr, int chain_cap = 0) {
    ShowerParams p;
    int rc = make_shower_params(g, &p);
    if (rc) return rc;
    if (d_edges) {
        JMC_REQUIRE(n_edges >= 2, "jmc_run: need at least 2 edges");
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_run: histogram outputs missing");
    } else {
        n_edges = 0;
    }
    if (n == 0) return JMC_OK;
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    if (n_edges >= 8 && n >= (1u << 16)) {
        std::vector<double> e(n_edges);
        JMC_CUDA(cudaMemcpyAsync(e.data(), d_edges, sizeof(double) * n_edges, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        JMC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        shower_bin_guess(e, &p);
    }
    const size_t nb = n_edges > 0 ? (size_t)n_edges - 1 : 0;
    const size_t smem = precision == JMC_F64 ? sizeof(double) * (size_t)n_edges + sizeof(unsigned int) * nb
                                             : shower_fast_smem(n_edges);
    if (smem > (size_t)di.max_smem_optin)
        return set_error(JMC_ELIMIT, "jmc_run: %d edges exceed shared memory", n_edges);
    cudaStream_t st = (cudaStream_t)stream;
    char* scr;
    rc = scratch(256 + sizeof(unsigned long long) * nb, (void**)&scr, (void*)st);
    if (rc) return rc;
    int* d_err = (int*)scr;
    unsigned long long* d_next = (unsigned long long*)(scr + 8);         // FP32 kernel: next unassigned jet
    unsigned long long* d_before = (unsigned long long*)(scr + 256);
    JMC_CUDA(cudaMemsetAsync(scr, 0, 16, st));
    if (nb) JMC_CUDA(cudaMemcpyAsync(d_before, d_count, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToDevice, st));
    const uint64_t want = (n + kShowerThreads - 1) / kShowerThreads;
    const bool record = d_obs || d_n_em || d_trials || d_chain;
    const bool exact = precision == JMC_F64;
    // instantiation for (shower accuracy, groomer, jet type, record); the grid fills every SM with resident blocks
    auto launch = [&](auto kern, auto... extra) -> cudaError_t {
        static int per_sm = 0;                 // per instantiation (the lambda body is instantiated per kernel)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int blocks = per_sm;
        if (blocks == 0 || smem > 8192) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, kShowerThreads, smem);
            if (e != cudaSuccess) return e;
            if (blocks < 1) blocks = 1;
            if (smem <= 8192) per_sm = blocks;
        }
        const uint64_t cap = (uint64_t)di.sm_count * blocks;
        const int grid = (int)(want < cap ? want : cap);
        kern<<<grid, kShowerThreads, smem, st>>>(p, n, seed, first_index, d_edges, n_edges, (unsigned long long*)d_count,
                                                 d_minmax, d_obs, d_n_em, d_trials, d_chain, chain_cap, d_err, extra...);
        return cudaSuccess;
    };
#define JMC_SH5(M, G, Q, R) (exact ? launch(shower_kernel<M, G, Q, R>) : launch(shower_fast_kernel<M, G, Q, R>, d_next))
#define JMC_SH4(M, G, Q) (record ? JMC_SH5(M, G, Q, true) : JMC_SH5(M, G, Q, false))
#define JMC_SH3(M, G) (p.jet == JMC_QUARK ? JMC_SH4(M, G, true) : JMC_SH4(M, G, false))
#define JMC_SH2(M) (p.groomer == 0 ? JMC_SH3(M, 0) : (p.groomer == 1 ? JMC_SH3(M, 1) : JMC_SH3(M, 2)))
    JMC_CUDA(p.mll_shower ? JMC_SH2(true) : JMC_SH2(false));
#undef JMC_SH2
#undef JMC_SH3
#undef JMC_SH4
#undef JMC_SH5
    JMC_LAUNCHED(exact ? "shower_kernel" : "shower_fast_kernel");
    if (nb) {
        shower_sums_kernel<<<(int)((nb + 255) / 256), 256, 0, st>>>(d_before, (const unsigned long long*)d_count, (int)nb,
                                                                     d_sum_w, d_sum_w2);
        JMC_LAUNCHED("shower_sums_kernel");
    }
    // the reference raises inside the veto loop when the overestimate fails; report it the same way
    int h_err = 0;
    JMC_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    JMC_CUDA(cudaStreamSynchronize(st));
    if (h_err & 1) return set_error(JMC_EDOMAIN, "The pdf must be everywhere less than the proposed pdf!");
    // the reference recurses until z theta <= cutoff; the device keeps at most kMaxEmissions per jet.  A jet that
    // reaches the cap is reported instead of being returned truncated.
    if (h_err & 2)
        return set_error(JMC_ELIMIT, "shower: a jet reached the cap of %d emissions (cutoff %g too low for the device "
                                     "record)", kMaxEmissions, g->shower_cutoff);
    return JMC_OK;
}

int run_shower(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
               const double* d_edges, int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count,
               double* d_minmax, void* stream) {
    JMC_REQUIRE(d_edges != nullptr || d_minmax != nullptr, "jmc_run: neither edges nor minmax requested");
    return shower_launch(g, n, seed, first_index, precision, d_edges, n_edges, d_sum_w, d_sum_w2, d_count, d_minmax,
                         nullptr, nullptr, nullptr, stream);
}

}  // namespace jmc

extern "C" int jmc_shower_dump(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                               double* d_obs, int32_t* d_n_emissions, int32_t* d_n_trials, void* stream) {
    JMC_REQUIRE(g != nullptr && g->kind == JMC_KIND_SHOWER, "jmc_shower_dump: integrand must be a shower");
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_shower_dump: precision must be JMC_F64 or JMC_F32");
    // the output arrays are indexed by the local jet number 0..n-1; jet i is stream index first_index + i
    return jmc::shower_launch(g, n, seed, first_index, precision, nullptr, 0, nullptr, nullptr, nullptr, nullptr, d_obs,
                              d_n_emissions, d_n_trials, stream);
}

extern "C" int jmc_chain_ecf(const jmc_integrand* g, const int32_t* d_n_emissions, const double* d_chains, uint64_t n,
                             int max_emissions, double* d_obs, void* stream) {
    JMC_REQUIRE(g != nullptr && g->kind == JMC_KIND_SHOWER, "jmc_chain_ecf: integrand must be a shower");
    jmc::ShowerParams p;
    int rc = jmc::make_shower_params(g, &p);
    if (rc) return rc;
    JMC_REQUIRE(max_emissions >= 1, "jmc_chain_ecf: need room for at least one emission");
    if (n == 0) return JMC_OK;
    JMC_REQUIRE(d_n_emissions && d_chains && d_obs, "jmc_chain_ecf: NULL array");
    jmc::DeviceInfo di;
    if (jmc::device_info(&di)) return JMC_ENODEV;
    const uint64_t want = (n + jmc::kShowerThreads - 1) / jmc::kShowerThreads, cap = (uint64_t)di.sm_count * 8;
    int* d_flag;
    rc = jmc::scratch(sizeof(int), (void**)&d_flag, stream);
    if (rc) return rc;
    JMC_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), (cudaStream_t)stream));
    jmc::chain_ecf_kernel<<<(int)(want < cap ? want : cap), jmc::kShowerThreads, 0, (cudaStream_t)stream>>>(
        p, d_n_emissions, d_chains, n, max_emissions, d_obs, d_flag);
    JMC_LAUNCHED("chain_ecf_kernel");
    int h_flag = 0;
    JMC_CUDA(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    JMC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (h_flag)
        return jmc::set_error(JMC_ELIMIT, "emission chains were truncated: raise max_emissions (a record is longer than "
                                          "its %d slots)", max_emissions);
    return JMC_OK;
}

extern "C" int jmc_shower_chains(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                                 int max_emissions, double* d_chains, int32_t* d_n_emissions, void* stream) {
    JMC_REQUIRE(g != nullptr && g->kind == JMC_KIND_SHOWER, "jmc_shower_chains: integrand must be a shower");
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_shower_chains: precision must be JMC_F64 or JMC_F32");
    JMC_REQUIRE(max_emissions >= 1 && d_chains && d_n_emissions, "jmc_shower_chains: need room for at least one emission");
    return jmc::shower_launch(g, n, seed, first_index, precision, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                              d_n_emissions, nullptr, stream, d_chains, max_emissions);
}
