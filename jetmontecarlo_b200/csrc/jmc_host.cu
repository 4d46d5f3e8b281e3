// Host side of libjmc_b200.so: error strings, device bookkeeping, argument
// validation shared by all kernels, precision dispatch of the fused path, and
// the host-buffer entry points a reference-side binding would call.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "jmc_common.cuh"

namespace jmc {

static thread_local std::string g_error;
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return set_error(JMC_ECUDA, "CUDA error in %s: %s", what, cudaGetErrorString(e));
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int device_info(DeviceInfo* out) {
    static thread_local DeviceInfo cache = {-1, 0, 0};
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (cache.device != dev) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
        if (prop.major != 10)
            return set_error(JMC_ENODEV, "libjmc_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major,
                             prop.minor);
        cache.device = dev;
        cache.sm_count = prop.multiProcessorCount;
        cache.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    }
    *out = cache;
    return JMC_OK;
}

// Small device scratch (poison flags, FP32 edges, staged batch parameters), one allocation per
// (host thread, device, stream) so that launches queued on different streams never share it; grows on demand.
int scratch(size_t bytes, void** ptr, void* stream) {
    struct Slot { int device; void* stream; void* p; size_t cap; };
    static thread_local std::vector<Slot> slots;
    int dev = -1;
    JMC_CUDA(cudaGetDevice(&dev));
    for (auto& s : slots) {
        if (s.device != dev || s.stream != stream) continue;
        if (s.cap < bytes) {
            JMC_CUDA(cudaFree(s.p));           // synchronises the device: nothing still reads the old block
            s.p = nullptr; s.cap = 0;
            JMC_CUDA(cudaMalloc(&s.p, bytes));
            s.cap = bytes;
        }
        *ptr = s.p;
        return JMC_OK;
    }
    Slot s{dev, stream, nullptr, bytes < 256 ? 256 : bytes};
    JMC_CUDA(cudaMalloc(&s.p, s.cap));
    slots.push_back(s);
    *ptr = s.p;
    return JMC_OK;
}

// Pinned host staging of the same kind (uploads of a parameter grid's edges): one block per host thread, grows on
// demand.  The caller synchronises the stream that reads it before asking again.
int pinned_scratch(size_t bytes, void** ptr) {
    static thread_local void* block = nullptr;
    static thread_local size_t cap = 0;
    if (cap < bytes) {
        if (block) JMC_CUDA(cudaFreeHost(block));
        block = nullptr; cap = 0;
        JMC_CUDA(cudaHostAlloc(&block, bytes, cudaHostAllocDefault));
        cap = bytes;
    }
    *ptr = block;
    return JMC_OK;
}

static bool valid_eps(double e) { return e > 0.0 && e < 1.0; }

// Validation mirrors the reference's assertions; the Python layer raises
// AssertionError with the reference's own messages before it ever gets here.
// ref: montecarlo/sampler.py:65-81, numerics/radiators/samplers.py:12-18,116,298
int make_run_params(const jmc_integrand* g, RunParams* out) {
    JMC_REQUIRE(g != nullptr, "integrand is NULL");
    JMC_REQUIRE(g->kind >= 0 && g->kind < JMC_KIND_COUNT, "unknown integrand kind %d", g->kind);
    JMC_REQUIRE(g->method == JMC_LIN || g->method == JMC_LOG, "%dis not a supported sample method.", g->method);
    JMC_REQUIRE(g->jet == JMC_QUARK || g->jet == JMC_GLUON, "jet type %d is not a supported jet type.", g->jet);
    if (g->method == JMC_LOG)
        JMC_REQUIRE(valid_eps(g->epsilon), "epsilon=%g is invalid for log sampling.", g->epsilon);
    const bool uses_zc = g->kind == JMC_KIND_CRIT || g->kind == JMC_KIND_CRIT_SUB || g->kind == JMC_KIND_PRE_CRIT_SUB;
    if (uses_zc) JMC_REQUIRE(g->z_cut > 0.0 && g->z_cut < 0.5, "zc=%g is an invalid grooming parameter.", g->z_cut);
    if (g->kind != JMC_KIND_SIMPLE) JMC_REQUIRE(g->radius > 0.0, "Jet radius must be greater than zero!");
    if (g->kind == JMC_KIND_PRE_CRIT_SUB && uses_zc) {
        JMC_REQUIRE(g->fixed_coupling != 0, "Can only accept fixed coupling for now!");
        JMC_REQUIRE(g->method != JMC_LOG || valid_eps(g->epsilon), "Weight requires valid cutoff for zero bin.");
    }
    JMC_REQUIRE(g->obs_acc == JMC_ACC_LL || g->obs_acc == JMC_ACC_MLL, "Accuracy must be LL or MLL");
    JMC_REQUIRE(g->split_acc >= JMC_ACC_LL && g->split_acc <= JMC_ACC_MLL, "Invalid accuracy.");
    RunParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = *g;
    if (g->method == JMC_LOG) {
        p.log_eps = std::log(g->epsilon);
        p.lw_crit_z = std::log(0.5 - g->z_cut);
        p.lw_radius = std::log(g->radius - 0.0);
        p.lw_half = std::log(0.5 - 0.0);
        p.lw_pre = g->z_cut > 0.0 ? std::log(g->z_cut - 0.0) : 0.0;
        p.lw_simple = g->hi > g->lo ? std::log(g->hi - g->lo) : 0.0;
    }
    *out = p;
    return JMC_OK;
}

int run_exact(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, const double* d_edges,
              int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, void* stream);
int run_fast(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, const double* d_edges,
             int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, void* stream);
int run_fast_batch(const jmc_integrand* g, int n_points, uint64_t n, uint64_t seed, uint64_t first_index,
                   const double* h_edges, int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count,
                   double* d_minmax, void* stream);
int dump_exact(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, double* d_samples,
               double* d_obs, double* d_weight, void* stream);
int dump_fast(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, double* d_samples,
              double* d_obs, double* d_weight, void* stream);
int run_shower(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
               const double* d_edges, int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count,
               double* d_minmax, void* stream);

// device buffer freed at scope exit (host-side helpers below synchronise before returning)
struct DevBufE {
    void* p = nullptr;
    ~DevBufE() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        JMC_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
        return JMC_OK;
    }
};

static int sampler_area(int kind, int method, double eps, double zc, double radius, double lo, double hi,
                        double* area) {
    // ref: samplers.py:30-36 (critical), :132-138 (precritical), :196-202 (ungroomed); sampler.py:118-122 (simple)
    if (method == JMC_LOG) {
        JMC_REQUIRE(valid_eps(eps), "epsilon=%g is invalid for log sampling.", eps);
        const double l = std::log(1.0 / eps);
        *area = (kind == JMC_SAMPLER_CRITICAL || kind == JMC_SAMPLER_UNGROOMED) ? std::pow(l, 2.0) : l;
        return JMC_OK;
    }
    switch (kind) {
        case JMC_SAMPLER_SIMPLE: *area = hi - lo; break;
        case JMC_SAMPLER_CRITICAL: *area = (1.0 / 2.0 - zc) * radius; break;
        case JMC_SAMPLER_PRECRITICAL: *area = zc; break;
        default: *area = radius / 2.0; break;
    }
    return JMC_OK;
}

}  // namespace jmc

using namespace jmc;

extern "C" int jmc_abi_version(void) { return JMC_ABI_VERSION; }
extern "C" const char* jmc_last_error(void) { return g_error.c_str(); }
extern "C" uint64_t jmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int jmc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int jmc_set_device(int device) {
    JMC_CUDA(cudaSetDevice(device));
    return JMC_OK;
}
extern "C" int jmc_sm_count(void) {
    DeviceInfo di;
    if (device_info(&di)) return JMC_ENODEV;
    return di.sm_count;
}

extern "C" int jmc_sampler_area(const jmc_sampler* s, double* area) {
    JMC_REQUIRE(s && area, "jmc_sampler_area: NULL argument");
    JMC_REQUIRE(s->method == JMC_LIN || s->method == JMC_LOG, "%dis not a supported sample method.", s->method);
    return sampler_area(s->kind, s->method, s->epsilon, s->zc, s->radius, s->lo, s->hi, area);
}

extern "C" int jmc_integrand_uniforms(const jmc_integrand* g) {
    if (!g) return set_error(JMC_EINVAL, "integrand is NULL");
    switch (g->kind) {
        case JMC_KIND_RADIATOR: case JMC_KIND_CRIT: return 2;
        case JMC_KIND_CRIT_SUB: return 4;
        case JMC_KIND_PRE_CRIT_SUB: return 6;
        case JMC_KIND_SIMPLE: return 1;
        case JMC_KIND_SHOWER: return 0;   // variable: 2-3 per veto trial
        default: return set_error(JMC_EINVAL, "unknown integrand kind %d", g->kind);
    }
}

extern "C" int jmc_integrand_area(const jmc_integrand* g, double* area) {
    JMC_REQUIRE(g && area, "jmc_integrand_area: NULL argument");
    double a = 1.0, t = 1.0;
    int rc = JMC_OK;
    auto mul = [&](int kind) {
        if (rc) return;
        rc = sampler_area(kind, g->method, g->epsilon, g->z_cut, g->radius, g->lo, g->hi, &t);
        a *= t;
    };
    switch (g->kind) {
        case JMC_KIND_RADIATOR: mul(JMC_SAMPLER_UNGROOMED); break;
        case JMC_KIND_CRIT: mul(JMC_SAMPLER_CRITICAL); break;
        case JMC_KIND_CRIT_SUB: mul(JMC_SAMPLER_CRITICAL); mul(JMC_SAMPLER_UNGROOMED); break;
        // order of the reference's product: crit * pre * sub (comparison_utils_probabilistic.py:452-454)
        case JMC_KIND_PRE_CRIT_SUB: mul(JMC_SAMPLER_CRITICAL); mul(JMC_SAMPLER_PRECRITICAL); mul(JMC_SAMPLER_UNGROOMED); break;
        case JMC_KIND_SIMPLE: mul(JMC_SAMPLER_SIMPLE); break;
        case JMC_KIND_SHOWER: a = 1.0; break;
        default: return set_error(JMC_EINVAL, "unknown integrand kind %d", g->kind);
    }
    if (rc) return rc;
    // RADIATOR, lin sampling: the reference rescales the AREA by theta_crit (generation.py:316-317)
    if (g->kind == JMC_KIND_RADIATOR && g->method == JMC_LIN) a = a * g->theta_scale;
    *area = a;
    return JMC_OK;
}

extern "C" int jmc_run(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                       const double* d_edges, int n_edges, double* d_sum_w, double* d_sum_w2, uint64_t* d_count,
                       double* d_minmax, void* stream) {
    JMC_REQUIRE(g != nullptr, "jmc_run: integrand is NULL");
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_run: precision must be JMC_F64 or JMC_F32");
    // Per-block bin counts are 32-bit in shared memory; one launch therefore takes at most 2^38 samples
    // (< 2^32 per block on any grid >= 64 blocks) and larger jobs are issued as consecutive index ranges.
    const uint64_t kMaxPerLaunch = 1ull << 38;
    do {
        const uint64_t chunk = n < kMaxPerLaunch ? n : kMaxPerLaunch;
        int rc;
        if (g->kind == JMC_KIND_SHOWER)
            rc = run_shower(g, chunk, seed, first_index, precision, d_edges, n_edges, d_sum_w, d_sum_w2, d_count,
                            d_minmax, stream);
        else if (precision == JMC_F64)
            rc = run_exact(g, chunk, seed, first_index, d_edges, n_edges, d_sum_w, d_sum_w2, d_count, d_minmax, stream);
        else
            rc = run_fast(g, chunk, seed, first_index, d_edges, n_edges, d_sum_w, d_sum_w2, d_count, d_minmax, stream);
        if (rc) return rc;
        n -= chunk;
        first_index += chunk;
    } while (n > 0);
    return JMC_OK;
}

extern "C" int jmc_run_batch(const jmc_integrand* g, int n_points, uint64_t n, uint64_t seed, uint64_t first_index,
                             int precision, const double* h_edges, int n_edges, double* d_sum_w, double* d_sum_w2,
                             uint64_t* d_count, double* d_minmax, void* stream) {
    JMC_REQUIRE(g != nullptr && n_points >= 1, "jmc_run_batch: need at least one parameter point");
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_run_batch: precision must be JMC_F64 or JMC_F32");
    if (precision == JMC_F32 && g[0].kind != JMC_KIND_SHOWER && g[0].kind != JMC_KIND_PRE_CRIT_SUB
        && g[0].kind != JMC_KIND_SIMPLE) {
        // a point may run in a single block whose bin counts are 32-bit: at most 2^31 samples per launch
        const uint64_t kMaxPerLaunch = 1ull << 31;
        do {
            const uint64_t chunk = n < kMaxPerLaunch ? n : kMaxPerLaunch;
            int rc = run_fast_batch(g, n_points, chunk, seed, first_index, h_edges, n_edges, d_sum_w, d_sum_w2, d_count,
                                    d_minmax, stream);
            if (rc) return rc;
            n -= chunk;
            first_index += chunk;
        } while (n > 0);
        return JMC_OK;
    }
    // FP64 (and the kinds without a batched kernel): one launch per point on the same stream
    const size_t nb = n_edges > 0 ? (size_t)n_edges - 1 : 0;
    DevBufE edges;
    if (h_edges) {
        int rc = edges.alloc(sizeof(double) * (size_t)n_points * n_edges);
        if (rc) return rc;
        JMC_CUDA(cudaMemcpyAsync(edges.p, h_edges, sizeof(double) * (size_t)n_points * n_edges, cudaMemcpyHostToDevice,
                                 (cudaStream_t)stream));
    }
    for (int i = 0; i < n_points; ++i) {
        int rc = jmc_run(&g[i], n, seed, first_index, precision, h_edges ? (const double*)edges.p + (size_t)i * n_edges : nullptr,
                         n_edges, d_sum_w ? d_sum_w + i * nb : nullptr, d_sum_w2 ? d_sum_w2 + i * nb : nullptr,
                         d_count ? d_count + i * nb : nullptr, d_minmax ? d_minmax + 2 * (size_t)i : nullptr, stream);
        if (rc) return rc;
    }
    JMC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return JMC_OK;
}

extern "C" int jmc_dump(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
                        double* d_samples, double* d_obs, double* d_weight, void* stream) {
    JMC_REQUIRE(g != nullptr, "jmc_dump: integrand is NULL");
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_dump: precision must be JMC_F64 or JMC_F32");
    JMC_REQUIRE(g->kind != JMC_KIND_SHOWER, "jmc_dump: use jmc_shower_dump for the shower");
    if (precision == JMC_F64) return dump_exact(g, n, seed, first_index, d_samples, d_obs, d_weight, stream);
    return dump_fast(g, n, seed, first_index, d_samples, d_obs, d_weight, stream);
}

// ---------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------
namespace {

// device buffer that frees itself; the host entry points are synchronous
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        JMC_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
        return JMC_OK;
    }
    template <typename T> T* as() { return (T*)p; }
};

int finalize_to_host(DevBuf& hist, int n_edges, const double* d_edges, double area, uint64_t n_total, int bc_kind,
                     double bc_value, double* h_density, double* h_density_err, double* h_integral,
                     double* h_integral_err, double* h_sum_w, double* h_sum_w2, uint64_t* h_count) {
    const size_t nb = (size_t)n_edges - 1;
    double* d_sum_w = hist.as<double>();
    double* d_sum_w2 = d_sum_w + nb;
    uint64_t* d_count = (uint64_t*)(d_sum_w2 + nb);
    DevBuf out;
    int rc = out.alloc(sizeof(double) * (4 * nb + 1));
    if (rc) return rc;
    double* d_density = out.as<double>();
    double* d_err = d_density + nb;
    double* d_ierr = d_err + nb;
    double* d_integral = d_ierr + nb;
    rc = jmc_finalize(d_sum_w, d_sum_w2, d_count, d_edges, n_edges, area, n_total, bc_kind, bc_value, d_density, d_err,
                      d_integral, d_ierr, nullptr);
    if (rc) return rc;
    if (h_density) JMC_CUDA(cudaMemcpy(h_density, d_density, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_density_err) JMC_CUDA(cudaMemcpy(h_density_err, d_err, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_integral_err) JMC_CUDA(cudaMemcpy(h_integral_err, d_ierr, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_integral) JMC_CUDA(cudaMemcpy(h_integral, d_integral, sizeof(double) * (nb + 1), cudaMemcpyDeviceToHost));
    if (h_sum_w) JMC_CUDA(cudaMemcpy(h_sum_w, d_sum_w, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_sum_w2) JMC_CUDA(cudaMemcpy(h_sum_w2, d_sum_w2, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_count) JMC_CUDA(cudaMemcpy(h_count, d_count, sizeof(uint64_t) * nb, cudaMemcpyDeviceToHost));
    JMC_CUDA(cudaDeviceSynchronize());
    return JMC_OK;
}

}  // namespace

extern "C" int jmc_set_density_host(const double* h_obs, const double* h_weights, uint64_t n, const double* h_edges,
                                    int n_edges, double area, int bc_kind, double bc_value, double* h_density,
                                    double* h_density_err, double* h_integral, double* h_integral_err,
                                    double* h_sum_w, double* h_sum_w2, uint64_t* h_count) {
    JMC_REQUIRE(n_edges >= 2 && h_edges, "jmc_set_density_host: Need bins to produce density histogram.");
    JMC_REQUIRE(n == 0 || (h_obs && h_weights), "jmc_set_density_host: NULL input array");
    const size_t nb = (size_t)n_edges - 1;
    DevBuf in, edges, hist;
    int rc = in.alloc(sizeof(double) * 2 * n);
    if (!rc) rc = edges.alloc(sizeof(double) * n_edges);
    if (!rc) rc = hist.alloc(sizeof(double) * 3 * nb);
    if (rc) return rc;
    double* d_obs = in.as<double>();
    double* d_w = d_obs + n;
    JMC_CUDA(cudaMemcpy(d_obs, h_obs, sizeof(double) * n, cudaMemcpyHostToDevice));
    JMC_CUDA(cudaMemcpy(d_w, h_weights, sizeof(double) * n, cudaMemcpyHostToDevice));
    JMC_CUDA(cudaMemcpy(edges.p, h_edges, sizeof(double) * n_edges, cudaMemcpyHostToDevice));
    JMC_CUDA(cudaMemset(hist.p, 0, sizeof(double) * 3 * nb));
    rc = jmc_hist(d_obs, d_w, n, edges.as<double>(), n_edges, hist.as<double>(), hist.as<double>() + nb,
                  (uint64_t*)(hist.as<double>() + 2 * nb), nullptr);
    if (rc) return rc;
    return finalize_to_host(hist, n_edges, edges.as<double>(), area, n, bc_kind, bc_value, h_density, h_density_err,
                            h_integral, h_integral_err, h_sum_w, h_sum_w2, h_count);
}

// ---------------------------------------------------------------------------
// handle: cached job description, device workspace and pinned staging, so that a job issued from host buffers costs
// its kernels plus ONE device-to-host copy and ONE synchronisation (jmc_create / jmc_destroy)
// ---------------------------------------------------------------------------
struct jmc_handle {
    int device = -1;
    // cached job: integrand + edges as last seen, and what was derived from them
    bool has_job = false;
    int precision = -1;
    jmc_integrand g;
    std::vector<double> edges;
    FastPlan* plan = nullptr;
    // device workspace: [poison 256 B][edges f64][edges f32][sum_w, sum_w2, count][density, err, ierr, integral]
    char* d_ws = nullptr;
    size_t ws_cap = 0;
    int ws_edges = 0;
    // pinned host staging: [edges f64][edges f32][results]
    char* h_pin = nullptr;
    size_t pin_cap = 0;
    cudaEvent_t staged = nullptr;      // last asynchronous copy that reads the pinned buffer
    cudaStream_t stream = nullptr;     // stream of the host-buffer entry points
};

namespace {

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WsLayout {
    size_t off_e64, off_e32, off_hist, off_out, total;
    explicit WsLayout(int n_edges) {
        const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1), nb = ne - 1;
        off_e64 = 256;
        off_e32 = off_e64 + align256(sizeof(double) * ne);
        off_hist = off_e32 + align256(sizeof(float) * ne);
        off_out = off_hist + sizeof(double) * 3 * nb;           // contiguous with the histograms: one copy back
        total = align256(off_out + sizeof(double) * (4 * nb + 1));
    }
};

int handle_reserve(jmc_handle* h, int n_edges) {
    const WsLayout L(n_edges);
    if (h->ws_cap < L.total) {
        if (h->d_ws) JMC_CUDA(cudaFree(h->d_ws));           // synchronises: nothing still uses the old block
        h->d_ws = nullptr; h->ws_cap = 0;
        JMC_CUDA(cudaMalloc((void**)&h->d_ws, L.total));
        h->ws_cap = L.total;
        JMC_CUDA(cudaMemset(h->d_ws, 0, 256));              // work counter / finished blocks of the fused kernels: 0
        JMC_CUDA(cudaMemset(h->d_ws, 0x7f, sizeof(int)));   // poison flag: "none"; the kernels re-arm all three
        h->has_job = false;
    }
    if (h->pin_cap < L.total) {
        if (h->h_pin) JMC_CUDA(cudaFreeHost(h->h_pin));
        h->h_pin = nullptr; h->pin_cap = 0;
        JMC_CUDA(cudaMallocHost((void**)&h->h_pin, L.total));
        h->pin_cap = L.total;
    }
    return JMC_OK;
}

// Make (g, edges, precision) the handle's current job: derive the FP32 plan and stage both edge arrays on the device,
// asynchronously on `st`.  A job identical to the cached one costs two memcmp.
int handle_set_job(jmc_handle* h, const jmc_integrand* g, int precision, const double* h_edges, int n_edges,
                   cudaStream_t st) {
    JMC_REQUIRE(h != nullptr && g != nullptr, "jmc_handle: NULL argument");
    int dev = -1;
    JMC_CUDA(cudaGetDevice(&dev));
    JMC_REQUIRE(dev == h->device, "jmc_handle: created on device %d, used on device %d", h->device, dev);
    if (!h_edges) n_edges = 0;
    if (h->has_job && h->precision == precision && (int)h->edges.size() == n_edges
        && std::memcmp(&h->g, g, sizeof(*g)) == 0
        && (n_edges == 0 || std::memcmp(h->edges.data(), h_edges, sizeof(double) * n_edges) == 0))
        return JMC_OK;
    h->has_job = false;
    if (h->plan) { fast_plan_free(h->plan); h->plan = nullptr; }
    int rc = handle_reserve(h, n_edges);
    if (rc) return rc;
    const bool fast = precision == JMC_F32 && g->kind != JMC_KIND_SHOWER;
    if (fast) {
        rc = fast_plan_make(g, h_edges, n_edges, &h->plan);
        if (rc) return rc;
    }
    if (n_edges > 0) {
        const WsLayout L(n_edges);
        JMC_CUDA(cudaEventSynchronize(h->staged));          // an earlier upload may still be reading the staging
        std::memcpy(h->h_pin + L.off_e64, h_edges, sizeof(double) * n_edges);
        const size_t bytes = fast ? L.off_e32 + sizeof(float) * n_edges - L.off_e64 : sizeof(double) * n_edges;
        if (fast) std::memcpy(h->h_pin + L.off_e32, fast_plan_edges(h->plan), sizeof(float) * n_edges);
        JMC_CUDA(cudaMemcpyAsync(h->d_ws + L.off_e64, h->h_pin + L.off_e64, bytes, cudaMemcpyHostToDevice, st));
        JMC_CUDA(cudaEventRecord(h->staged, st));
    }
    h->g = *g;
    h->precision = precision;
    h->edges.assign(h_edges, h_edges + n_edges);
    h->ws_edges = n_edges;
    h->has_job = true;
    return JMC_OK;
}

int handle_run(jmc_handle* h, const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index, int precision,
               double* d_sum_w, double* d_sum_w2, uint64_t* d_count, double* d_minmax, cudaStream_t st,
               bool defer_poison = false) {
    const WsLayout L(h->ws_edges);
    const int ne = h->ws_edges;
    const double* d_e64 = ne ? (const double*)(h->d_ws + L.off_e64) : nullptr;
    if (n == 0) return JMC_OK;
    if (h->plan) {
        // FP32 kernels take at most 2^38 samples per launch (32-bit per-block bin counts)
        const uint64_t kMaxPerLaunch = 1ull << 38;
        while (n) {
            const uint64_t chunk = n < kMaxPerLaunch ? n : kMaxPerLaunch;
            int rc = fast_plan_run(h->plan, ne ? (const float*)(h->d_ws + L.off_e32) : nullptr, (int*)h->d_ws, chunk,
                                   seed, first_index, d_sum_w, d_sum_w2, d_count, d_minmax, (void*)st, defer_poison);
            if (rc) return rc;
            n -= chunk;
            first_index += chunk;
        }
        return JMC_OK;
    }
    return jmc_run(g, n, seed, first_index, precision, d_e64, ne, d_sum_w, d_sum_w2, d_count, d_minmax, (void*)st);
}

}  // namespace

extern "C" int jmc_create(jmc_handle** out) {
    JMC_REQUIRE(out != nullptr, "jmc_create: NULL argument");
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc) return rc;
    jmc_handle* h = new jmc_handle;
    h->device = di.device;
    cudaError_t e = cudaEventCreateWithFlags(&h->staged, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "jmc_create"); }
    *out = h;
    return JMC_OK;
}

extern "C" int jmc_destroy(jmc_handle* h) {
    if (!h) return JMC_OK;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->plan) fast_plan_free(h->plan);
    if (h->d_ws) cudaFree(h->d_ws);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->staged) cudaEventDestroy(h->staged);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return JMC_OK;
}

extern "C" int jmc_handle_run(jmc_handle* h, const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index,
                              int precision, const double* h_edges, int n_edges, double* d_sum_w, double* d_sum_w2,
                              uint64_t* d_count, double* d_minmax, void* stream) {
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_handle_run: precision must be JMC_F64 or JMC_F32");
    if (h_edges) {
        JMC_REQUIRE(n_edges >= 2, "jmc_handle_run: need at least 2 edges");
        JMC_REQUIRE(d_sum_w && d_sum_w2 && d_count, "jmc_handle_run: histogram outputs missing");
    } else {
        JMC_REQUIRE(d_minmax != nullptr, "jmc_handle_run: neither edges nor minmax requested");
    }
    int rc = handle_set_job(h, g, precision, h_edges, n_edges, (cudaStream_t)stream);
    if (rc) return rc;
    return handle_run(h, g, n, seed, first_index, precision, d_sum_w, d_sum_w2, d_count, d_minmax, (cudaStream_t)stream);
}

extern "C" int jmc_handle_integrate_host(jmc_handle* h, const jmc_integrand* g, uint64_t n, uint64_t seed,
                                         uint64_t first_index, int precision, const double* h_edges, int n_edges,
                                         int bc_kind, double bc_value, double* h_density, double* h_density_err,
                                         double* h_integral, double* h_integral_err, double* h_sum_w,
                                         double* h_sum_w2, uint64_t* h_count) {
    JMC_REQUIRE(n_edges >= 2 && h_edges, "jmc_integrate_host: Need bins to produce density histogram.");
    JMC_REQUIRE(precision == JMC_F64 || precision == JMC_F32, "jmc_integrate_host: precision must be JMC_F64 or JMC_F32");
    JMC_REQUIRE(h != nullptr, "jmc_integrate_host: handle is NULL");
    double area;
    int rc = jmc_integrand_area(g, &area);
    if (rc) return rc;
    cudaStream_t st = h->stream;
    rc = handle_set_job(h, g, precision, h_edges, n_edges, st);
    if (rc) return rc;
    const WsLayout L(n_edges);
    const size_t nb = (size_t)n_edges - 1;
    double* d_sum_w = (double*)(h->d_ws + L.off_hist);
    double* d_sum_w2 = d_sum_w + nb;
    uint64_t* d_count = (uint64_t*)(d_sum_w2 + nb);
    double* d_density = (double*)(h->d_ws + L.off_out);
    double* d_err = d_density + nb;
    double* d_ierr = d_err + nb;
    double* d_integral = d_ierr + nb;
    JMC_CUDA(cudaMemsetAsync(d_sum_w, 0, sizeof(double) * 3 * nb, st));
    // FP32 plans leave their NaN flag in the handle's scratch: the finalize kernel applies it (no kernel of its own)
    const bool deferred = h->plan != nullptr;
    rc = handle_run(h, g, n, seed, first_index, precision, d_sum_w, d_sum_w2, d_count, nullptr, st, deferred);
    if (rc) return rc;
    JMC_REQUIRE(n > 0, "jmc_integrate_host: zero samples (the reference divides by len(observables) here)");
    rc = deferred ? finalize_poisoned(d_sum_w, d_sum_w2, d_count, (const double*)(h->d_ws + L.off_e64), n_edges, area, n,
                                      bc_kind, bc_value, d_density, d_err, d_integral, d_ierr,
                                      g->nan_to_num ? nullptr : (int*)h->d_ws, (void*)st)
                  : jmc_finalize(d_sum_w, d_sum_w2, d_count, (const double*)(h->d_ws + L.off_e64), n_edges, area, n,
                                 bc_kind, bc_value, d_density, d_err, d_integral, d_ierr, (void*)st);
    if (rc) return rc;
    // histograms and results are contiguous: one copy into pinned memory, one synchronisation
    const size_t bytes = sizeof(double) * (7 * nb + 1);
    char* stage = h->h_pin + L.off_hist;
    JMC_CUDA(cudaMemcpyAsync(stage, d_sum_w, bytes, cudaMemcpyDeviceToHost, st));
    JMC_CUDA(cudaStreamSynchronize(st));
    const double* r = (const double*)stage;
    if (h_sum_w) std::memcpy(h_sum_w, r, sizeof(double) * nb);
    if (h_sum_w2) std::memcpy(h_sum_w2, r + nb, sizeof(double) * nb);
    if (h_count) std::memcpy(h_count, r + 2 * nb, sizeof(uint64_t) * nb);
    if (h_density) std::memcpy(h_density, r + 3 * nb, sizeof(double) * nb);
    if (h_density_err) std::memcpy(h_density_err, r + 4 * nb, sizeof(double) * nb);
    if (h_integral_err) std::memcpy(h_integral_err, r + 5 * nb, sizeof(double) * nb);
    if (h_integral) std::memcpy(h_integral, r + 6 * nb, sizeof(double) * (nb + 1));
    return JMC_OK;
}

// the handle-less form keeps one handle per (host thread, device)
extern "C" int jmc_integrate_host(const jmc_integrand* g, uint64_t n, uint64_t seed, uint64_t first_index,
                                  int precision, const double* h_edges, int n_edges, int bc_kind, double bc_value,
                                  double* h_density, double* h_density_err, double* h_integral,
                                  double* h_integral_err, double* h_sum_w, double* h_sum_w2, uint64_t* h_count) {
    // argument validation comes before any CUDA call
    JMC_REQUIRE(n_edges >= 2 && h_edges, "jmc_integrate_host: Need bins to produce density histogram.");
    JMC_REQUIRE(g != nullptr, "jmc_integrate_host: integrand is NULL");
    static thread_local std::vector<jmc_handle*> handles;
    int dev = -1;
    JMC_CUDA(cudaGetDevice(&dev));
    jmc_handle* h = nullptr;
    for (jmc_handle* c : handles) if (c->device == dev) h = c;
    if (!h) {
        int rc = jmc_create(&h);
        if (rc) return rc;
        handles.push_back(h);
    }
    return jmc_handle_integrate_host(h, g, n, seed, first_index, precision, h_edges, n_edges, bc_kind, bc_value,
                                     h_density, h_density_err, h_integral, h_integral_err, h_sum_w, h_sum_w2, h_count);
}

extern "C" int jmc_sample_host(const jmc_sampler* s, uint64_t n, uint64_t seed, uint64_t first_index,
                               double* h_samples, double* h_jac) {
    JMC_REQUIRE(s != nullptr, "jmc_sample_host: sampler is NULL");
    const int dim = (s->kind == JMC_SAMPLER_CRITICAL || s->kind == JMC_SAMPLER_UNGROOMED) ? 2 : 1;
    DevBuf buf;
    int rc = buf.alloc(sizeof(double) * n * (dim + 1));
    if (rc) return rc;
    double* d_samples = buf.as<double>();
    double* d_jac = d_samples + n * dim;
    rc = jmc_sample(s, n, seed, first_index, d_samples, d_jac, nullptr);
    if (rc) return rc;
    if (n && h_samples) JMC_CUDA(cudaMemcpy(h_samples, d_samples, sizeof(double) * n * dim, cudaMemcpyDeviceToHost));
    if (n && h_jac) JMC_CUDA(cudaMemcpy(h_jac, d_jac, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return JMC_OK;
}

extern "C" int jmc_eval_fn_host(int fn, const jmc_fn_params* p, uint64_t n, const double* h_a0, int64_t stride0,
                                const double* h_a1, int64_t stride1, const double* h_a2, int64_t stride2,
                                const double* h_a3, int64_t stride3, double* h_out) {
    JMC_REQUIRE(n == 0 || h_out, "jmc_eval_fn_host: out is NULL");
    if (n == 0) return JMC_OK;
    const double* h[4] = {h_a0, h_a1, h_a2, h_a3};
    const int64_t st[4] = {stride0, stride1, stride2, stride3};
    DevBuf in[4], out;
    const double* d[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) {
        if (!h[k]) continue;
        JMC_REQUIRE(st[k] >= 1, "jmc_eval_fn_host: stride of argument %d must be >= 1", k);
        const size_t count = (size_t)(n - 1) * st[k] + 1;
        int rc = in[k].alloc(sizeof(double) * count);
        if (rc) return rc;
        JMC_CUDA(cudaMemcpy(in[k].p, h[k], sizeof(double) * count, cudaMemcpyHostToDevice));
        d[k] = in[k].as<double>();
    }
    int rc = out.alloc(sizeof(double) * n);
    if (rc) return rc;
    rc = jmc_eval_fn(fn, p, n, d[0], st[0], d[1], st[1], d[2], st[2], d[3], st[3], out.as<double>(), nullptr);
    if (rc) return rc;
    JMC_CUDA(cudaMemcpy(h_out, out.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return JMC_OK;
}
