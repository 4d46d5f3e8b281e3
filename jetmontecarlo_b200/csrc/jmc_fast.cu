// placeholder until the FP32 path lands
#include "jmc_common.cuh"
namespace jmc {
int run_fast(const jmc_integrand*, uint64_t, uint64_t, uint64_t, const double*, int, double*, double*, uint64_t*, double*, void*) {
    return set_error(JMC_EINVAL, "jmc_run: FP32 mode not built");
}
}
