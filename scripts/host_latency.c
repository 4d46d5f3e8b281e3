/* Wall time of a small job through the host-buffer entry point, from C (no interpreter in the loop):
 *   host_latency <libjmc_b200.so> <n> <reps> [kind: 0 = C1 radiator fc (default), 2 = C3 crit x sub rc counts on request]
 * prints microseconds per jmc_integrate_host call (host edges in, host results out). */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "jmc_b200.h"

typedef int (*integrate_fn)(const jmc_integrand*, uint64_t, uint64_t, uint64_t, int, const double*, int, int, double,
                            double*, double*, double*, double*, double*, double*, uint64_t*);

static double now_us(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: host_latency lib n reps [kind]\n"); return 2; }
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    integrate_fn integrate = (integrate_fn)dlsym(h, "jmc_integrate_host");
    const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "jmc_last_error");
    const uint64_t n = (uint64_t)atof(argv[2]);
    const int reps = atoi(argv[3]), kind = argc > 4 ? atoi(argv[4]) : 0;
    enum { NE = 100, NB = NE - 1 };
    double edges[NE], dens[NB], err[NB], integ[NE], ierr[NB], sw[NB], sw2[NB];
    uint64_t cnt[NB];
    for (int k = 0; k < NE; ++k) edges[k] = pow(10.0, -11.0 + k * (log10(0.5) + 11.0) / (NE - 1));
    jmc_integrand g;
    memset(&g, 0, sizeof(g));
    g.kind = kind == 2 ? JMC_KIND_CRIT_SUB : JMC_KIND_RADIATOR;
    g.method = JMC_LOG; g.jet = JMC_QUARK; g.fixed_coupling = kind == 2 ? 0 : 1;
    g.obs_acc = kind == 2 ? JMC_ACC_MLL : JMC_ACC_LL; g.split_acc = JMC_ACC_LL;
    g.nan_to_num = kind == 2; g.counts = kind == 2 ? JMC_COUNT_WEIGHTED : JMC_COUNT_ALL;
    g.epsilon = 1e-10; g.z_cut = 0.1; g.radius = 1.0; g.beta = 2.0; g.f_soft = 0.0; g.z_pre = 0.1; g.theta_scale = 1.0;
    const int bc = kind == 2 ? JMC_BC_LAST_MINUS : JMC_BC_LAST_PLUS;
    const double bcv = kind == 2 ? 1.0 : 0.0;
    for (int k = 0; k < 50; ++k)
        if (integrate(&g, n, 1, (uint64_t)k * n, JMC_F32, edges, NE, bc, bcv, dens, err, integ, ierr, sw, sw2, cnt)) {
            fprintf(stderr, "jmc_integrate_host: %s\n", last_error());
            return 1;
        }
    const double t0 = now_us();
    for (int k = 0; k < reps; ++k)
        integrate(&g, n, 1, (uint64_t)(50 + k) * n, JMC_F32, edges, NE, bc, bcv, dens, err, integ, ierr, sw, sw2, cnt);
    const double us = (now_us() - t0) / reps;
    printf("{\"n\": %llu, \"kind\": %d, \"reps\": %d, \"us_per_job\": %.3f, \"integral_first\": %.9g}\n",
           (unsigned long long)n, kind, reps, us, integ[0]);
    return 0;
}
