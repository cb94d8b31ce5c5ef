/*
 * c_driver.c -- replays BASELINE config C1 (SIS [100,1] on data/pooley.csv, theta = (0.003, 0.1)) through include/dpomp.h
 * with no Python and no Julia: what a host in any language does through the C ABI.  Compiled with gcc and run by
 * tests/test_gpu_cdriver.py on the GPU box.
 *
 *   get_particle_filter_lpdf(model, y)(theta)   src/hmm_utils.jl:281-284   -> dpomp_model_create / dpomp_pf_create / dpomp_pf_loglik
 *   partial_log_likelihood! call sequence       src/hmm_particle_filter.jl:39-76 -> dpomp_pf_partial (1..2 then 3..5)
 *   rs_systematic                               src/hmm_resample.jl:44-62  -> dpomp_resample_indices
 *   pop2[p] .= pop[nidx[p]]                     src/hmm_ibis.jl:74         -> dpomp_pf_permute / dpomp_pf_resample_migrate (world 1)
 *
 * Prints one line per check and exits 0 only if all pass.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dpomp.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != DPOMP_OK) {                                                             \
            fprintf(stderr, "FAIL %s -> %d: %s\n", #call, rc_, dpomp_last_error());        \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)
#define EXPECT(cond, ...)                                                                  \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            fprintf(stderr, "FAIL %s: ", #cond);                                           \
            fprintf(stderr, __VA_ARGS__);                                                  \
            fprintf(stderr, "\n");                                                         \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main(void) {
    /* generate_model("SIS", [100, 1]) (src/hmm_examples.jl:107-110, :181): rates theta1*S*I, theta2*I; rows [-1 1; 1 -1];
     * Gaussian observation model on compartment 2 (seq = 2:2), sigma = 2 (src/hmm_examples.jl:59-67) */
    dpomp_model_desc d;
    memset(&d, 0, sizeof(d));
    d.n_compartments = 2; d.n_events = 2; d.n_params = 2; d.t0_index = 0;
    for (int e = 0; e < DPOMP_MAX_EVENTS; ++e) d.rate_par[e] = -1;
    d.rate_par[0] = 0; d.rate_f1[0][0] = 1; d.rate_f2[0][1] = 1;               /* theta1 * S * I */
    d.rate_par[1] = 1; d.rate_f1[1][1] = 1; d.rate_k2[1] = 1;                  /* theta2 * I * 1 */
    d.trans[0][0] = -1; d.trans[0][1] = 1; d.trans[1][0] = 1; d.trans[1][1] = -1;
    d.initial_condition[0] = 100; d.initial_condition[1] = 1;
    d.obs_sigma = 2.0; d.obs_xmask[1] = 1; d.n_obs_vals = 2; d.obs_ymask[1] = 1;
    /* data/pooley.csv: t = 20..100, val = [0, I] */
    const double times[5] = {20, 40, 60, 80, 100};
    const int32_t ids[5] = {1, 1, 1, 1, 1};
    const int64_t vals[10] = {0, 18, 0, 65, 0, 70, 0, 66, 0, 67};
    d.n_obs = 5; d.obs_time = times; d.obs_id = ids; d.obs_val = vals;

    int ndev = 0;
    CHECK(dpomp_device_count(&ndev));
    EXPECT(ndev >= 1, "no CUDA device");
    EXPECT(dpomp_version() >= 100, "version");

    dpomp_model* model = NULL;
    CHECK(dpomp_model_create(&d, &model));

    /* 1. estimate_likelihood, 16 replicate filters of 4096 particles: anchor -15.69 +- 0.01, sd of one
     *    estimate ~0.06 -> the mean of 16 lies within 0.1 */
    enum { NB = 16, NP = 4096 };
    dpomp_pf* pf = NULL;
    CHECK(dpomp_pf_create(model, NP, NB, DPOMP_RS_SYSTEMATIC, 2026, -1, &pf));
    double theta[2 * NB], ll[NB], mean = 0.0;
    for (int b = 0; b < NB; ++b) { theta[2 * b] = 0.003; theta[2 * b + 1] = 0.1; }
    CHECK(dpomp_pf_loglik(pf, theta, NB, ll));
    for (int b = 0; b < NB; ++b) mean += ll[b] / NB;
    EXPECT(fabs(mean + 15.69) < 0.1, "mean log-likelihood %.4f", mean);
    printf("ok loglik: mean of %d filters = %.4f (anchor -15.69)\n", NB, mean);

    /* 2. determinism: the same stream key reproduces the estimate bit for bit; partial calls compose (1..2 then 3..5 with
     *    device-resident populations) into a finite value with the same law */
    double a[NB], b2[NB], g1[NB], g2[NB];
    CHECK(dpomp_pf_set_stream_key(pf, 77)); CHECK(dpomp_pf_loglik(pf, theta, NB, a));
    CHECK(dpomp_pf_set_stream_key(pf, 77)); CHECK(dpomp_pf_loglik(pf, theta, NB, b2));
    EXPECT(memcmp(a, b2, sizeof(a)) == 0, "same key, different result");
    CHECK(dpomp_pf_partial(pf, theta, NB, 1, 2, g1));
    CHECK(dpomp_pf_partial(pf, theta, NB, 3, 5, g2));
    double mean2 = 0.0;
    for (int b = 0; b < NB; ++b) mean2 += (g1[b] + g2[b]) / NB;
    EXPECT(fabs(mean2 + 15.69) < 0.1, "composed partial calls: %.4f", mean2);
    printf("ok determinism + partial composition: %.4f\n", mean2);

    /* 3. populations: S + I = 101 for every particle; permute / resample_migrate (world-size-1 communicator) gather whole filters */
    int64_t* pop = (int64_t*)malloc(sizeof(int64_t) * NP * 2);
    int64_t* pop3 = (int64_t*)malloc(sizeof(int64_t) * NP * 2);
    CHECK(dpomp_pf_get_pop(pf, 3, pop3));
    for (int p = 0; p < NP; ++p) EXPECT(pop3[p] + pop3[NP + p] == 101 && pop3[p] >= 0 && pop3[NP + p] >= 0, "particle %d: %lld + %lld", p, (long long)pop3[p], (long long)pop3[NP + p]);
    int64_t nidx[NB];
    for (int b = 0; b < NB; ++b) nidx[b] = 3;
    CHECK(dpomp_pf_permute(pf, nidx, NB));
    CHECK(dpomp_pf_get_pop(pf, 9, pop));
    EXPECT(memcmp(pop, pop3, sizeof(int64_t) * NP * 2) == 0, "permute");
    dpomp_comm* comm = NULL;
    CHECK(dpomp_comm_create(NULL, 0, 0, 1, -1, &comm));
    int32_t rank = -1, world = -1;
    CHECK(dpomp_comm_info(comm, &rank, &world));
    EXPECT(rank == 0 && world == 1, "comm info");
    for (int b = 0; b < NB; ++b) nidx[b] = NB - b;
    CHECK(dpomp_pf_resample_migrate(pf, comm, nidx, NB));
    CHECK(dpomp_pf_get_pop(pf, 1, pop));
    EXPECT(memcmp(pop, pop3, sizeof(int64_t) * NP * 2) == 0, "resample_migrate");
    double gall[NB];
    CHECK(dpomp_pf_partial_allgather(pf, comm, theta, NB, 1, 5, NB, gall));
    for (int b = 0; b < NB; ++b) EXPECT(isfinite(gall[b]) && gall[b] < -10 && gall[b] > -25, "partial_allgather[%d] = %g", b, gall[b]);
    CHECK(dpomp_comm_destroy(comm));
    printf("ok populations, permute, world-size-1 communicator\n");

    /* 4. rs_systematic known answers (hand cases of tests/test_oracle.py): w = [1,1,1,1], r = 0.5 -> 1,2,3,4;
     *    w = [0,0,1,0] -> all 3; w = [3,1], r = 0.9 -> 1,2 */
    int64_t idx[4];
    const double w1[4] = {1, 1, 1, 1}, w2[4] = {0, 0, 1, 0}, w3[2] = {3, 1}, r5 = 0.5, r9 = 0.9;
    CHECK(dpomp_resample_indices(DPOMP_RS_SYSTEMATIC, 0, w1, 4, &r5, 1, 4, idx, -1));
    EXPECT(idx[0] == 1 && idx[1] == 2 && idx[2] == 3 && idx[3] == 4, "rs_systematic uniform weights");
    CHECK(dpomp_resample_indices(DPOMP_RS_SYSTEMATIC, 0, w2, 4, &r5, 1, 4, idx, -1));
    EXPECT(idx[0] == 3 && idx[1] == 3 && idx[2] == 3 && idx[3] == 3, "rs_systematic single weight");
    CHECK(dpomp_resample_indices(DPOMP_RS_SYSTEMATIC, 0, w3, 2, &r9, 1, 2, idx, -1));
    EXPECT(idx[0] == 1 && idx[1] == 2, "rs_systematic [3,1] r = 0.9 -> %lld %lld", (long long)idx[0], (long long)idx[1]);
    printf("ok rs_systematic known answers\n");

    /* 5. errors come back as codes + messages, never as crashes */
    EXPECT(dpomp_pf_partial(pf, theta, NB, 0, 2, g1) == DPOMP_ERR_ARG, "ymin = 0 must be rejected");
    EXPECT(strlen(dpomp_last_error()) > 0, "error message");
    int64_t ovf = -1;
    CHECK(dpomp_pf_overflow_count(pf, &ovf));
    EXPECT(ovf == 0, "event cap hit %lld times", (long long)ovf);
    printf("ok error handling\n");

    free(pop); free(pop3);
    CHECK(dpomp_pf_destroy(pf));
    CHECK(dpomp_model_destroy(model));
    printf("C DRIVER OK\n");
    return 0;
}
