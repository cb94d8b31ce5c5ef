/*
 * dpomp_oracle_ibis.c -- CPU restatement of the outer layers that call the particle filter:
 *   run_pibis   (SMC^2, Chopin et al. 2013)            src/hmm_ibis.jl:12-135
 *   pMCMC spec  (run_pmcmc + commented generic_mcmc!)  src/hmm_mcmc.jl:349-365, 166-211
 * TEST INFRASTRUCTURE ONLY (see dpomp_oracle.c).  "parity unpinned": the only reference numbers for these layers are
 * the seeded single-run statistics of test/runtests.jl:35,51 (Julia RNG stream, not reproducible draw for draw).
 *
 * Host randomness (theta proposals, accept/reject, outer resampling) comes from xoshiro256++ seeded by the caller;
 * particle-filter randomness from Philox keys derived from the same seed.  Priors are products of uniforms
 * (Distributions.Product(Uniform.(lower, upper)), src/hmm_examples.jl:33-35, test/runtests.jl:29).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/dpomp.h"

int orc_pf_partial(const dpomp_model_desc* m, const double* theta, int64_t n, int64_t* pop, int ymin, int ymax,
                   int rs_type, uint64_t key, uint32_t filter, int mode, int tile, int items, int64_t max_events,
                   int threads, double* out_ll, double* logw_last, int64_t* anc_last, int64_t* n_events,
                   int64_t* n_overflow);
void orc_search_systematic(const double* cw, int64_t n, double r, int64_t* out);
double orc_compute_ess(const double* w, int64_t n);
void orc_compute_is_mu_covar(double* mu, double* cv, const double* theta, const double* w, int n_theta, int64_t n);

/* ---- host RNG ------------------------------------------------------------------------------------------------- */
typedef struct { uint64_t s[4]; int has_spare; double spare; } orc_rng;
static uint64_t sm64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(orc_rng* r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) r->s[i] = sm64(&seed);
    r->has_spare = 0;
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(orc_rng* r) {
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_uniform(orc_rng* r) { return (double)(rng_next(r) >> 11) * 0x1.0p-53; } /* [0,1) like rand() */
static double rng_normal(orc_rng* r) {
    if (r->has_spare) { r->has_spare = 0; return r->spare; }
    double u, v, s;
    do { u = 2.0 * rng_uniform(r) - 1.0; v = 2.0 * rng_uniform(r) - 1.0; s = u * u + v * v; } while (s >= 1.0 || s == 0.0);
    const double f = sqrt(-2.0 * log(s) / s);
    r->spare = v * f; r->has_spare = 1;
    return u * f;
}

/* logpdf of Product(Uniform.(lo, hi)) */
static double prior_logpdf(const double* lo, const double* hi, const double* th, int d) {
    double lp = 0.0;
    for (int i = 0; i < d; ++i) {
        if (th[i] < lo[i] || th[i] > hi[i]) return -INFINITY;
        lp -= log(hi[i] - lo[i]);
    }
    return lp;
}
/* lower Cholesky factor of a d x d row-major matrix; returns 0 if not positive definite
 * (get_prop_density: isposdef(Hermitian(cv)) ? MvNormal(cv) : old, src/hmm_cmn.jl:33-42) */
static int cholesky(const double* a, int d, double* l) {
    memset(l, 0, sizeof(double) * (size_t)(d * d));
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = a[i * d + j];
            for (int k = 0; k < j; ++k) s -= l[i * d + k] * l[j * d + k];
            if (i == j) {
                if (!(s > 0.0)) return 0;
                l[i * d + i] = sqrt(s);
            } else l[i * d + j] = s / l[j * d + j];
        }
    return 1;
}
/* get_mv_param (src/hmm_cmn.jl:13-18): theta_i + sclr * rand(MvNormal(L L')) */
static void mv_param(orc_rng* r, const double* l, int d, double sclr, const double* theta_i, double* out) {
    double z[DPOMP_MAX_PARAMS];
    for (int i = 0; i < d; ++i) z[i] = rng_normal(r);
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s += l[i * d + k] * z[k];
        out[i] = theta_i[i] + sclr * s;
    }
}

/*
 * run_pibis (src/hmm_ibis.jl:12-135).  theta: n_theta x outer_p column-major, overwritten with the final sample.
 * Outputs: mu[n_theta], cv[n_theta^2], w[outer_p], bme[2] (already negated like ImportanceSample.bme), k_log[2].
 * `threads` > 1 evaluates the independent particle filters of one sweep with OpenMP (a pure re-ordering of work;
 * every random draw is made in the reference's order when ind_prop is true, the default of run_ibis_analysis;
 * with ind_prop false the scale tj is frozen within a sweep unless threads == 1, where the sweep is fully literal).
 */
int orc_run_pibis(const dpomp_model_desc* m, double* theta, int64_t outer_p, const double* prior_lo,
                  const double* prior_hi, double ess_rs_crit, int ind_prop, double alpha, int64_t npf, int n_props,
                  uint64_t seed, int threads, int64_t max_events, double* mu, double* cv, double* w, double* bme,
                  int64_t* k_log, int64_t* pf_steps) {
    const int d = m->n_params, C = m->n_compartments, T = m->n_obs;
    const size_t popsz = (size_t)npf * C;
    orc_rng rng; rng_seed(&rng, seed);
    uint64_t key_ctr = seed ^ 0xC0FFEEull;
    const double ess_crit = ess_rs_crit * (double)outer_p;                       /* :17 */
    double* aw = (double*)malloc(sizeof(double) * outer_p), *aw2 = (double*)malloc(sizeof(double) * outer_p);
    double* gx = (double*)calloc(outer_p, sizeof(double)), *mtd = (double*)calloc(outer_p, sizeof(double));
    double* theta2 = (double*)malloc(sizeof(double) * d * outer_p);
    double* theta_f = (double*)malloc(sizeof(double) * d * outer_p);
    double* prtf = (double*)malloc(sizeof(double) * outer_p), *uacc = (double*)malloc(sizeof(double) * outer_p);
    double* awf = (double*)malloc(sizeof(double) * outer_p), *gxf = (double*)malloc(sizeof(double) * outer_p);
    int64_t* pop = (int64_t*)calloc(popsz * outer_p, sizeof(int64_t));
    int64_t* pop2 = (int64_t*)calloc(popsz * outer_p, sizeof(int64_t));
    int64_t* popf = (int64_t*)calloc(popsz * outer_p, sizeof(int64_t));
    int64_t* nidx = (int64_t*)malloc(sizeof(int64_t) * outer_p);
    double* cwv = (double*)malloc(sizeof(double) * outer_p);
    double propd[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS], chol_tmp[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS];
    memset(propd, 0, sizeof(propd));
    for (int i = 0; i < d; ++i) propd[i * d + i] = 1.0;                          /* MvNormal(I) :45 */
    double tj = 0.2;                                                             /* :46 */
    for (int64_t i = 0; i < outer_p; ++i) { w[i] = 1.0; aw[i] = prior_logpdf(prior_lo, prior_hi, theta + i * d, d); }
    bme[0] = bme[1] = 0.0; k_log[0] = k_log[1] = 0;
    int64_t steps = 0;
    int obs_min = 1;
    for (int oi = 1; oi <= T; ++oi) {                                            /* :50 */
        if (m->obs_id[oi - 1] <= 0) continue;
        const uint64_t key = sm64(&key_ctr);
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) if (threads > 1)
        for (int64_t p = 0; p < outer_p; ++p)                                    /* :53-56 */
            orc_pf_partial(m, theta + p * d, npf, pop + p * popsz, obs_min, oi, DPOMP_RS_SYSTEMATIC, key, (uint32_t)p, 0,
                           1024, 4, max_events, 1, &gx[p], NULL, NULL, NULL, NULL);
        steps += outer_p * npf * (oi - obs_min + 1);
        double swg = 0.0, sw = 0.0;
        for (int64_t p = 0; p < outer_p; ++p) { aw[p] += gx[p]; gx[p] = exp(gx[p]); swg += w[p] * gx[p]; sw += w[p]; }
        const double lml = log(swg / sw);                                        /* :60 */
        bme[0] += lml;
        for (int64_t p = 0; p < outer_p; ++p) w[p] *= gx[p];
        orc_compute_is_mu_covar(mu, cv, theta, w, d, outer_p);                   /* :63 */
        if (orc_compute_ess(w, outer_p) < ess_crit) {                            /* :65 */
            if (cholesky(cv, d, chol_tmp)) memcpy(propd, chol_tmp, sizeof(double) * d * d);   /* :68 */
            cwv[0] = w[0];
            for (int64_t p = 1; p < outer_p; ++p) cwv[p] = cwv[p - 1] + w[p];
            orc_search_systematic(cwv, outer_p, rng_uniform(&rng), nidx);        /* rs_systematic(w) :70 */
            double gmean = 0.0;
            for (int64_t p = 0; p < outer_p; ++p) {                              /* :71-75 */
                const int64_t a = nidx[p] - 1;
                memcpy(theta2 + p * d, theta + a * d, sizeof(double) * d);
                aw2[p] = aw[a];
                memcpy(pop2 + p * popsz, pop + a * popsz, sizeof(int64_t) * popsz);
                mtd[p] = gx[a];
                gmean += gx[a];
            }
            const double mlr = gmean / (double)outer_p * exp(lml);                /* :76 */
            { double* t = theta; (void)t; memcpy(theta, theta2, sizeof(double) * d * outer_p); }
            memcpy(aw, aw2, sizeof(double) * outer_p);
            { int64_t* t = pop; pop = pop2; pop2 = t; }
            k_log[0] += outer_p;
            for (int mk = 0; mk < n_props; ++mk) {                               /* :83-116, sweep over p */
                const int literal = (threads == 1);
                if (literal) {
                    for (int64_t p = 0; p < outer_p; ++p) {
                        double* tf = theta_f + p * d;
                        if (ind_prop) mv_param(&rng, propd, d, 1.0, mu, tf); else mv_param(&rng, propd, d, tj, theta + p * d, tf);
                        const double pr = prior_logpdf(prior_lo, prior_hi, tf, d);
                        if (pr == -INFINITY) continue;                             /* :89 */
                        const uint64_t kf = sm64(&key_ctr);
                        double a1 = 0.0, g1 = 0.0;
                        int64_t* pf_ = popf;
                        if (oi == 1) {
                            orc_pf_partial(m, tf, npf, pf_, 1, 1, DPOMP_RS_SYSTEMATIC, kf, (uint32_t)p, 0, 1024, 4, max_events, 1, &g1, NULL, NULL, NULL, NULL);
                            a1 = g1;
                        } else {
                            orc_pf_partial(m, tf, npf, pf_, 1, oi - 1, DPOMP_RS_SYSTEMATIC, kf, (uint32_t)p, 0, 1024, 4, max_events, 1, &a1, NULL, NULL, NULL, NULL);
                            orc_pf_partial(m, tf, npf, pf_, oi, oi, DPOMP_RS_SYSTEMATIC, kf ^ 1, (uint32_t)p, 0, 1024, 4, max_events, 1, &g1, NULL, NULL, NULL, NULL);
                            a1 += g1;
                        }
                        steps += npf * oi;
                        a1 += pr;
                        if (exp(a1 - aw[p]) > rng_uniform(&rng)) {                 /* :104 */
                            mtd[p] = exp(g1);
                            memcpy(theta + p * d, tf, sizeof(double) * d);
                            aw[p] = a1;
                            memcpy(pop + p * popsz, pf_, sizeof(int64_t) * popsz);
                            k_log[1] += 1;
                            tj *= alpha;
                        } else tj *= 0.999;
                    }
                } else {
                    for (int64_t p = 0; p < outer_p; ++p) {
                        double* tf = theta_f + p * d;
                        if (ind_prop) mv_param(&rng, propd, d, 1.0, mu, tf); else mv_param(&rng, propd, d, tj, theta + p * d, tf);
                        prtf[p] = prior_logpdf(prior_lo, prior_hi, tf, d);
                        uacc[p] = (prtf[p] == -INFINITY) ? 2.0 : rng_uniform(&rng);
                    }
                    const uint64_t kf = sm64(&key_ctr);
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
                    for (int64_t p = 0; p < outer_p; ++p) {
                        if (prtf[p] == -INFINITY) continue;
                        double a1 = 0.0, g1 = 0.0;
                        int64_t* pf_ = popf + p * popsz;
                        if (oi == 1) {
                            orc_pf_partial(m, theta_f + p * d, npf, pf_, 1, 1, DPOMP_RS_SYSTEMATIC, kf, (uint32_t)p, 0, 1024, 4, max_events, 1, &g1, NULL, NULL, NULL, NULL);
                            a1 = g1;
                        } else {
                            orc_pf_partial(m, theta_f + p * d, npf, pf_, 1, oi - 1, DPOMP_RS_SYSTEMATIC, kf, (uint32_t)p, 0, 1024, 4, max_events, 1, &a1, NULL, NULL, NULL, NULL);
                            orc_pf_partial(m, theta_f + p * d, npf, pf_, oi, oi, DPOMP_RS_SYSTEMATIC, kf ^ 1, (uint32_t)p, 0, 1024, 4, max_events, 1, &g1, NULL, NULL, NULL, NULL);
                            a1 += g1;
                        }
                        awf[p] = a1 + prtf[p];
                        gxf[p] = g1;
                    }
                    for (int64_t p = 0; p < outer_p; ++p) {
                        if (prtf[p] == -INFINITY) continue;
                        steps += npf * oi;
                        if (exp(awf[p] - aw[p]) > uacc[p]) {
                            mtd[p] = exp(gxf[p]);
                            memcpy(theta + p * d, theta_f + p * d, sizeof(double) * d);
                            aw[p] = awf[p];
                            memcpy(pop + p * popsz, popf + p * popsz, sizeof(int64_t) * popsz);
                            k_log[1] += 1;
                            tj *= alpha;
                        } else tj *= 0.999;
                    }
                }
            }
            double mmean = 0.0;
            for (int64_t p = 0; p < outer_p; ++p) mmean += mtd[p];
            bme[1] += log(mlr / (mmean / (double)outer_p));                      /* :118 */
            for (int64_t p = 0; p < outer_p; ++p) w[p] = 1.0;                    /* :119 */
        } else {
            double swg2 = 0.0, sw2 = 0.0;                                        /* :122 (uses the updated w, as written) */
            for (int64_t p = 0; p < outer_p; ++p) { swg2 += w[p] * gx[p]; sw2 += w[p]; }
            bme[1] += log(swg2 / sw2);
        }
        obs_min = oi + 1;                                                        /* :124 */
    }
    orc_compute_is_mu_covar(mu, cv, theta, w, d, outer_p);                       /* :128 */
    bme[0] = -bme[0]; bme[1] = -bme[1];                                          /* ImportanceSample(..., -bme) :132 */
    if (pf_steps) *pf_steps = steps;
    free(aw); free(aw2); free(gx); free(mtd); free(theta2); free(theta_f); free(prtf); free(uacc); free(awf); free(gxf);
    free(pop); free(pop2); free(popf); free(nidx); free(cwv);
    return 0;
}

/*
 * Particle MCMC as specified by run_pmcmc (src/hmm_mcmc.jl:349-365) and the commented generic_mcmc!
 * (src/hmm_mcmc.jl:166-211): adaptive random-walk Metropolis on theta, target = log prior + PF log-likelihood estimate
 * (estimate_likelihood with rsp_systematic); proposal MvNormal(covar) scaled by c, covar initialised to
 * diag(0.1 * theta0^2) (1 where theta0 == 0), c = C_INITIAL, c *= 1.002 on accept / 0.999 on reject while
 * i < adapt_period, covariance re-estimated from the chain every adapt_period / 10 steps.
 * samples: n_theta x steps x chains (Julia column-major: theta index fastest).  theta_init: n_theta x chains.
 */
int orc_run_pmcmc(const dpomp_model_desc* m, const double* theta_init, int n_chains, int steps, int adapt_period,
                  int64_t npf, const double* prior_lo, const double* prior_hi, double c_initial, uint64_t seed,
                  int threads, int64_t max_events, double* samples, int64_t* accepted) {
    const int d = m->n_params, T = m->n_obs;
    const double adapt_interval = (double)adapt_period / 10.0;   /* ADAPT_INTERVAL = adapt_period / 10, Float64 (:168) */
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (int mc = 0; mc < n_chains; ++mc) {
        orc_rng rng; rng_seed(&rng, seed + 0x9E37ull * (uint64_t)(mc + 1));
        uint64_t key_ctr = seed ^ (0xABCDull * (uint64_t)(mc + 1));
        int64_t* pop = (int64_t*)calloc((size_t)npf * m->n_compartments, sizeof(int64_t));
        double* chain = samples + (size_t)mc * d * steps;
        double covar[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS], l[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS], tmp[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS];
        memset(covar, 0, sizeof(covar));
        for (int i = 0; i < d; ++i) {
            const double t0 = theta_init[mc * d + i];
            covar[i * d + i] = 0.1 * (t0 == 0.0 ? 1.0 : t0 * t0);
        }
        cholesky(covar, d, l);
        double c = c_initial;
        memcpy(chain, theta_init + mc * d, sizeof(double) * d);
        double ll_i = prior_logpdf(prior_lo, prior_hi, chain, d);
        if (ll_i != -INFINITY) {
            double e = 0.0;
            orc_pf_partial(m, chain, npf, pop, 1, T, DPOMP_RS_SYSTEMATIC, sm64(&key_ctr), (uint32_t)mc, 0, 1024, 4, max_events, 1, &e, NULL, NULL, NULL, NULL);
            ll_i += e;
        }
        int64_t acc = 0;
        for (int i = 1; i < steps; ++i) {
            double* cur = chain + (size_t)i * d;
            const double* prv = chain + (size_t)(i - 1) * d;
            mv_param(&rng, l, d, c, prv, cur);
            double ll_f = prior_logpdf(prior_lo, prior_hi, cur, d);
            int ok = 0;
            if (ll_f != -INFINITY) {
                double e = 0.0;
                orc_pf_partial(m, cur, npf, pop, 1, T, DPOMP_RS_SYSTEMATIC, sm64(&key_ctr), (uint32_t)mc, 0, 1024, 4, max_events, 1, &e, NULL, NULL, NULL, NULL);
                ll_f += e;
                if (ll_f != -INFINITY) {
                    const double mh = exp(ll_f - ll_i);
                    ok = (mh > 1.0 || mh > rng_uniform(&rng));
                }
            }
            if (ok) { ll_i = ll_f; ++acc; } else memcpy(cur, prv, sizeof(double) * d);
            if (i + 1 < adapt_period) {   /* 1-based step index i+1 < adapt_period */
                c *= ok ? 1.002 : 0.999;
                if (fmod((double)(i + 1), adapt_interval) == 0.0) {   /* i % ADAPT_INTERVAL == 0 (:200) */
                    /* covar = cov(theta[:, 1:i, mc]) (sample covariance, n-1 denominator) */
                    const int n = i + 1;
                    double mean[DPOMP_MAX_PARAMS] = {0};
                    for (int s = 0; s < n; ++s) for (int a = 0; a < d; ++a) mean[a] += chain[(size_t)s * d + a];
                    for (int a = 0; a < d; ++a) mean[a] /= n;
                    double sum = 0.0;
                    for (int a = 0; a < d; ++a) for (int b2 = 0; b2 < d; ++b2) {
                        double v = 0.0;
                        for (int s = 0; s < n; ++s) v += (chain[(size_t)s * d + a] - mean[a]) * (chain[(size_t)s * d + b2] - mean[b2]);
                        tmp[a * d + b2] = v / (n - 1);
                        sum += tmp[a * d + b2];
                    }
                    if (sum != 0.0) {
                        double l2[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS];
                        if (cholesky(tmp, d, l2)) { memcpy(covar, tmp, sizeof(double) * d * d); memcpy(l, l2, sizeof(double) * d * d); }
                    }
                }
            }
        }
        if (accepted) accepted[mc] = acc;
        free(pop);
    }
    return 0;
}
