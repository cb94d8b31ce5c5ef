/*
 * dpomp_oracle_mbp.c -- CPU restatement of the MBP-IBIS layer (TEST INFRASTRUCTURE ONLY, see dpomp_oracle.c):
 *   iterate_particle!              src/hmm_sim.jl:6-25      (single-trajectory Gillespie that records events)
 *   iterate_mbp!                   src/hmm_mbp.jl:7-44      (model-based proposal walk, Pooley 2015)
 *   initialise_trajectory!         src/hmm_mbp.jl:47-80
 *   partial_model_based_proposal   src/hmm_mbp.jl:83-108
 *   run_mbp_ibis                   src/hmm_ibis.jl:140-244
 * "parity unpinned": the reference only prints the MBP-IBIS evidence (test/runtests.jl:55-59), no value is asserted.
 *
 * Draw streams (DESIGN.md "random streams"): (K, A, B) = Philox4x32-10(ctr = (0, id, obs, 2<<30 | which), call key);
 * draw j of particle `id` is Philox2x32-10(ctr = (id ^ A, j ^ B), key = K).  An event attempt uses (w0, w1) as
 * (waiting time, event type); a keep decision uses the 53-bit uniform of (w0, w1).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/dpomp.h"

void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]);
void orc_philox2x32_10(const uint32_t ctr_in[2], uint32_t key, uint32_t out[2]);
void orc_search_systematic(const double* cw, int64_t n, double r, int64_t* out);
void orc_search_stratified(const double* cw, int64_t n, const double* r, int64_t* out);
double orc_compute_ess(const double* w, int64_t n);
void orc_compute_is_mu_covar(double* mu, double* cv, const double* theta, const double* w, int n_theta, int64_t n);

#define TAG_MBP 2u
typedef struct { uint32_t k, a, b, id, j; } mbp_stream;
static mbp_stream mbp_stream_init(uint64_t key, uint32_t id, uint32_t obs, uint32_t which) {
    uint32_t ctr[4] = {0u, id, obs, (TAG_MBP << 30) | which}, k[2] = {(uint32_t)key, (uint32_t)(key >> 32)}, w[4];
    orc_philox4x32_10(ctr, k, w);
    mbp_stream s = {w[0], w[1], w[2], id, 0u};
    return s;
}
static void mbp_draw(mbp_stream* s, uint32_t out[2]) {
    uint32_t ctr[2] = {s->id ^ s->a, s->j ^ s->b};
    orc_philox2x32_10(ctr, s->k, out);
    s->j += 1;
}
static inline double u32o(uint32_t w) { return ((double)w + 0.5) * 0x1.0p-32; }
static inline double u53o(uint32_t hi, uint32_t lo) { return (double)(((uint64_t)hi << 21) | (uint64_t)(lo >> 11)) * 0x1.0p-53; }

static void rates(const dpomp_model_desc* m, const double* th, const int64_t* x, double* out) {
    for (int e = 0; e < m->n_events; ++e) {
        int64_t l1 = m->rate_k1[e], l2 = m->rate_k2[e], dn = m->rate_kd[e];
        for (int c = 0; c < m->n_compartments; ++c) {
            l1 += (int64_t)m->rate_f1[e][c] * x[c];
            l2 += (int64_t)m->rate_f2[e][c] * x[c];
            dn += (int64_t)m->rate_dn[e][c] * x[c];
        }
        double p = m->rate_par[e] >= 0 ? th[m->rate_par[e]] : 1.0;
        double r = (p * (double)l1) * (double)l2;
        if (m->rate_has_den[e]) r = (dn == 0) ? 0.0 : r / (double)dn;
        out[e] = r;
    }
}
static void cumsum_e(double* v, int n) { for (int i = 1; i < n; ++i) v[i] = v[i - 1] + v[i]; }
static int choose(const double* cum, int n, double u) {
    double etc = u * cum[n - 1];
    for (int i = 0; i < n - 1; ++i) if (cum[i] > etc) return i;
    return n - 1;
}
static double obs_ll(const dpomp_model_desc* m, int t, const int64_t* x) {
    double tmp1 = log(1.0 / (sqrt(2.0 * M_PI) * m->obs_sigma)), tmp2 = 2.0 * m->obs_sigma * m->obs_sigma;
    int64_t ys = 0, xs = 0;
    for (int v = 0; v < m->n_obs_vals; ++v) ys += (int64_t)m->obs_ymask[v] * m->obs_val[(int64_t)t * m->n_obs_vals + v];
    for (int c = 0; c < m->n_compartments; ++c) xs += (int64_t)m->obs_xmask[c] * x[c];
    int64_t d = ys - xs;
    return tmp1 - ((double)(d * d) / tmp2);
}

/* iterate_particle! (src/hmm_sim.jl:6-25).  Events are appended to (ev_time, ev_type)[*len]; `cap` plays MAX_TRAJ.
 * Returns the observation log-likelihood (or -Inf after a trajectory overflow, with log_like[0] = -Inf). */
double orc_mbp_iterate(const dpomp_model_desc* m, const double* theta, int64_t* fc, double* ev_time, int32_t* ev_type,
                       int64_t* len, int64_t cap, double* log_like, double time, int obs_i, uint64_t key, uint32_t id) {
    double cum[DPOMP_MAX_EVENTS];
    const int E = m->n_events, t = obs_i - 1;
    mbp_stream st = mbp_stream_init(key, id, (uint32_t)t, 0u);
    for (;;) {
        rates(m, theta, fc, cum); cumsum_e(cum, E);
        if (!(cum[E - 1] > 0.0)) break;                          /* :11 */
        uint32_t w[2]; mbp_draw(&st, w);
        time -= log(u32o(w[0])) / cum[E - 1];                    /* :12 */
        if (time > m->obs_time[t]) break;                        /* :13 */
        int et = choose(cum, E, u32o(w[1]));                     /* :14 */
        for (int c = 0; c < m->n_compartments; ++c) fc[c] += m->trans[et][c];   /* :15 */
        if (*len >= cap) { log_like[0] = -INFINITY; return -INFINITY; }        /* :17-20 (cap = MAX_TRAJ) */
        ev_time[*len] = time; ev_type[*len] = et + 1; *len += 1;               /* :16, 1-based type */
    }
    double out = obs_ll(m, t, fc);                               /* :22 */
    if (m->obs_id[t] > 0) log_like[0] += out;                    /* :23 */
    return out;
}

/* partial_model_based_proposal (src/hmm_mbp.jl:83-108) with iterate_mbp! (:7-44) and initialise_trajectory! (:47-80).
 * The caller has checked the prior of theta_f.  xf_loglike[2] must be zero on entry.  Returns 0, or 1 on overflow
 * (xf_loglike[0] = -Inf). */
int orc_mbp_propose(const dpomp_model_desc* m, const double* theta_i, const double* theta_f, const double* xi_time,
                    const int32_t* xi_type, int64_t xi_len, double* xf_time, int32_t* xf_type, int64_t* xf_len, int64_t cap,
                    int64_t* xf_fc, double* xf_loglike, int ymax, uint64_t key, uint32_t id) {
    const int E = m->n_events, C = m->n_compartments;
    int64_t pop_i[DPOMP_MAX_COMPARTMENTS];
    double lf[DPOMP_MAX_EVENTS], li[DPOMP_MAX_EVENTS], ld[DPOMP_MAX_EVENTS];
    mbp_stream st = mbp_stream_init(key, id, 0u, 1u);
    for (int c = 0; c < C; ++c) { xf_fc[c] = m->initial_condition[c]; pop_i[c] = m->initial_condition[c]; }
    *xf_len = 0;
    int64_t evt = 0;  /* 0-based index of the next old event */
    double time = 0.0;
#define PUSH(tt, ee) do { if (*xf_len >= cap) { xf_loglike[0] = -INFINITY; return 1; } \
                          xf_time[*xf_len] = (tt); xf_type[*xf_len] = (ee) + 1; *xf_len += 1; } while (0)
    if (m->t0_index > 0) {                                       /* initialise_trajectory! */
        const double t0f = theta_f[m->t0_index - 1], t0i = theta_i[m->t0_index - 1];
        if (t0f < t0i) {                                         /* 'sim' :53-67 */
            double t = t0f;
            for (;;) {
                rates(m, theta_f, xf_fc, lf); cumsum_e(lf, E);
                if (!(lf[E - 1] > 0.0)) break;
                uint32_t w[2]; mbp_draw(&st, w);
                t -= log(u32o(w[0])) / lf[E - 1];
                if (t > t0i) break;
                int et = choose(lf, E, u32o(w[1]));
                PUSH(t, et);
                for (int c = 0; c < C; ++c) xf_fc[c] += m->trans[et][c];
            }
        } else {                                                 /* 'delete' :69-76 */
            while (evt < xi_len && !(xi_time[evt] > t0f)) {
                for (int c = 0; c < C; ++c) pop_i[c] += m->trans[xi_type[evt] - 1][c];
                ++evt;
            }
        }
        time = t0f > t0i ? t0f : t0i;                            /* :94 */
    }
    for (int oi = 1; oi <= ymax; ++oi) {                         /* :95 */
        const double t_obs = m->obs_time[oi - 1];
        for (;;) {                                               /* iterate_mbp! :14-42 */
            const double tmax = (evt >= xi_len) ? t_obs : (t_obs < xi_time[evt] ? t_obs : xi_time[evt]);
            rates(m, theta_i, pop_i, li);
            for (;;) {
                rates(m, theta_f, xf_fc, lf);
                for (int e = 0; e < E; ++e) { double dlt = lf[e] - li[e]; ld[e] = dlt > 0.0 ? dlt : 0.0; }
                cumsum_e(ld, E);
                if (!(ld[E - 1] > 0.0)) break;
                uint32_t w[2]; mbp_draw(&st, w);
                time -= log(u32o(w[0])) / ld[E - 1];
                if (time > tmax) break;
                int et = choose(ld, E, u32o(w[1]));
                for (int c = 0; c < C; ++c) xf_fc[c] += m->trans[et][c];
                PUSH(time, et);
            }
            if (evt >= xi_len) break;
            if (xi_time[evt] > t_obs) break;
            const int et = xi_type[evt] - 1;
            time = xi_time[evt];
            const double prob_keep = lf[et] / li[et];
            int keep = prob_keep > 1.0;
            if (!keep) { uint32_t w[2]; mbp_draw(&st, w); keep = prob_keep > u53o(w[0], w[1]); }
            if (keep) { PUSH(time, et); for (int c = 0; c < C; ++c) xf_fc[c] += m->trans[et][c]; }
            for (int c = 0; c < C; ++c) pop_i[c] += m->trans[et][c];
            ++evt;
        }
        time = t_obs;                                            /* :102 */
        xf_loglike[1] = obs_ll(m, oi - 1, xf_fc);                /* :103 */
        if (m->obs_id[oi - 1] > 0) xf_loglike[0] += xf_loglike[1];
    }
#undef PUSH
    return 0;
}

/* ---- host RNG (same generator as dpomp_oracle_ibis.c) ---------------------------------------------------------- */
typedef struct { uint64_t s[4]; int has_spare; double spare; } mrng;
static uint64_t sm64(uint64_t* x) { uint64_t z = (*x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rnext(mrng* r) { uint64_t* s = r->s; const uint64_t res = rotl(s[0] + s[3], 23) + s[0]; const uint64_t t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return res; }
static double runif(mrng* r) { return (double)(rnext(r) >> 11) * 0x1.0p-53; }
static double rnorm(mrng* r) {
    if (r->has_spare) { r->has_spare = 0; return r->spare; }
    double u, v, s;
    do { u = 2.0 * runif(r) - 1.0; v = 2.0 * runif(r) - 1.0; s = u * u + v * v; } while (s >= 1.0 || s == 0.0);
    const double f = sqrt(-2.0 * log(s) / s);
    r->spare = v * f; r->has_spare = 1; return u * f;
}
static double prior_lp(const double* lo, const double* hi, const double* th, int d) {
    double lp = 0.0;
    for (int i = 0; i < d; ++i) { if (th[i] < lo[i] || th[i] > hi[i]) return -INFINITY; lp -= log(hi[i] - lo[i]); }
    return lp;
}
static int chol(const double* a, int d, double* l) {
    memset(l, 0, sizeof(double) * (size_t)(d * d));
    for (int i = 0; i < d; ++i) for (int j = 0; j <= i; ++j) {
        double s = a[i * d + j];
        for (int k = 0; k < j; ++k) s -= l[i * d + k] * l[j * d + k];
        if (i == j) { if (!(s > 0.0)) return 0; l[i * d + i] = sqrt(s); } else l[i * d + j] = s / l[j * d + j];
    }
    return 1;
}

/*
 * run_mbp_ibis (src/hmm_ibis.jl:140-244).  theta: n_theta x outer_p column-major (overwritten).  rs_type selects the
 * outer resampler (1 systematic as hard-coded in the reference :194, 2 stratified for BASELINE config C5).
 * `cap` = per-trajectory event capacity (MAX_TRAJ in the reference).
 */
int orc_run_mbp_ibis(const dpomp_model_desc* m, double* theta, int64_t outer_p, const double* prior_lo, const double* prior_hi,
                     double ess_rs_crit, int n_props, int ind_prop, double alpha, int rs_type, int64_t cap, uint64_t seed,
                     int threads, double* mu, double* cv, double* w, double* bme, int64_t* k_log) {
    const int d = m->n_params, C = m->n_compartments, T = m->n_obs;
    mrng rng; { uint64_t s = seed; for (int i = 0; i < 4; ++i) rng.s[i] = sm64(&s); rng.has_spare = 0; }
    uint64_t key_ctr = seed ^ 0xBADC0DEull;
    const double ess_crit = ess_rs_crit * (double)outer_p;
    /* particle stores: current, resample workspace, proposal */
    double *tm[3]; int32_t* ty[3]; int64_t* ln[3]; int64_t* fc[3]; double* ll[3]; double* pr[3]; double* th[3];
    for (int s = 0; s < 3; ++s) {
        tm[s] = (double*)malloc(sizeof(double) * outer_p * cap); ty[s] = (int32_t*)malloc(sizeof(int32_t) * outer_p * cap);
        ln[s] = (int64_t*)calloc(outer_p, sizeof(int64_t)); fc[s] = (int64_t*)calloc(outer_p * C, sizeof(int64_t));
        ll[s] = (double*)calloc(outer_p * 2, sizeof(double)); pr[s] = (double*)calloc(outer_p, sizeof(double));
        th[s] = (double*)malloc(sizeof(double) * outer_p * d);
    }
    memcpy(th[0], theta, sizeof(double) * outer_p * d);
    for (int64_t p = 0; p < outer_p; ++p) {                      /* :149-154 */
        for (int c = 0; c < C; ++c) fc[0][p * C + c] = m->initial_condition[c];
        pr[0][p] = prior_lp(prior_lo, prior_hi, th[0] + p * d, d);
    }
    double propd[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS], tmpl[DPOMP_MAX_PARAMS * DPOMP_MAX_PARAMS];
    memset(propd, 0, sizeof(propd));
    for (int i = 0; i < d; ++i) propd[i * d + i] = 1.0;
    double tj = 0.2;
    double* gx = (double*)calloc(outer_p, sizeof(double)); double* mtd = (double*)calloc(outer_p, sizeof(double));
    double* tcur = (double*)calloc(outer_p, sizeof(double)); int64_t* nidx = (int64_t*)malloc(sizeof(int64_t) * outer_p);
    double* cwv = (double*)malloc(sizeof(double) * outer_p); double* rr = (double*)malloc(sizeof(double) * outer_p);
    double* thf = (double*)malloc(sizeof(double) * outer_p * d); double* uacc = (double*)malloc(sizeof(double) * outer_p);
    for (int64_t p = 0; p < outer_p; ++p) { w[p] = 1.0; tcur[p] = m->t0_index > 0 ? th[0][p * d + m->t0_index - 1] : 0.0; }
    bme[0] = bme[1] = 0.0; k_log[0] = k_log[1] = 0;
    for (int oi = 1; oi <= T; ++oi) {
        const uint64_t key = sm64(&key_ctr);
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) if (threads > 1)
        for (int64_t p = 0; p < outer_p; ++p) {                  /* :176-179 */
            double g = orc_mbp_iterate(m, th[0] + p * d, fc[0] + p * C, tm[0] + p * cap, ty[0] + p * cap, &ln[0][p], cap,
                                       ll[0] + 2 * p, tcur[p], oi, key, (uint32_t)p);
            gx[p] = exp(g);
        }
        if (m->obs_id[oi - 1] > 0) {
            double swg = 0.0, sw = 0.0;
            for (int64_t p = 0; p < outer_p; ++p) { swg += w[p] * gx[p]; sw += w[p]; }
            const double lml = log(swg / sw);                    /* :181 */
            bme[0] += lml;
            for (int64_t p = 0; p < outer_p; ++p) w[p] *= gx[p];
            orc_compute_is_mu_covar(mu, cv, th[0], w, d, outer_p);
            if (orc_compute_ess(w, outer_p) < ess_crit) {        /* :190 */
                if (chol(cv, d, tmpl)) memcpy(propd, tmpl, sizeof(double) * d * d);
                cwv[0] = w[0];
                for (int64_t p = 1; p < outer_p; ++p) cwv[p] = cwv[p - 1] + w[p];
                if (rs_type == DPOMP_RS_STRATIFIED) { for (int64_t p = 0; p < outer_p; ++p) rr[p] = runif(&rng); orc_search_stratified(cwv, outer_p, rr, nidx); }
                else orc_search_systematic(cwv, outer_p, runif(&rng), nidx);
                double gmean = 0.0;
                for (int64_t p = 0; p < outer_p; ++p) {          /* :196-199 deepcopy */
                    const int64_t a = nidx[p] - 1;
                    mtd[p] = gx[a]; gmean += gx[a];
                    memcpy(tm[1] + p * cap, tm[0] + a * cap, sizeof(double) * ln[0][a]);
                    memcpy(ty[1] + p * cap, ty[0] + a * cap, sizeof(int32_t) * ln[0][a]);
                    ln[1][p] = ln[0][a]; memcpy(fc[1] + p * C, fc[0] + a * C, sizeof(int64_t) * C);
                    ll[1][2 * p] = ll[0][2 * a]; ll[1][2 * p + 1] = ll[0][2 * a + 1]; pr[1][p] = pr[0][a];
                    memcpy(th[1] + p * d, th[0] + a * d, sizeof(double) * d);
                }
                const double mlr = gmean / (double)outer_p * exp(lml);
                { double* t; int32_t* ti; int64_t* tl;
                  t = tm[0]; tm[0] = tm[1]; tm[1] = t; ti = ty[0]; ty[0] = ty[1]; ty[1] = ti; tl = ln[0]; ln[0] = ln[1]; ln[1] = tl;
                  tl = fc[0]; fc[0] = fc[1]; fc[1] = tl; t = ll[0]; ll[0] = ll[1]; ll[1] = t; t = pr[0]; pr[0] = pr[1]; pr[1] = t;
                  t = th[0]; th[0] = th[1]; th[1] = t; }
                k_log[0] += outer_p * n_props;
                for (int mk = 0; mk < n_props; ++mk) {           /* :203-219; tj frozen within a sweep when threads > 1 */
                    const uint64_t kf = sm64(&key_ctr);
                    if (threads == 1) {
                        for (int64_t p = 0; p < outer_p; ++p) {
                            double* tf = thf + p * d;
                            double z[DPOMP_MAX_PARAMS];
                            for (int i = 0; i < d; ++i) z[i] = rnorm(&rng);
                            for (int i = 0; i < d; ++i) { double s = 0.0; for (int k = 0; k <= i; ++k) s += propd[i * d + k] * z[k];
                                tf[i] = (ind_prop ? mu[i] + s : th[0][p * d + i] + tj * s); }
                            const double prf = prior_lp(prior_lo, prior_hi, tf, d);
                            double llf[2] = {0.0, 0.0};
                            if (prf == -INFINITY) { llf[0] = llf[1] = -INFINITY; }
                            else orc_mbp_propose(m, th[0] + p * d, tf, tm[0] + p * cap, ty[0] + p * cap, ln[0][p], tm[2] + p * cap, ty[2] + p * cap,
                                                 &ln[2][p], cap, fc[2] + p * C, llf, oi, kf, (uint32_t)p);
                            if (exp(prf - pr[0][p]) * exp(llf[0] - ll[0][2 * p]) > runif(&rng)) {   /* :212 */
                                mtd[p] = exp(llf[1]);
                                memcpy(tm[0] + p * cap, tm[2] + p * cap, sizeof(double) * ln[2][p]);
                                memcpy(ty[0] + p * cap, ty[2] + p * cap, sizeof(int32_t) * ln[2][p]);
                                ln[0][p] = ln[2][p]; memcpy(fc[0] + p * C, fc[2] + p * C, sizeof(int64_t) * C);
                                ll[0][2 * p] = llf[0]; ll[0][2 * p + 1] = llf[1]; pr[0][p] = prf;
                                memcpy(th[0] + p * d, tf, sizeof(double) * d);
                                k_log[1] += 1; tj *= alpha;
                            } else tj *= 0.999;
                        }
                    } else {
                        for (int64_t p = 0; p < outer_p; ++p) {
                            double z[DPOMP_MAX_PARAMS];
                            for (int i = 0; i < d; ++i) z[i] = rnorm(&rng);
                            for (int i = 0; i < d; ++i) { double s = 0.0; for (int k = 0; k <= i; ++k) s += propd[i * d + k] * z[k];
                                thf[p * d + i] = (ind_prop ? mu[i] + s : th[0][p * d + i] + tj * s); }
                            uacc[p] = runif(&rng);
                        }
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
                        for (int64_t p = 0; p < outer_p; ++p) {
                            pr[2][p] = prior_lp(prior_lo, prior_hi, thf + p * d, d);
                            ll[2][2 * p] = ll[2][2 * p + 1] = 0.0;
                            if (pr[2][p] == -INFINITY) ll[2][2 * p] = ll[2][2 * p + 1] = -INFINITY;
                            else orc_mbp_propose(m, th[0] + p * d, thf + p * d, tm[0] + p * cap, ty[0] + p * cap, ln[0][p], tm[2] + p * cap,
                                                 ty[2] + p * cap, &ln[2][p], cap, fc[2] + p * C, ll[2] + 2 * p, oi, kf, (uint32_t)p);
                        }
                        for (int64_t p = 0; p < outer_p; ++p) {
                            if (exp(pr[2][p] - pr[0][p]) * exp(ll[2][2 * p] - ll[0][2 * p]) > uacc[p]) {
                                mtd[p] = exp(ll[2][2 * p + 1]);
                                memcpy(tm[0] + p * cap, tm[2] + p * cap, sizeof(double) * ln[2][p]);
                                memcpy(ty[0] + p * cap, ty[2] + p * cap, sizeof(int32_t) * ln[2][p]);
                                ln[0][p] = ln[2][p]; memcpy(fc[0] + p * C, fc[2] + p * C, sizeof(int64_t) * C);
                                ll[0][2 * p] = ll[2][2 * p]; ll[0][2 * p + 1] = ll[2][2 * p + 1]; pr[0][p] = pr[2][p];
                                memcpy(th[0] + p * d, thf + p * d, sizeof(double) * d);
                                k_log[1] += 1; tj *= alpha;
                            } else tj *= 0.999;
                        }
                    }
                }
                double mmean = 0.0;
                for (int64_t p = 0; p < outer_p; ++p) mmean += mtd[p];
                bme[1] += log(mlr / (mmean / (double)outer_p));  /* :224 */
                for (int64_t p = 0; p < outer_p; ++p) w[p] = 1.0;
            } else {
                double swg2 = 0.0, sw2 = 0.0;
                for (int64_t p = 0; p < outer_p; ++p) { swg2 += w[p] * gx[p]; sw2 += w[p]; }
                bme[1] += log(swg2 / sw2);                       /* :228 */
            }
        }
        for (int64_t p = 0; p < outer_p; ++p) tcur[p] = m->obs_time[oi - 1];   /* :234 */
    }
    orc_compute_is_mu_covar(mu, cv, th[0], w, d, outer_p);
    memcpy(theta, th[0], sizeof(double) * outer_p * d);
    bme[0] = -bme[0]; bme[1] = -bme[1];
    for (int s = 0; s < 3; ++s) { free(tm[s]); free(ty[s]); free(ln[s]); free(fc[s]); free(ll[s]); free(pr[s]); free(th[s]); }
    free(gx); free(mtd); free(tcur); free(nidx); free(cwv); free(rr); free(thf); free(uacc);
    return 0;
}
