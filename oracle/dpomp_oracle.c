/*
 * dpomp_oracle.c -- CPU restatement of DiscretePOMP.jl's particle-filter hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libdpomp.so) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference at function level -- the reference (pure Julia; no Julia toolchain
 * in this image) cannot be run here and ships no known-answer vectors for this path.  What pins the
 * restatement instead (tests/test_oracle.py): the reference's seeded end-to-end numbers (test/runtests.jl:35,43,51) and
 * the surveyor's anchors (PF log-lik -15.69 +- 0.01 on data/pooley.csv), to Monte-Carlo accuracy; closed-form laws of the
 * simulated process; and an exactly solvable case (pure-death process: likelihood by the forward recursion over its
 * hidden states, evidence and posterior by quadrature) for the filter and for the outer layers.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference repository).
 * Uniform random numbers come from Philox4x32-10 with the counter layout of DESIGN.md ("random streams") so
 * that the GPU path and this oracle consume IDENTICAL draws; the reference uses Julia's global RNG, whose
 * stream is version dependent, so draw-level parity with the reference itself is impossible.
 *
 * Two arithmetic modes for the weight normalisation / resampling part:
 *   ORC_MODE_LITERAL (0): exactly the reference: running linear-domain cumulative sum of exp(log g),
 *                         log(cw[end]/N), sequential walk `while u[i] > cw[j]` (src/hmm_particle_filter.jl:29-30,60,
 *                         src/hmm_pf_resample.jl:24-42).
 *   ORC_MODE_DEVICE  (1): same mathematics evaluated in the device's deterministic order: log-sum-exp with the
 *                         tile-relative scaling, the fixed scan tree and the counting form of the search
 *                         (DESIGN.md "deterministic scan tree").  Used for bit-exact ancestor parity.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/dpomp.h"

#define ORC_MODE_LITERAL 0
#define ORC_MODE_DEVICE 1
#define ORC_MODE_DEVICE_INTERLEAVED 2 /* device order + dpomp_pf_set_scatter(DPOMP_SCATTER_INTERLEAVED) */
#define ORC_GROUP_TILES 32 /* tiles per group of the two-level combine (kGroupTiles of the device code) */

/* ------------------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants).                                         */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox2x32-10 (same paper): 64-bit counter, 32-bit key, two output words */
void orc_philox2x32_10(const uint32_t ctr_in[2], uint32_t key, uint32_t out[2]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t n0 = (uint32_t)(p >> 32) ^ key ^ c1;
        c1 = (uint32_t)p;
        c0 = n0;
        key += 0x9E3779B9u;
    }
    out[0] = c0; out[1] = c1;
}

#define ORC_TAG_SIM 0u
#define ORC_TAG_RESAMPLE 1u

static inline void stream_draw(uint64_t key, uint32_t particle, uint32_t filter, uint32_t obs, uint32_t tag,
                               uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {particle, filter, obs, (tag << 30) | block};
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    orc_philox4x32_10(ctr, k, out);
}
/* Event-loop draws (DESIGN.md "random streams"): the (call key, filter, observation) triple is hashed once with
 * Philox4x32-10 into a 32-bit Philox2x32 key K and two counter masks A, B; attempt k of particle n then uses
 * Philox2x32-10(ctr = (n ^ A, k ^ B), key = K): word 0 -> waiting time, word 1 -> event type. */
typedef struct { uint32_t k, a, b; } sim_stream;
static inline sim_stream sim_stream_init(uint64_t key, uint32_t filter, uint32_t obs) {
    uint32_t w[4];
    stream_draw(key, 0u, filter, obs, ORC_TAG_SIM, 0u, w);
    sim_stream s = {w[0], w[1], w[2]};
    return s;
}
static inline void sim_draw(const sim_stream* s, uint32_t particle, uint32_t attempt, uint32_t out[2]) {
    uint32_t ctr[2] = {particle ^ s->a, attempt ^ s->b};
    orc_philox2x32_10(ctr, s->k, out);
}
static inline double u32_open(uint32_t w) { return ((double)w + 0.5) * 0x1.0p-32; }               /* (0,1)  */
static inline double u53(uint32_t hi, uint32_t lo) {                                                /* [0,1)  */
    return (double)(((uint64_t)hi << 21) | (uint64_t)(lo >> 11)) * 0x1.0p-53;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Model closures.                                                                                              */
/* ------------------------------------------------------------------------------------------------------------ */
/* rate_function + cumsum! (src/hmm_particle_filter.jl:20-21; rate expressions src/hmm_examples.jl:103-168) */
static void cum_rates(const dpomp_model_desc* m, const double* th, const int64_t* x, double* cum) {
    double acc = 0.0;
    for (int e = 0; e < m->n_events; ++e) {
        int64_t l1 = m->rate_k1[e], l2 = m->rate_k2[e], dn = m->rate_kd[e];
        for (int c = 0; c < m->n_compartments; ++c) {
            l1 += (int64_t)m->rate_f1[e][c] * x[c];
            l2 += (int64_t)m->rate_f2[e][c] * x[c];
            dn += (int64_t)m->rate_dn[e][c] * x[c];
        }
        double p = m->rate_par[e] >= 0 ? th[m->rate_par[e]] : 1.0;
        double r = (p * (double)l1) * (double)l2;
        if (m->rate_has_den[e]) r = (dn == 0) ? 0.0 : r / (double)dn; /* reference gives NaN on 0/0 and hangs */
        acc = (e == 0) ? r : acc + r;
        cum[e] = acc;
    }
}

/* choose_event (src/hmm_cmn.jl:4-10): 0-based event */
static int choose_event(const double* cum, int n_events, double u) {
    double etc = u * cum[n_events - 1];
    for (int i = 0; i < n_events - 1; ++i)
        if (cum[i] > etc) return i;
    return n_events - 1;
}

/* partial_gaussian_obs_model / gom2 (src/hmm_examples.jl:59-67) */
static double obs_model(const dpomp_model_desc* m, int t, const int64_t* x) {
    double tmp1 = log(1.0 / (sqrt(2.0 * M_PI) * m->obs_sigma));
    double tmp2 = 2.0 * m->obs_sigma * m->obs_sigma;
    int64_t ys = 0, xs = 0;
    for (int v = 0; v < m->n_obs_vals; ++v) ys += (int64_t)m->obs_ymask[v] * m->obs_val[(int64_t)t * m->n_obs_vals + v];
    for (int c = 0; c < m->n_compartments; ++c) xs += (int64_t)m->obs_xmask[c] * x[c];
    int64_t d = ys - xs;
    return tmp1 - ((double)(d * d) / tmp2);
}

/* function-level probes for the reference known-answer vectors (tests/test_reference_kats.py) */
void orc_cum_rates(const dpomp_model_desc* m, const double* th, const int64_t* x, double* cum) { cum_rates(m, th, x, cum); }
int orc_choose_event(const double* cum, int n_events, double u) { return choose_event(cum, n_events, u) + 1; } /* 1-based */
double orc_obs_model(const dpomp_model_desc* m, int t, const int64_t* x) { return obs_model(m, t, x); }

/* the event loop of iterate_particles! (src/hmm_particle_filter.jl:19-27) for ONE particle over (t, tmax].
 * Attempt k (k = number of events so far) draws Philox2x32 words (waiting time, event type). */
static int64_t sim_interval(const dpomp_model_desc* m, const double* th, int64_t* x, double time, double tmax,
                            const sim_stream* st, uint32_t particle, int64_t max_events, int* overflow) {
    double cum[DPOMP_MAX_EVENTS];
    int64_t k = 0;
    *overflow = 0;
    for (;;) {
        cum_rates(m, th, x, cum);
        if (!(cum[m->n_events - 1] > 0.0)) break; /* `== 0.0 && break` (:22); rates are >= 0 here */
        if (k >= max_events) { *overflow = 1; break; } /* event cap: documented divergence (no cap in the reference) */
        uint32_t w[2];
        sim_draw(st, particle, (uint32_t)k, w);
        double u_wait = u32_open(w[0]), u_evt = u32_open(w[1]);
        time -= log(u_wait) / cum[m->n_events - 1]; /* :23 */
        if (time > tmax) break;                     /* :24 */
        int e = choose_event(cum, m->n_events, u_evt);
        for (int c = 0; c < m->n_compartments; ++c) x[c] += m->trans[e][c]; /* :26 */
        ++k;
    }
    return k;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Resamplers on raw weights: rs_* (src/hmm_resample.jl).  `u` holds the rand() draws in consumption order.      */
/* out: 1-based ancestors.  w is cumulated in place like cumsum!/cumsum.                                         */
/* ------------------------------------------------------------------------------------------------------------ */
static void walk_search(const double* cw, int64_t n, const double* u, int64_t n_out, int64_t* out) {
    /* `j = 1; for i: while u[i] > cw[j] j += 1; output[i] = j` (src/hmm_resample.jl:55-60, src/hmm_pf_resample.jl:34-40) */
    int64_t j = 0;
    for (int64_t i = 0; i < n_out; ++i) {
        while (j < n - 1 && u[i] > cw[j]) ++j; /* j<n-1 guard: the reference would throw a BoundsError instead */
        out[i] = j + 1;
    }
}

/* rs_systematic (src/hmm_resample.jl:44-62) / rsp_systematic (src/hmm_pf_resample.jl:24-42) given cumulative weights */
void orc_search_systematic(const double* cw, int64_t n, double r, int64_t* out) {
    double* u = (double*)malloc(sizeof(double) * (size_t)n);
    u[0] = r / (double)n;
    for (int64_t i = 1; i < n; ++i) u[i] = u[0] + ((double)i / (double)n);
    for (int64_t i = 0; i < n; ++i) u[i] *= cw[n - 1];
    walk_search(cw, n, u, n, out);
    free(u);
}
/* rs_stratified (src/hmm_resample.jl:66-83); also the intended rsp_stratified (src/hmm_pf_resample.jl:46-63, broken F6) */
void orc_search_stratified(const double* cw, int64_t n, const double* r, int64_t* out) {
    double* u = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) u[i] = r[i] / (double)n;
    for (int64_t i = 0; i < n; ++i) u[i] += ((double)i / (double)n);
    for (int64_t i = 0; i < n; ++i) u[i] *= cw[n - 1];
    walk_search(cw, n, u, n, out);
    free(u);
}
/* rs_multinomial (src/hmm_resample.jl:4-20); also the intended rsp_multinomial (src/hmm_pf_resample.jl:5-20, broken F6) */
void orc_search_multinomial(const double* cw, int64_t n, const double* r, int64_t n_out, int64_t* out) {
    for (int64_t p = 0; p < n_out; ++p) {
        out[p] = n;
        double chs = r[p] * cw[n - 1];
        for (int64_t p2 = 0; p2 < n - 1; ++p2) {
            if (chs < cw[p2]) { out[p] = p2 + 1; break; }
        }
    }
}
static void cumsum_inplace(double* w, int64_t n) {
    for (int64_t i = 1; i < n; ++i) w[i] = w[i - 1] + w[i];
}
/* rs_type 1/2/3 on RAW weights (w is overwritten with its cumulative sum, as in the reference) */
void orc_rs(int rs_type, double* w, int64_t n, const double* r, int64_t n_out, int64_t* out) {
    cumsum_inplace(w, n);
    if (rs_type == DPOMP_RS_STRATIFIED) orc_search_stratified(w, n, r, out);
    else if (rs_type == DPOMP_RS_MULTINOMIAL) orc_search_multinomial(w, n, r, n_out, out);
    else orc_search_systematic(w, n, r[0], out);
}
/* same searches on ALREADY cumulative weights (rsp_* semantics) */
void orc_rsp(int rs_type, const double* cw, int64_t n, const double* r, int64_t n_out, int64_t* out) {
    if (rs_type == DPOMP_RS_STRATIFIED) orc_search_stratified(cw, n, r, out);
    else if (rs_type == DPOMP_RS_MULTINOMIAL) orc_search_multinomial(cw, n, r, n_out, out);
    else orc_search_systematic(cw, n, r[0], out);
}

/* compute_ess (src/hmm_particle_filter.jl:4-6) */
double orc_compute_ess(const double* w, int64_t n) {
    double s = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < n; ++i) { s += w[i]; s2 += w[i] * w[i]; }
    return s * s / s2;
}

/* compute_is_mu_covar! (src/cmn.jl:91-99); theta is n_theta x n column-major */
void orc_compute_is_mu_covar(double* mu, double* cv, const double* theta, const double* w, int n_theta, int64_t n) {
    double sw = 0.0;
    for (int64_t p = 0; p < n; ++p) sw += w[p];
    for (int i = 0; i < n_theta; ++i) {
        double a = 0.0;
        for (int64_t p = 0; p < n; ++p) a += w[p] * theta[p * n_theta + i];
        mu[i] = a / sw;
        double v = 0.0;
        for (int64_t p = 0; p < n; ++p) { double d = theta[p * n_theta + i] - mu[i]; v += w[p] * (d * d); }
        cv[i * n_theta + i] = v / sw;
        for (int j = 0; j < i; ++j) {
            double c = 0.0;
            for (int64_t p = 0; p < n; ++p) c += w[p] * (theta[p * n_theta + i] - mu[i]) * (theta[p * n_theta + j] - mu[j]);
            cv[i * n_theta + j] = cv[j * n_theta + i] = c / sw;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Device-order arithmetic (ORC_MODE_DEVICE): the fixed scan tree of DESIGN.md.                                  */
/* A tile holds `tile` values; lane l of warp w owns items (w*32+l)*items .. +items-1.                           */
/* ------------------------------------------------------------------------------------------------------------ */
/* inclusive (incl) and exclusive (excl) tile-local scans and the tile total, in the device's association order */
static double tile_scan(const double* a, int tile, int items, double* incl, double* excl) {
    int nwarps = tile / (32 * items);
    double warp_prefix = 0.0, total = 0.0;
    for (int w = 0; w < nwarps; ++w) {
        double lane_tot[32], scan[32];
        for (int l = 0; l < 32; ++l) {
            const double* q = a + ((size_t)w * 32 + l) * items;
            double r = q[0];
            for (int k = 1; k < items; ++k) r = r + q[k];
            lane_tot[l] = r;
            scan[l] = r;
        }
        for (int d = 1; d < 32; d <<= 1) { /* Kogge-Stone over lanes, as __shfl_up_sync rounds */
            double nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = (l >= d) ? scan[l - d] + scan[l] : scan[l];
            memcpy(scan, nxt, sizeof(scan));
        }
        for (int l = 0; l < 32; ++l) {
            double base = warp_prefix + (l > 0 ? scan[l - 1] : 0.0);
            const double* q = a + ((size_t)w * 32 + l) * items;
            double r = 0.0;
            for (int k = 0; k < items; ++k) {
                size_t idx = ((size_t)w * 32 + l) * items + k;
                if (excl) excl[idx] = base + r;
                r = (k == 0) ? q[0] : r + q[k];
                if (incl) incl[idx] = base + r;
            }
        }
        total = warp_prefix + scan[31];
        warp_prefix = total;
        (void)lane_tot;
    }
    return total;
}

typedef struct {
    int rs_type;
    int64_t n;
    double s;      /* grand total (cw[end]) */
    double r1;     /* systematic: the single rand() */
    uint64_t key;  /* stratified: per-offspring draws */
    uint32_t filter, obs;
} ecount_ctx;

static inline double strat_draw(const ecount_ctx* c, int64_t i0 /*0-based offspring*/) {
    uint32_t w[4];
    stream_draw(c->key, (uint32_t)i0, c->filter, c->obs, ORC_TAG_RESAMPLE, 1u, w);
    return u53(w[0], w[1]);
}
/* u_i of rsp_systematic / rs_stratified for 1-based i, exactly the reference's expression order */
static inline double u_of(const ecount_ctx* c, int64_t i) {
    double n = (double)c->n;
    if (c->rs_type == DPOMP_RS_STRATIFIED) return ((strat_draw(c, i - 1) / n) + ((double)(i - 1) / n)) * c->s;
    return ((c->r1 / n) + ((double)(i - 1) / n)) * c->s;
}
/* E(v) = #{ i in 1..n : u_i <= v }  (u_i is non-decreasing in i) */
static int64_t ecount(const ecount_ctx* c, double v) {
    if (!(c->s > 0.0)) return c->n;
    double n = (double)c->n;
    double g = (c->rs_type == DPOMP_RS_STRATIFIED) ? floor((v / c->s) * n) : floor(((v / c->s) - (c->r1 / n)) * n) + 1.0;
    if (!(g >= 0.0)) g = 0.0;
    if (g > n) g = n;
    int64_t e = (int64_t)g;
    while (e < c->n && u_of(c, e + 1) <= v) ++e;
    while (e > 0 && u_of(c, e) > v) --e;
    return e;
}

/* normalise + resample one filter in device order.  logw[n] -> returns log-lik increment; anc (1-based, may be NULL
 * when do_resample == 0). */
static double device_normalise_resample(const double* logw, int64_t n, int tile, int items, int do_resample,
                                        int rs_type, uint64_t key, uint32_t filter, uint32_t obs, int64_t* anc) {
    int64_t ntiles = (n + tile - 1) / tile;
    double* a = (double*)malloc(sizeof(double) * (size_t)tile);
    double* incl = (double*)malloc(sizeof(double) * (size_t)tile);
    double* m_b = (double*)malloc(sizeof(double) * (size_t)ntiles);
    double* s_b = (double*)malloc(sizeof(double) * (size_t)ntiles);
    double* f_b = (double*)malloc(sizeof(double) * (size_t)ntiles);
    double* off = (double*)malloc(sizeof(double) * (size_t)(ntiles + 1));
    double big_m = -INFINITY;
    for (int64_t b = 0; b < ntiles; ++b) {
        double mb = -INFINITY;
        for (int q = 0; q < tile; ++q) {
            int64_t p = b * tile + q;
            if (p < n && logw[p] > mb) mb = logw[p];
        }
        m_b[b] = mb;
        double ref = (mb == -INFINITY) ? 0.0 : mb;
        for (int q = 0; q < tile; ++q) {
            int64_t p = b * tile + q;
            a[q] = (p < n && logw[p] != -INFINITY) ? exp(logw[p] - ref) : 0.0;
        }
        s_b[b] = tile_scan(a, tile, items, NULL, NULL);
        if (mb > big_m) big_m = mb;
    }
    /* two-level combine (DESIGN.md "deterministic scan tree"): groups of ORC_GROUP_TILES tiles, then the groups.
     *   level 1: m_g = max m_b; f_{b|g} = exp(m_b - m_g); o_{b|g} = Kogge-Stone exclusive scan of f_{b|g} s_b over the 32 lanes
     *   level 2: M = max m_g; F_g = exp(m_g - M); O_g = the same 32-lane scan of F_g s_g, chunks of 32 chained sequentially
     *   cw_q = O_g + F_g * (o_{b|g} + f_{b|g} * incl_q);  off[b] = cw at incl = 0 */
    const int64_t ngroups = (ntiles + ORC_GROUP_TILES - 1) / ORC_GROUP_TILES;
    double* g_m = (double*)malloc(sizeof(double) * (size_t)ngroups);
    double* g_s = (double*)malloc(sizeof(double) * (size_t)ngroups);
    double* g_f = (double*)malloc(sizeof(double) * (size_t)ngroups);
    double* g_off = (double*)malloc(sizeof(double) * (size_t)ngroups);
    double* t_o = (double*)malloc(sizeof(double) * (size_t)ntiles);
    big_m = -INFINITY;
    for (int64_t g = 0; g < ngroups; ++g) {
        double mg = -INFINITY, scan[32];
        for (int l = 0; l < 32; ++l) {
            const int64_t b = g * ORC_GROUP_TILES + l;
            if (b < ntiles && m_b[b] > mg) mg = m_b[b];
        }
        for (int l = 0; l < 32; ++l) {
            const int64_t b = g * ORC_GROUP_TILES + l;
            double f = 0.0, sb = 0.0;
            if (b < ntiles) { f = (m_b[b] == -INFINITY) ? 0.0 : exp(m_b[b] - mg); sb = s_b[b]; f_b[b] = f; }
            scan[l] = f * sb;
        }
        for (int d = 1; d < 32; d <<= 1) {
            double nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = (l >= d) ? scan[l - d] + scan[l] : scan[l];
            memcpy(scan, nxt, sizeof(scan));
        }
        for (int l = 0; l < 32; ++l) {
            const int64_t b = g * ORC_GROUP_TILES + l;
            if (b < ntiles) t_o[b] = (l > 0) ? scan[l - 1] : 0.0;
        }
        g_m[g] = mg; g_s[g] = scan[31];
        if (mg > big_m) big_m = mg;
    }
    double carry = 0.0;
    for (int64_t c0 = 0; c0 < ngroups; c0 += 32) { /* same tree as level 1: Kogge-Stone over 32 lanes, chunks chained */
        double scan[32];
        for (int l = 0; l < 32; ++l) {
            const int64_t g = c0 + l;
            if (g < ngroups) {
                g_f[g] = (g_m[g] == -INFINITY) ? 0.0 : exp(g_m[g] - big_m);
                scan[l] = g_f[g] * g_s[g];
            } else scan[l] = 0.0;
        }
        for (int d = 1; d < 32; d <<= 1) {
            double nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = (l >= d) ? scan[l - d] + scan[l] : scan[l];
            memcpy(scan, nxt, sizeof(scan));
        }
        for (int l = 0; l < 32; ++l)
            if (c0 + l < ngroups) g_off[c0 + l] = carry + ((l > 0) ? scan[l - 1] : 0.0);
        carry = carry + scan[31];
    }
    double big_s = carry;
#define ORC_CW(b, inc) (g_off[(b) / ORC_GROUP_TILES] + g_f[(b) / ORC_GROUP_TILES] * (t_o[b] + f_b[b] * (inc)))
    for (int64_t b = 0; b < ntiles; ++b) off[b] = g_off[b / ORC_GROUP_TILES] + g_f[b / ORC_GROUP_TILES] * (t_o[b] + 1.0 * 0.0);
    off[ntiles] = big_s;
    double ll = big_m + log(big_s / (double)n);
    if (do_resample && rs_type != DPOMP_RS_MULTINOMIAL) {
        uint32_t w4[4];
        stream_draw(key, 0u, filter, obs, ORC_TAG_RESAMPLE, 0u, w4);
        ecount_ctx ctx = {rs_type, n, big_s, u53(w4[0], w4[1]), key, filter, obs};
        int64_t* e = (int64_t*)malloc(sizeof(int64_t) * (size_t)tile);
        for (int64_t b = 0; b < ntiles; ++b) {
            double ref = (m_b[b] == -INFINITY) ? 0.0 : m_b[b];
            for (int q = 0; q < tile; ++q) {
                int64_t p = b * tile + q;
                a[q] = (p < n && logw[p] != -INFINITY) ? exp(logw[p] - ref) : 0.0;
            }
            tile_scan(a, tile, items, incl, NULL);
            int64_t lo = (b == 0) ? 0 : ecount(&ctx, off[b]);
            int64_t hi = (b == ntiles - 1) ? n : ecount(&ctx, off[b + 1]);
            int64_t nvalid = (n - b * tile < tile) ? (n - b * tile) : tile;
            for (int q = 0; q < nvalid; ++q) {
                int64_t ev = ecount(&ctx, ORC_CW(b, incl[q]));
                if (ev < lo) ev = lo;
                if (ev > hi) ev = hi;
                e[q] = ev;
            }
            e[nvalid - 1] = hi;
            int q = 0;
            for (int64_t i = lo + 1; i <= hi; ++i) {
                while (e[q] < i) ++q;
                anc[i - 1] = b * tile + q + 1;
            }
        }
        free(e);
    } else if (do_resample) {
        /* multinomial: offspring i draws chs = r_i * S; ancestor = first p2 < n with chs < cw[p2], else n
         * (src/hmm_resample.jl:9-16), on the device-order cumulative weights */
        double* cw = (double*)malloc(sizeof(double) * (size_t)n);
        for (int64_t b = 0; b < ntiles; ++b) {
            double ref = (m_b[b] == -INFINITY) ? 0.0 : m_b[b];
            for (int q = 0; q < tile; ++q) {
                int64_t p = b * tile + q;
                a[q] = (p < n && logw[p] != -INFINITY) ? exp(logw[p] - ref) : 0.0;
            }
            tile_scan(a, tile, items, incl, NULL);
            for (int q = 0; q < tile; ++q)
                if (b * tile + q < n) cw[b * tile + q] = ORC_CW(b, incl[q]);
        }
        for (int64_t i = 0; i < n; ++i) {
            uint32_t w4[4];
            stream_draw(key, (uint32_t)i, filter, obs, ORC_TAG_RESAMPLE, 1u, w4);
            double chs = u53(w4[0], w4[1]) * big_s;
            /* first p2 in [0, n-1) with chs < cw[p2]: cw is non-decreasing within a tile but tile seams may be off by
             * an ulp, so search tiles by their offsets first exactly like the device does */
            int64_t lo_t = 0, hi_t = ntiles; /* first tile whose END offset (off[b+1]) is > chs */
            while (lo_t < hi_t) { int64_t mid = (lo_t + hi_t) / 2; if (chs < off[mid + 1]) hi_t = mid; else lo_t = mid + 1; }
            int64_t res = n;
            if (lo_t < ntiles) {
                int64_t b = lo_t, base = b * tile;
                int64_t nvalid = (n - base < tile) ? (n - base) : tile;
                int64_t lo_q = 0, hi_q = nvalid;
                while (lo_q < hi_q) { int64_t mid = (lo_q + hi_q) / 2; if (chs < cw[base + mid]) hi_q = mid; else lo_q = mid + 1; }
                res = base + lo_q + 1;
                if (lo_q >= nvalid) { /* seam: chs < off[b+1] but not < the last cw of the tile: the walk of
                                       * src/hmm_resample.jl:9-16 continues into the following tiles */
                    res = n;
                    for (int64_t q2 = base + nvalid; q2 < n; ++q2)
                        if (chs < cw[q2]) { res = q2 + 1; break; }
                }
            }
            if (res > n) res = n;
            anc[i] = res;
        }
        free(cw);
    }
    free(a); free(incl); free(m_b); free(s_b); free(f_b); free(off);
    free(g_m); free(g_s); free(g_f); free(g_off); free(t_o);
#undef ORC_CW
    return ll;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* partial_log_likelihood! (src/hmm_particle_filter.jl:39-76) with iterate_particles! (:9-33) inlined.           */
/*   pop: n x C column-major int64 (the reference's Matrix{Int64}); ymin/ymax 1-based inclusive.                 */
/*   threads: 1 = faithful single thread; >1 = OpenMP over particles (particles are independent given           */
/*            counter-based draws; the running weight sum stays sequential).                                     */
/*   Optional outputs (may be NULL): logw_last[n], anc_last[n] (1-based, of the last observation if it resampled)*/
/* ------------------------------------------------------------------------------------------------------------ */
int orc_pf_partial(const dpomp_model_desc* m, const double* theta, int64_t n, int64_t* pop, int ymin, int ymax,
                   int rs_type, uint64_t key, uint32_t filter, int mode, int tile, int items, int64_t max_events,
                   int threads, double* out_ll, double* logw_last, int64_t* anc_last, int64_t* n_events,
                   int64_t* n_overflow) {
    const int C = m->n_compartments;
    double t_prev;
    if (ymin == 1) { /* :41-45 */
        for (int64_t p = 0; p < n; ++p)
            for (int c = 0; c < C; ++c) pop[c * n + p] = m->initial_condition[c];
        t_prev = (m->t0_index == 0) ? 0.0 : theta[m->t0_index - 1];
    } else {
        t_prev = m->obs_time[ymin - 2]; /* :47 */
    }
    int64_t* old_p = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n * C));
    double* cw = (double*)malloc(sizeof(double) * (size_t)n);
    double* lw = (double*)malloc(sizeof(double) * (size_t)n);
    int64_t* anc = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    double output = 0.0;
    int64_t ev_total = 0, ovf_total = 0;
    (void)threads;
    for (int oi = ymin; oi <= ymax; ++oi) { /* :54 */
        const int t = oi - 1;
        const double tmax = m->obs_time[t];
        /* iterate_particles! (:9-33) */
        const sim_stream st = sim_stream_init(key, filter, (uint32_t)t);
#pragma omp parallel for schedule(static) num_threads(threads) reduction(+ : ev_total, ovf_total) if (threads > 1)
        for (int64_t p = 0; p < n; ++p) {
            int64_t x[DPOMP_MAX_COMPARTMENTS];
            for (int c = 0; c < C; ++c) x[c] = pop[c * n + p];
            int ovf = 0;
            ev_total += sim_interval(m, theta, x, t_prev, tmax, &st, (uint32_t)p, max_events, &ovf);
            ovf_total += ovf;
            lw[p] = ovf ? -INFINITY : obs_model(m, t, x);
            for (int c = 0; c < C; ++c) pop[c * n + p] = x[c];
        }
        const int has_lik = m->obs_id[t] > 0;                    /* :58 */
        const int do_rs = has_lik && (oi < m->n_obs);            /* :62 */
        if (mode == ORC_MODE_LITERAL) {
            double total = 0.0;
            for (int64_t p = 0; p < n; ++p) { total += exp(lw[p]); cw[p] = total; } /* :29-30 */
            if (has_lik) {
                output += log(cw[n - 1] / (double)n);             /* :60 */
                if (do_rs) {
                    memcpy(old_p, pop, sizeof(int64_t) * (size_t)(n * C)); /* :66 */
                    if (rs_type == DPOMP_RS_SYSTEMATIC) {
                        uint32_t w4[4];
                        stream_draw(key, 0u, filter, (uint32_t)t, ORC_TAG_RESAMPLE, 0u, w4);
                        orc_search_systematic(cw, n, u53(w4[0], w4[1]), anc);
                    } else {
                        double* r = (double*)malloc(sizeof(double) * (size_t)n);
                        for (int64_t i = 0; i < n; ++i) {
                            uint32_t w4[4];
                            stream_draw(key, (uint32_t)i, filter, (uint32_t)t, ORC_TAG_RESAMPLE, 1u, w4);
                            r[i] = u53(w4[0], w4[1]);
                        }
                        if (rs_type == DPOMP_RS_STRATIFIED) orc_search_stratified(cw, n, r, anc);
                        else orc_search_multinomial(cw, n, r, n, anc);
                        free(r);
                    }
                }
            }
        } else {
            double inc = device_normalise_resample(lw, n, tile, items, do_rs, rs_type, key, filter, (uint32_t)t, anc);
            if (has_lik) output += inc;
            if (do_rs) memcpy(old_p, pop, sizeof(int64_t) * (size_t)(n * C));
        }
        if (do_rs) { /* m_pop[i,:] .= old_p[j,:] (src/hmm_pf_resample.jl:38) */
            if (mode == ORC_MODE_DEVICE_INTERLEAVED) {
                /* offspring i goes to row pos(i): 32-row chunk k -> chunk sigma(k) = rank of k by (k mod M, k div M), M = tiles,
                 * over the floor(n / 32) full chunks; identity for one tile and for the trailing partial chunk (include/dpomp.h) */
                const int64_t m_t = (n + tile - 1) / tile, ncf = (m_t > 1) ? (n >> 5) : 0;
                const int64_t pq = ncf / (m_t > 0 ? m_t : 1), pr = ncf % (m_t > 0 ? m_t : 1);
                int64_t* anc2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
                for (int64_t i = 0; i < n; ++i) {
                    const int64_t k = i >> 5;
                    int64_t row = i;
                    if (k < ncf) {
                        const int64_t rr = k % m_t, qq = k / m_t;
                        row = ((rr * pq + (rr < pr ? rr : pr) + qq) << 5) | (i & 31);
                    }
                    anc2[row] = anc[i];
                }
                memcpy(anc, anc2, sizeof(int64_t) * (size_t)n);
                free(anc2);
            }
            for (int c = 0; c < C; ++c)
                for (int64_t i = 0; i < n; ++i) pop[c * n + i] = old_p[c * n + (anc[i] - 1)];
        }
        if (oi == ymax) {
            if (logw_last) memcpy(logw_last, lw, sizeof(double) * (size_t)n);
            if (anc_last && do_rs) memcpy(anc_last, anc, sizeof(int64_t) * (size_t)n);
        }
        t_prev = tmax; /* :72 */
    }
    free(old_p); free(cw); free(lw); free(anc);
    *out_ll = output;
    if (n_events) *n_events = ev_total;
    if (n_overflow) *n_overflow = ovf_total;
    return 0;
}

/* estimate_likelihood (src/hmm_particle_filter.jl:79-84) */
int orc_pf_loglik(const dpomp_model_desc* m, const double* theta, int64_t n, int rs_type, uint64_t key,
                  uint32_t filter, int mode, int tile, int items, int64_t max_events, int threads, double* out_ll,
                  int64_t* n_events) {
    int64_t* pop = (int64_t*)calloc((size_t)(n * m->n_compartments), sizeof(int64_t));
    int rc = orc_pf_partial(m, theta, n, pop, 1, m->n_obs, rs_type, key, filter, mode, tile, items, max_events,
                            threads, out_ll, NULL, NULL, n_events, NULL);
    free(pop);
    return rc;
}

/* a batch of independent filters (the loop of run_pibis src/hmm_ibis.jl:53-56), optionally OpenMP over filters.
 * theta: n_params x B column-major; pops: B consecutive n x C matrices (may be NULL when ymin == 1 and state is not
 * needed afterwards). */
int orc_pf_partial_batch(const dpomp_model_desc* m, const double* theta, int32_t nb, int64_t n, int64_t* pops,
                         int ymin, int ymax, int rs_type, uint64_t key, uint32_t filter0, int mode, int tile, int items,
                         int64_t max_events, int threads, double* out_ll, int64_t* n_events) {
    int64_t ev_total = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : ev_total) if (threads > 1)
    for (int32_t b = 0; b < nb; ++b) {
        int64_t* pop = pops ? pops + (size_t)b * n * m->n_compartments
                            : (int64_t*)calloc((size_t)(n * m->n_compartments), sizeof(int64_t));
        int64_t ev = 0;
        orc_pf_partial(m, theta + (size_t)b * m->n_params, n, pop, ymin, ymax, rs_type, key, filter0 + (uint32_t)b,
                       mode, tile, items, max_events, 1, &out_ll[b], NULL, NULL, &ev, NULL);
        ev_total += ev;
        if (!pops) free(pop);
    }
    if (n_events) *n_events = ev_total;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Single-trajectory Gillespie simulation with the reference's draw order, for generating synthetic data:       */
/* gillespie_sim (src/hmm_sim.jl:86-102) with iterate_particle! (:55-70) and dmy_obs_fn (src/hmm_examples.jl:6-8).*/
/* out_states: T x C row-major final state at each observation time.                                             */
/* ------------------------------------------------------------------------------------------------------------ */
int orc_gillespie_sim(const dpomp_model_desc* m, const double* theta, uint64_t key, int64_t max_events,
                      int64_t* out_states, int64_t* n_events) {
    int64_t x[DPOMP_MAX_COMPARTMENTS];
    for (int c = 0; c < m->n_compartments; ++c) x[c] = m->initial_condition[c];
    double t = (m->t0_index == 0) ? 0.0 : theta[m->t0_index - 1];
    int64_t ev = 0;
    for (int i = 0; i < m->n_obs; ++i) {
        int ovf = 0;
        const sim_stream st = sim_stream_init(key, 0u, (uint32_t)i);
        ev += sim_interval(m, theta, x, t, m->obs_time[i], &st, 0u, max_events, &ovf);
        for (int c = 0; c < m->n_compartments; ++c) out_states[(size_t)i * m->n_compartments + c] = x[c];
        t = m->obs_time[i];
    }
    if (n_events) *n_events = ev;
    return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
