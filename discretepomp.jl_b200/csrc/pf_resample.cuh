// pf_resample.cuh -- the resample phase of one (filter, ancestor tile): counting search + offspring map + gather.
// Shared by the stand-alone resample kernel (pf_kernels.cu) and the fused step kernel (pf_sim.cuh).
//
// Replaces rsp_systematic (src/hmm_pf_resample.jl:24-42) / the intended rsp_stratified (:46-63) and the `old_p .= pop`
// copy of partial_log_likelihood! (src/hmm_particle_filter.jl:66).
#pragma once
#include <type_traits>

#include "dpomp_dev.cuh"

namespace dpomp {

struct RsArgs {              // what the phase needs besides the tile's own data
    const double* tile_f;    // [B][ntiles]
    const double* tile_off;  // [B][ntiles]
    const double* grp_f;     // [B][ngroups]
    const double* grp_off;   // [B][ngroups]
    const double* filt_s;    // [B]
    // filter-local views of level 2 computed by this CTA (deferred level 2 of the stand-alone kernel), or nullptr: then the
    // arrays above, written by the simulate kernel, are read
    const double* l2_f;      // [ngroups]
    const double* l2_off;    // [ngroups]
    const double* l2_s;      // [1]
    int32_t* pop_dst;        // [B][C][n_pad]
    int32_t* anc;            // [B][n_pad] or nullptr
    long long n, n_pad;
    int ntiles, ngroups, n_comp, t, rs_type;
    uint64_t key;
    ChunkPerm perm;          // offspring i is stored at row perm_pos(perm, i)
};

//   cw_q = O_g + F_g * (o_{b|g} + f_{b|g} * incl_q)   (incl_q: tile-local scan, deterministic tree)
//   e_q  = E(cw_q) = #{ i : u_i <= cw_q }               (counting form of `while u[i] > cw[j]`, src/hmm_pf_resample.jl:34-40)
//   offspring (lo_b, hi_b] belong to this tile; offspring i takes the first q with e_q >= i, i.e. ancestor q owns
//   (max_{q'<q} e_q', max_{q'<=q} e_q'].
// One block barrier publishes (lo_b, hi_b) and the per-warp maxima; afterwards every warp works alone on the 32*ITEMS
// ancestors it owns: the offspring -> ancestor map of a window of 32*ITEMS offspring is built in the warp's slice of
// `am_s` by a scatter of the range starts followed by an inclusive max-scan; state rows are copied from the staged
// ancestor states `st_tile[c * stride + q]` (shared memory) with coalesced 128-byte stores.
// The combine outputs are read with ld.cg: in the fused kernel they were written by other SMs during the same launch.
template <typename T, bool CG>
__device__ __forceinline__ T ld_combine(const T* p) {  // CG: the value was written by another SM during this launch
    if constexpr (CG) return __ldcg(p); else return *p;
}

// NC: compartments at compile time (the gather of the stand-alone kernel is instantiated for 2, 3 and 4), 0 = a.n_comp
template <int ITEMS, typename SrcT, bool CG, int RS = 0, bool PERM = true, int NC = 0>
__device__ __forceinline__ void resample_tile(const RsArgs& a, int b, int tile, uint32_t gfilter, const double (&incl)[ITEMS],
                                              const SrcT* st_tile, int stride, int* am_all, int* warp_max_s, long long* lohi_s) {
    constexpr int TILE = kBlockThreads * ITEMS;
    constexpr int CHUNK = 32 * ITEMS;
    constexpr int NW = kBlockThreads / 32;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kHeavyFactor = 8;  // offspring / ancestors of a tile above which the CTA-wide path is used
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long base_n = (long long)tile * TILE;
    const double big_s = a.l2_s ? *a.l2_s : ld_combine<double, CG>(a.filt_s + b);
    const int grp = tile / kGroupTiles;
    auto grp_off_of = [&](int g) -> double { return a.l2_off ? a.l2_off[g] : ld_combine<double, CG>(a.grp_off + (size_t)b * a.ngroups + g); };
    auto grp_f_of = [&](int g) -> double { return a.l2_f ? a.l2_f[g] : ld_combine<double, CG>(a.grp_f + (size_t)b * a.ngroups + g); };
    const double g_off = grp_off_of(grp), g_f = grp_f_of(grp);
    const double t_off = ld_combine<double, CG>(a.tile_off + (size_t)b * a.ntiles + tile), t_f = ld_combine<double, CG>(a.tile_f + (size_t)b * a.ntiles + tile);

    const Philox4 p = stream_draw(a.key, 0u, gfilter, (uint32_t)a.t, kTagResample, 0u);
    const ResampleCtx ctx = make_resample_ctx(a.rs_type, a.n, big_s, a.key, gfilter, (uint32_t)a.t, u53(p.w0, p.w1));

    // tile seams: the offset of a tile is its cw at incl = 0
    if (tid == 0) lohi_s[0] = (tile == 0) ? 0 : resample_ecount<RS>(ctx, tile_cw(g_off, g_f, t_off, t_f, 0.0));
    if (tid == kBlockThreads / 2) {
        long long hi_b = a.n;
        if (tile != a.ntiles - 1) {
            const int g2 = (tile + 1) / kGroupTiles;
            hi_b = resample_ecount<RS>(ctx, tile_cw(grp_off_of(g2), grp_f_of(g2),
                                                ld_combine<double, CG>(a.tile_off + (size_t)b * a.ntiles + tile + 1), 1.0, 0.0));
        }
        lohi_s[1] = hi_b;
    }
    const long long rem = a.n - base_n;
    const int nvalid = rem < TILE ? (int)rem : TILE;
    // raw counts (n_particles < 2^31); the last valid item and the padding close the tile's range (clamped to hi below)
    int er[ITEMS];
#ifdef DPOMP_RS_INLINE_CORRECT
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) er[k] = (int)resample_ecount<RS>(ctx, tile_cw(g_off, g_f, t_off, t_f, incl[k]));
#else
    unsigned todo = 0;  // items whose closed-form guess needs the exact correction
    // (RS == 0, the fused step kernel: the resampler is a block-uniform run-time value)
    const bool sys = RS == DPOMP_RS_SYSTEMATIC || (RS == 0 && a.rs_type == DPOMP_RS_SYSTEMATIC);
#ifndef DPOMP_NO_FUSED_GUESS
    if (sys) {
        // Systematic: the count of an item is floor(t) + 1 with t = cw_q N / S - r unless t lies within 2^-12 of an integer
        // (resample_ecount_guess).  For that decision t may come from ONE fused multiply-add per item,
        //   t ~ K0 + K1 * incl_q,  K1 = F_g f_{b|g} N / S,  K0 = (O_g + F_g o_{b|g}) N / S - r:
        // K0 and K1 carry a few ulp of relative error, so |t_fma - t| < 2 N 2^-50 <= 2^-18 index units for N <= 2^31, far
        // inside the 2^-12 margin; the exactly rounded cw_q of the reference's expression order (tile_cw) is only formed for
        // the items that fall inside the margin, by the correction path below (the same code as before: same counts).
        if (ctx.s > 0.0) {
            const double k1 = g_f * t_f * ctx.inv_s * ctx.dn;
            const double k0 = ((g_off + g_f * t_off) * ctx.inv_s - ctx.r1_over_n) * ctx.dn;
            const double t_hi = ctx.dn - 1.0;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const double t = fma(incl[k], k1, k0);
                const double fl = floor(t);
                const double frac = t - fl;
                const bool exact = frac > 0x1.0p-12 && frac < 1.0 - 0x1.0p-12 && t > 0.0 && t < t_hi;
                er[k] = (int)fl + 1;  // (garbage when !exact: replaced below)
                if (!exact) todo |= 1u << k;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) er[k] = (int)ctx.n;
        }
    } else
#endif
    {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            bool exact;
            er[k] = resample_ecount_guess<RS>(ctx, tile_cw(g_off, g_f, t_off, t_f, incl[k]), exact);
            if (!exact) todo |= 1u << k;
        }
    }
    // one copy of the correction code (not ITEMS inlined ones: instruction-cache footprint); the arrays stay in registers
    // (select chains instead of dynamic indexing).  Systematic: rare (|t - round(t)| < 2^-12); stratified: every item.
#pragma unroll 1
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        double inc_k = incl[0];
        int e_k = er[0];
#pragma unroll
        for (int j = 1; j < ITEMS; ++j) {
            inc_k = (k == j) ? incl[j] : inc_k;
            e_k = (k == j) ? er[j] : e_k;
        }
#ifndef DPOMP_NO_FUSED_GUESS
        if (sys) {  // the fused guess of this item was not usable: the full exact count
            e_k = (int)resample_ecount<RS>(ctx, tile_cw(g_off, g_f, t_off, t_f, inc_k));
        } else
#endif
        e_k = resample_ecount_correct<RS>(ctx, tile_cw(g_off, g_f, t_off, t_f, inc_k), e_k);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) er[j] = (k == j) ? e_k : er[j];
    }
#endif
#pragma unroll
    for (int k = 0; k < ITEMS; ++k)
        if (tid * ITEMS + k >= nvalid - 1) er[k] = 0x7fffffff;
    // running maximum in item order: lane-serial, then Kogge-Stone over the lanes; the warp maximum goes to shared memory
#pragma unroll
    for (int k = 1; k < ITEMS; ++k) er[k] = max(er[k], er[k - 1]);
    int inc = er[ITEMS - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc = max(inc, y);
    }
    int prev_raw = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) prev_raw = -1;
    if (lane == 31) warp_max_s[warp] = inc;
    __syncthreads();  // the only block barrier
    pdl_trigger();
    // the stand-alone kernel stages the ancestor states with cp.async while the counts above are computed: this thread's
    // copies are complete after the wait, its warp's after the __syncwarp (a no-op wait in the fused kernel)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#ifdef DPOMP_PHASE_TIMERS
    if (threadIdx.x == 0 && a.t == DPOMP_PHASE_OBS && blockIdx.x < 4096) g_dpomp_phase[1][blockIdx.x][3] = dpomp_gtime();
#endif

    const int lo = (int)lohi_s[0], hi = (int)lohi_s[1];  // offspring counts of a filter: n_particles < 2^31
    DPOMP_CHECK_IDX(lo, a.n + 1);
    DPOMP_CHECK_IDX(hi, a.n + 1);
    DPOMP_CHECK_IDX(hi - lo, a.n + 1);
    int wprev_raw = -1;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        if (w < warp) wprev_raw = max(wprev_raw, warp_max_s[w]);
    prev_raw = max(prev_raw, wprev_raw);
    // clamp is monotone, so clamp(running max) == running max of the clamped counts; offsets are relative to lo
    auto clamp_off = [&](int v) -> int { return min(max(v, lo), hi) - lo; };
    int emax[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) emax[k] = clamp_off(max(er[k], prev_raw));
    const int prev = clamp_off(prev_raw);                       // owned range of item 0: (prev, emax[0]]
    const int wfirst = clamp_off(wprev_raw);                    // the warp owns offspring offsets (wfirst, wlast]
    const int wlast = __shfl_sync(FULL, emax[ITEMS - 1], 31);

    int32_t* dst_b = a.pop_dst + (size_t)b * a.n_comp * a.n_pad;  // row of offspring i: perm_pos(a.perm, i)
    // Heavy tile (weight collapse: this tile feeds far more offspring than it has ancestors): per-warp windows would leave
    // the warp that owns the heavy ancestors looping alone, so the whole CTA walks the tile's offspring range instead and
    // finds each ancestor by binary search over the running maxima (block-uniform decision: lo and hi are shared).
    if ((long long)hi - lo > (long long)kHeavyFactor * TILE) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am_all[tid * ITEMS + k] = emax[k];
        __syncthreads();
        const int total = (int)(hi - lo);
        for (int o = 1 + tid; o <= total; o += kBlockThreads) {  // 1-based offset in the tile's offspring range
            int lq = 0, hq = TILE - 1;                             // first q with emax_q >= o (emax of the last items = total)
            while (lq < hq) {
                const int mid = (lq + hq) >> 1;
                if (am_all[mid] >= o) hq = mid; else lq = mid + 1;
            }
            const long long row = perm_pos(a.perm, lo + o - 1);
            DPOMP_CHECK_IDX(lq, TILE);
            DPOMP_CHECK_IDX(base_n + lq, a.n);
            DPOMP_CHECK_IDX(row, a.n);
            for (int c = 0; c < a.n_comp; ++c) dst_b[(size_t)c * a.n_pad + row] = (int)st_tile[(size_t)c * stride + lq];
            if (a.anc) a.anc[(size_t)b * a.n_pad + row] = (int32_t)(base_n + lq);
        }
        return;
    }
    int* am_w = am_all + warp * CHUNK;
    for (int wlo = wfirst; wlo < wlast; wlo += CHUNK) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am_w[lane * ITEMS + k] = -1;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int first = (k == 0) ? prev : emax[k - 1];
            if (emax[k] > first && first < wlo + CHUNK && emax[k] > wlo) {
                DPOMP_CHECK_IDX(max(first - wlo, 0), CHUNK);
                am_w[max(first - wlo, 0)] = lane * ITEMS + k;
            }
        }
        __syncwarp();
        int am[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am[k] = am_w[lane * ITEMS + k];
#pragma unroll
        for (int k = 1; k < ITEMS; ++k) am[k] = max(am[k], am[k - 1]);
        int ainc = am[ITEMS - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(FULL, ainc, d);
            if (lane >= d) ainc = max(ainc, y);
        }
        int aprev = __shfl_up_sync(FULL, ainc, 1);
        if (lane == 0) aprev = -1;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am_w[lane * ITEMS + k] = max(am[k], aprev);
        __syncwarp();
        if (!PERM || a.perm.ncf == 0) {  // the reference's row order: offspring i in row i
            // gather: one row of 32 offspring per step (coalesced 128-byte stores), only the live rows of the window -- the
            // second, nearly empty window of a warp with a few offspring more than CHUNK costs one row, not a full window
            // (round 2: the ITEMS-wide predicated form was 79 SASS instructions per compartment and window, 30 % of the kernel)
            const int nlive = min(wlast - wlo, CHUNK);  // warp-uniform, > 0
            const SrcT* sw_ = st_tile + warp * CHUNK;
            int32_t* d0 = dst_b + lo + wlo + lane;
#pragma unroll 2
            for (int r0 = 0; r0 < nlive; r0 += 32) {
                const int pidx = r0 + lane;
                const bool live = pidx < nlive;
                const int sq = am_w[live ? pidx : 0];  // slot 0 of a non-empty window is live
                if (live) {
                    DPOMP_CHECK_IDX(sq, CHUNK);
                    DPOMP_CHECK_IDX(base_n + warp * CHUNK + sq, a.n);
                    DPOMP_CHECK_IDX(lo + wlo + pidx, a.n);
                }
                if constexpr (NC > 0) {
                    int vals[NC];
#pragma unroll
                    for (int c = 0; c < NC; ++c) vals[c] = (int)sw_[(size_t)c * stride + sq];
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        if (live) d0[(size_t)c * a.n_pad + r0] = vals[c];
                } else {
                    for (int c = 0; c < a.n_comp; ++c) {
                        const int v = (int)sw_[(size_t)c * stride + sq];
                        if (live) d0[(size_t)c * a.n_pad + r0] = v;
                    }
                }
                if (a.anc && live) a.anc[(size_t)b * a.n_pad + lo + wlo + pidx] = (int32_t)(base_n + warp * CHUNK + sq);
            }
        } else {
            // gather of the interleaved placement: striped over the window; per compartment the ITEMS loads of a lane are
            // issued before its stores
            int srcq[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) {
                const int pidx = j * 32 + lane;
                srcq[j] = (wlo + pidx < wlast) ? am_w[pidx] : -1;
                if (srcq[j] >= 0) {  // window entry: an ancestor of this warp's chunk; offspring row inside the filter
                    DPOMP_CHECK_IDX(srcq[j], CHUNK);
                    DPOMP_CHECK_IDX(base_n + warp * CHUNK + srcq[j], a.n);
                    DPOMP_CHECK_IDX(lo + wlo + pidx, a.n);
                } else {
                    DPOMP_CHECK_IDX(wlo + pidx < wlast ? -1 : 0, 1);  // a live window slot must have an ancestor
                }
            }
            // Row of offspring i = lo + wlo + j * 32 + lane: consecutive j advance the chunk index by one, so
            // (k mod M, k div M) is kept incrementally (one division per window).  n_pad < 2^31: int rows.
            int row[ITEMS];
            {
                const int i0 = (int)(lo + wlo) + lane;
                int k = i0 >> 5;
                int rr = k % a.perm.m, qq = k / a.perm.m;
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) {
                    row[j] = (k < a.perm.ncf) ? ((chunk_sigma(a.perm, rr, qq) << 5) | (i0 & 31)) : i0 + j * 32;
                    if (srcq[j] >= 0) DPOMP_CHECK_IDX(row[j], a.n);
                    ++k;
                    if (++rr == a.perm.m) { rr = 0; ++qq; }
                }
            }
            for (int c = 0; c < a.n_comp; ++c) {
                const SrcT* sc = st_tile + (size_t)c * stride + warp * CHUNK;
                int32_t* dc = dst_b + (size_t)c * a.n_pad;
                int vals[ITEMS];
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                    if (srcq[j] >= 0) vals[j] = (int)sc[srcq[j]];
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                    if (srcq[j] >= 0) dc[row[j]] = vals[j];
            }
            if (a.anc) {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                    if (srcq[j] >= 0) a.anc[(size_t)b * a.n_pad + row[j]] = (int32_t)(base_n + warp * CHUNK + srcq[j]);
            }
        }
        __syncwarp();
    }
}

}  // namespace dpomp
