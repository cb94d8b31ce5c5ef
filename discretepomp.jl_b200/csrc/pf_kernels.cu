// pf_kernels.cu -- kernel 2 of the particle filter (scan + search + gather fused), the multinomial variant, the
// outer-layer filter gather / pack kernels and the bit-exactness hook for the resampling searches.
//
// Replaces rsp_systematic (src/hmm_pf_resample.jl:24-42), the intended rsp_stratified / rsp_multinomial (:46-63, :5-20,
// broken in the reference, specified by rs_stratified / rs_multinomial src/hmm_resample.jl:66-83, 4-20), the
// `old_p .= pop` copy of partial_log_likelihood! (src/hmm_particle_filter.jl:66) and the population gathers of
// run_pibis (src/hmm_ibis.jl:71-79, 105-108).
#include "dpomp_dev.cuh"
#include "dpomp_internal.cuh"

namespace dpomp {

cudaError_t launch_sim_f32(const ModelHost& m, int items, const SimLaunch& a, cudaStream_t stream);
cudaError_t launch_sim_f64(const ModelHost& m, int items, const SimLaunch& a, cudaStream_t stream);

cudaError_t launch_sim_weight(const ModelHost& m, int sim_precision, int items, const SimLaunch& a, cudaStream_t stream) {
    return sim_precision == DPOMP_SIM_F64 ? launch_sim_f64(m, items, a, stream) : launch_sim_f32(m, items, a, stream);
}

// does the rate table equal one of the hand-specialised predefined models (pf_sim.cuh Builtin<>)?
namespace {
struct BuiltinSpec { int id, C, E; int A[3], B[3]; int T[3][4]; };
const BuiltinSpec kSpecs[] = {
    {1, 2, 1, {0}, {1}, {{-1, 1}}},
    {2, 3, 2, {0, 1}, {1, -1}, {{-1, 1, 0}, {0, -1, 1}}},
    {3, 2, 2, {0, 1}, {1, -1}, {{-1, 1}, {1, -1}}},
    {4, 3, 2, {0, 1}, {2, -1}, {{-1, 1, 0}, {0, -1, 1}}},
    {5, 4, 3, {0, 1, 2}, {2, -1, -1}, {{-1, 1, 0, 0}, {0, -1, 1, 0}, {0, 0, -1, 1}}},
    {6, 3, 3, {0, 1, 2}, {2, -1, -1}, {{-1, 1, 0}, {0, -1, 1}, {1, 0, -1}}},
    {7, 2, 3, {1, 0, 0}, {-1, 1, -1}, {{0, 1}, {1, -1}, {-1, 0}}},
};
}  // namespace
int builtin_model_id(const dpomp_model_desc& d) {
    for (const BuiltinSpec& s : kSpecs) {
        if (d.n_compartments != s.C || d.n_events != s.E) continue;
        bool ok = true;
        for (int e = 0; e < s.E && ok; ++e) {
            ok = d.rate_par[e] == e && d.rate_has_den[e] == 0 && d.rate_k1[e] == 0 && d.rate_k2[e] == (s.B[e] >= 0 ? 0 : 1);
            for (int c = 0; c < s.C && ok; ++c)
                ok = d.rate_f1[e][c] == (c == s.A[e] ? 1 : 0) && d.rate_f2[e][c] == (c == s.B[e] ? 1 : 0) && d.trans[e][c] == s.T[e][c];
        }
        if (ok) return s.id;
    }
    return 0;
}

int sim_kernel_supported(int n_comp, int n_events) {
    return n_comp >= 1 && n_comp <= 8 && n_events >= 1 && n_events <= 8;
}

// ------------------------------------------------------------------------------------------------------------
// Kernel 2: one CTA per (filter, ancestor tile).  The tile's cumulative weights never leave the SM:
//   w_q = exp(logw_q - m_b);  cw_q = off_b + f_b * incl_q   (deterministic scan tree)
//   e_q = E(cw_q) = number of offspring whose uniform is <= cw_q  (counting form of `while u[i] > cw[j]`)
//   offspring (lo_b, hi_b] belong to this tile; offspring i takes the first q with e_q >= i, i.e. ancestor q owns
//   (max_{q'<q} e_q', max_{q'<=q} e_q'].  The offspring -> ancestor map of a window of TILE offspring is built in
//   shared memory by a scatter of the range starts followed by an inclusive max-scan (no per-offspring search);
//   state rows are then copied with coalesced writes over i and near-sorted reads over q.
// ------------------------------------------------------------------------------------------------------------
// inclusive max-scan over the block in blocked order (thread owns ITEMS consecutive values); returns in `v`, and the
// inclusive value of the previous thread's last item in `prev` (identity for thread 0).  Two __syncthreads().
template <int ITEMS>
__device__ __forceinline__ void block_max_scan(int (&v)[ITEMS], int& prev, int identity, int* warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 1; k < ITEMS; ++k) v[k] = max(v[k], v[k - 1]);
    int inc = v[ITEMS - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = max(inc, y);
    }
    int pl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) pl = identity;
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int wp = identity;
#pragma unroll
    for (int w = 0; w < kBlockThreads / 32; ++w)
        if (w < warp) wp = max(wp, warp_tot[w]);
    prev = max(wp, pl);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) v[k] = max(v[k], prev);
    __syncthreads();
}

template <int ITEMS>
__global__ void __launch_bounds__(kBlockThreads) pf_resample_kernel(const __grid_constant__ ResampleLaunch a) {
    constexpr int TILE = kBlockThreads * ITEMS;
    __shared__ int am_s[TILE];  // offspring window -> local ancestor index
    __shared__ double warp_scratch[kBlockThreads / 32];
    __shared__ int warp_iscratch[kBlockThreads / 32];
    __shared__ long long lohi_s[2];

    const int tid = threadIdx.x;
    const int tile = blockIdx.x % a.ntiles;
    const int b = blockIdx.x / a.ntiles;
    const long long base_n = (long long)tile * TILE;
    const uint32_t gfilter = a.filter_ids ? a.filter_ids[b] : a.filter0 + (uint32_t)b;

    const double big_s = a.filt_s[b];
    const double off_b = a.tile_off[(size_t)b * (a.ntiles + 1) + tile];
    const double off_n = a.tile_off[(size_t)b * (a.ntiles + 1) + tile + 1];
    const double f_b = a.tile_f[(size_t)b * a.ntiles + tile];

    double av[ITEMS], incl[ITEMS], excl[ITEMS];
    const double* wt = a.wtile + (size_t)b * a.n_pad + base_n + (size_t)tid * ITEMS;  // exp(logw - m_b) from kernel 1
    if constexpr (ITEMS % 2 == 0) {  // 128-bit loads
#pragma unroll
        for (int k = 0; k < ITEMS; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(wt + k);
            av[k] = v.x;
            av[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) av[k] = wt[k];
    }
    tile_scan<ITEMS>(av, incl, excl, warp_scratch);

    if (a.rs_type == DPOMP_RS_MULTINOMIAL) {  // materialise cw; the per-offspring search is a second kernel
        double* cw = a.cw + (size_t)b * a.n_pad + base_n + (size_t)tid * ITEMS;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) cw[k] = __dadd_rn(off_b, __dmul_rn(f_b, incl[k]));
        return;
    }

    const Philox4 p = stream_draw(a.key, 0u, gfilter, (uint32_t)a.t, kTagResample, 0u);
    const ResampleCtx ctx = make_resample_ctx(a.rs_type, a.n, big_s, a.key, gfilter, (uint32_t)a.t, u53(p.w0, p.w1));

    if (tid == 0) lohi_s[0] = (tile == 0) ? 0 : resample_ecount(ctx, off_b);
    if (tid == 32) lohi_s[1] = (tile == a.ntiles - 1) ? a.n : resample_ecount(ctx, off_n);
    int er[ITEMS];  // n_particles < 2^31
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) er[k] = (int)resample_ecount(ctx, __dadd_rn(off_b, __dmul_rn(f_b, incl[k])));
    __syncthreads();

    const long long lo = lohi_s[0], hi = lohi_s[1];
    const long long rem = a.n - base_n;
    const int nvalid = rem < TILE ? (int)rem : TILE;
    // owned offspring as offsets from lo: item q owns (prev_q, emax_q]; the last valid item closes the tile's range
    int emax[ITEMS], prev;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int qk = tid * ITEMS + k;
        long long ev = er[k] < lo ? lo : (er[k] > hi ? hi : (long long)er[k]);
        if (qk >= nvalid - 1) ev = hi;
        emax[k] = (int)(ev - lo);
    }
    block_max_scan<ITEMS>(emax, prev, 0, warp_iscratch);

    const int total = (int)(hi - lo);
    const int32_t* src_b = a.pop_src + (size_t)b * a.n_comp * a.n_pad;
    int32_t* dst_b = a.pop_dst + (size_t)b * a.n_comp * a.n_pad;
    for (int wlo = 0; wlo < total; wlo += TILE) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am_s[tid * ITEMS + k] = -1;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int first = (k == 0) ? prev : emax[k - 1];  // offspring offsets (first, emax[k]] belong to item k
            if (emax[k] > first && first < wlo + TILE && emax[k] > wlo) am_s[max(first - wlo, 0)] = tid * ITEMS + k;
        }
        __syncthreads();
        int am[ITEMS], dummy;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am[k] = am_s[tid * ITEMS + k];
        block_max_scan<ITEMS>(am, dummy, -1, warp_iscratch);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) am_s[tid * ITEMS + k] = am[k];
        __syncthreads();
        // gather: striped over the window; per compartment, the ITEMS loads of a thread are issued before its stores
        int srcq[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int pidx = j * kBlockThreads + tid;
            srcq[j] = (wlo + pidx < total) ? am_s[pidx] : -1;
        }
        const long long i_base = lo + wlo + tid;  // 0-based offspring index of j = 0
        for (int c = 0; c < a.n_comp; ++c) {
            const int32_t* sc = src_b + (size_t)c * a.n_pad + base_n;
            int32_t* dc = dst_b + (size_t)c * a.n_pad + i_base;
            int vals[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
                if (srcq[j] >= 0) vals[j] = sc[srcq[j]];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
                if (srcq[j] >= 0) dc[j * kBlockThreads] = vals[j];
        }
        if (a.anc) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
                if (srcq[j] >= 0) a.anc[(size_t)b * a.n_pad + i_base + j * kBlockThreads] = (int32_t)(base_n + srcq[j]);
        }
        __syncthreads();
    }
}

// multinomial: offspring i draws chs = r_i * S; ancestor = first p2 < N with chs < cw[p2], else N (src/hmm_resample.jl:9-16)
__global__ void __launch_bounds__(kBlockThreads) pf_multinomial_gather_kernel(const __grid_constant__ ResampleLaunch a, int tile_size) {
    const long long gi = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
    const int b = (int)(gi / a.n_pad);
    const long long i = gi % a.n_pad;
    if (b >= a.n_filters || i >= a.n) return;
    const uint32_t gfilter = a.filter_ids ? a.filter_ids[b] : a.filter0 + (uint32_t)b;
    const double big_s = a.filt_s[b];
    const Philox4 p = stream_draw(a.key, (uint32_t)i, gfilter, (uint32_t)a.t, kTagResample, 1u);
    const double chs = __dmul_rn(u53(p.w0, p.w1), big_s);
    const double* off = a.tile_off + (size_t)b * (a.ntiles + 1);
    int lt = 0, ht = a.ntiles;  // first tile with chs < off[tile + 1]
    while (lt < ht) {
        const int mid = (lt + ht) >> 1;
        if (chs < off[mid + 1]) ht = mid; else lt = mid + 1;
    }
    long long res = a.n - 1;
    if (lt < a.ntiles) {
        const long long base = (long long)lt * tile_size;
        const long long rem = a.n - base;
        const int nvalid = rem < tile_size ? (int)rem : tile_size;
        const double* cw = a.cw + (size_t)b * a.n_pad + base;
        int lq = 0, hq = nvalid;
        while (lq < hq) {
            const int mid = (lq + hq) >> 1;
            if (chs < cw[mid]) hq = mid; else lq = mid + 1;
        }
        if (lq >= nvalid) lq = nvalid - 1;
        res = base + lq;
    }
    const int32_t* src_b = a.pop_src + (size_t)b * a.n_comp * a.n_pad;
    int32_t* dst_b = a.pop_dst + (size_t)b * a.n_comp * a.n_pad;
    for (int c = 0; c < a.n_comp; ++c) dst_b[(size_t)c * a.n_pad + i] = src_b[(size_t)c * a.n_pad + res];
    if (a.anc) a.anc[(size_t)b * a.n_pad + i] = (int32_t)res;
}

cudaError_t launch_resample(int items, const ResampleLaunch& a, cudaStream_t stream) {
    const unsigned grid = (unsigned)(a.n_filters * a.ntiles);
    if (items == 1) pf_resample_kernel<1><<<grid, kBlockThreads, 0, stream>>>(a);
    else pf_resample_kernel<4><<<grid, kBlockThreads, 0, stream>>>(a);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (a.rs_type == DPOMP_RS_MULTINOMIAL) {
        const long long total = (long long)a.n_filters * a.n_pad;
        pf_multinomial_gather_kernel<<<(unsigned)((total + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, stream>>>(
            a, kBlockThreads * items);
        err = cudaGetLastError();
    }
    return err;
}

// ------------------------------------------------------------------------------------------------------------
// Whole-filter copies for the outer layer: dst filter dst_slots[k] <- src filter src_slots[k] (1-based slots).
// 128-bit vectorised; filter_stride_words = C * n_pad is a multiple of 256.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) gather_filters_kernel(int32_t* dst, const int32_t* src, const int64_t* dst_slots,
                                                                        const int64_t* src_slots, long long stride_words) {
    const int k = blockIdx.y;
    const long long d = dst_slots ? dst_slots[k] - 1 : k;
    const long long s = src_slots ? src_slots[k] - 1 : k;
    const int4* sp = reinterpret_cast<const int4*>(src + s * stride_words);
    int4* dp = reinterpret_cast<int4*>(dst + d * stride_words);
    const long long nvec = stride_words / 4;
    for (long long v = (long long)blockIdx.x * kBlockThreads + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlockThreads)
        dp[v] = sp[v];
}

cudaError_t launch_gather_filters(int32_t* dst, const int32_t* src, const int64_t* dst_slots_dev, const int64_t* src_slots_dev,
                                  int n, long long filter_stride_words, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const long long nvec = filter_stride_words / 4;
    long long bx = (nvec + kBlockThreads - 1) / kBlockThreads;
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    for (int k0 = 0; k0 < n; k0 += 65535) {  // gridDim.y limit
        const int cnt = (n - k0) < 65535 ? (n - k0) : 65535;
        gather_filters_kernel<<<dim3((unsigned)bx, (unsigned)cnt), kBlockThreads, 0, stream>>>(
            dst, src, dst_slots_dev ? dst_slots_dev + k0 : nullptr, src_slots_dev ? src_slots_dev + k0 : nullptr,
            filter_stride_words);
        if (!dst_slots_dev || !src_slots_dev) {
            // identity side: offset the dense buffer by k0 filters
            if (!dst_slots_dev) dst += (long long)cnt * filter_stride_words;
            if (!src_slots_dev) src += (long long)cnt * filter_stride_words;
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_pack_filters(int32_t* packed, const int32_t* pop, const int64_t* slots_dev, int n,
                                long long filter_stride_words, int unpack, cudaStream_t stream) {
    if (unpack) return launch_gather_filters(const_cast<int32_t*>(pop), packed, slots_dev, nullptr, n, filter_stride_words, stream);
    return launch_gather_filters(packed, pop, nullptr, slots_dev, n, filter_stride_words, stream);
}

// ------------------------------------------------------------------------------------------------------------
// Bit-exactness hook: literal searches on a given cumulative-weight array (1-based output).
//   systematic / stratified: first j with NOT (u_i > cw_j)   (walk of src/hmm_pf_resample.jl:34-40 as a lower bound;
//                            u_i is non-decreasing so the walk and the bound agree)
//   multinomial:             first p2 < n with chs < cw[p2], else n  (src/hmm_resample.jl:9-16)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) search_hook_kernel(int rs_type, const double* cw, long long n, const double* u,
                                                                     long long n_out, int64_t* out) {
    const long long i = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
    if (i >= n_out) return;
    const double s = cw[n - 1];
    const double dn = (double)n;
    long long lo = 0, hi = n - 1;  // answer in [0, n-1]
    if (rs_type == DPOMP_RS_MULTINOMIAL) {
        const double chs = __dmul_rn(u[i], s);
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (chs < cw[mid]) hi = mid; else lo = mid + 1;
        }
    } else {
        const double r = (rs_type == DPOMP_RS_STRATIFIED) ? u[i] : u[0];
        const double ui = __dmul_rn(__dadd_rn(__ddiv_rn(r, dn), __ddiv_rn((double)i, dn)), s);
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (!(ui > cw[mid])) hi = mid; else lo = mid + 1;
        }
    }
    out[i] = lo + 1;
}

cudaError_t launch_search_hook(int rs_type, const double* cw_dev, long long n, const double* u_dev, long long n_out,
                               int64_t* out_dev, cudaStream_t stream) {
    search_hook_kernel<<<(unsigned)((n_out + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, stream>>>(
        rs_type, cw_dev, n, u_dev, n_out, out_dev);
    return cudaGetLastError();
}

}  // namespace dpomp
