// dpomp_internal.cuh -- host-side handle layouts and the launcher interface between capi.cu and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dpomp.h"
#include "dpomp_dev.cuh"

namespace dpomp {

// Rate / observation table in the arithmetic of the event loop (Real = float or double), padded to the
// instantiated (C, E).  Passed to kernels by value (constant bank).
template <typename Real, int C, int E>
struct DevModel {
    int par[E];       // parameter index or -1
    Real f1[E][C], k1[E];
    Real f2[E][C], k2[E];
    Real dn[E][C], kd[E];
    int has_den[E];
    int any_den;
    Real trans[E][C];
    int xmask_i[C];
    int ic[C];
    double obs_inv_tmp2;
    int obs_tmp2_pow2;
    double obs_tmp1, obs_tmp2;  // log(1/(sqrt(2 pi) sigma)), 2 sigma^2 (src/hmm_examples.jl:61-62)
    int t0_index;
    int n_params;
};

// Everything one launch of the simulate+weight kernel needs besides the model.
struct SimLaunch {
    int32_t* pop;            // [B][C][n_pad] current populations (read unless fresh, written)
    double* logw;            // [B][n_pad] log weights (diagnostics; written only when record_logw)
    double* wtile;           // [B][n_pad] tile-local inclusive scan of exp(logw - tile max) (deterministic tree)
    int record_logw;
    const double* theta;     // [B][n_params] device
    const double* obs_time;  // [T]
    const double* obs_ysum;  // [T]  sum_v ymask[v] * y.val[v]
    double* tile_m;          // [B][ntiles]
    double* tile_s;          // [B][ntiles]
    double* tile_f;          // [B][ntiles]  f_{b|g} = exp(m_b - m_g): tile scale inside its group of kGroupTiles tiles
    double* tile_off;        // [B][ntiles]  o_{b|g}: exclusive offset of the tile inside its group (group scale)
    double* grp_m;           // [B][ngroups] group maxima / sums / scales F_g = exp(m_g - M) / exclusive offsets O_g
    double* grp_s;
    double* grp_f;
    double* grp_off;
    unsigned int* grp_counter;  // [B][ngroups] tickets of the tiles of a group
    int ngroups;
    double* filt_m;          // [B]
    double* filt_s;          // [B]
    double* ll_acc;          // [B]
    unsigned int* tile_counter;      // [B]
    unsigned long long* ev_count;    // [1]
    unsigned long long* ovf_count;   // [1]
    long long n;             // particles per filter
    long long n_pad;         // ntiles * tile
    int ntiles;
    int n_filters;           // filters in this launch
    int n_comp;              // real C (<= instantiated C)
    int t;                   // 0-based observation index
    int fresh;               // 1: start from the initial condition at t_prev = 0 / theta[t0_index]
    int has_lik;             // obs_id[t] > 0
    // fused step (simulate + resample in ONE launch; all tiles of a filter must be co-resident): see pf_sim.cuh
    int do_resample;         // fused kernel only: resample after the combine
    int rs_type;
    int32_t* pop_dst;        // [B][C][n_pad] offspring populations
    int32_t* anc;            // [B][n_pad] 0-based ancestors or nullptr
    ChunkPerm perm;          // offspring placement (identity or chunk-interleaved)
    unsigned long long* work_counter;  // dynamic logical CTA index (arrival order) = atomicAdd(work_counter) - work_base
    unsigned long long work_base;
    unsigned int* filt_gen;  // [B] set to `gen` by the CTA that finished the filter's combine
    unsigned int gen;
    uint64_t key;
    uint32_t filter0;        // global id of local filter 0
    const uint32_t* filter_ids;  // optional explicit global ids [n_filters] (overrides filter0 + b)
    long long max_events;
};

struct ResampleLaunch {
    const int32_t* pop_src;  // [B][C][n_pad]
    int32_t* pop_dst;
    const double* wtile;     // [B][n_pad] tile-local inclusive scans written by the simulate kernel
    const double* tile_m;
    const double* tile_f;
    const double* tile_off;
    const double* grp_f;
    const double* grp_off;
    int ngroups;
    const double* filt_s;
    int32_t* anc;            // [B][n_pad] 0-based ancestors, or nullptr
    double* cw;              // [B][n_pad] cumulative weights (multinomial only), or nullptr
    long long n, n_pad;
    int ntiles, n_filters, n_comp;
    int t, rs_type;
    ChunkPerm perm;          // offspring placement (identity or chunk-interleaved)
    uint64_t key;
    uint32_t filter0;
    const uint32_t* filter_ids;
};

struct ModelHost {
    dpomp_model_desc desc;  // pointers re-targeted to the vectors below
    std::vector<double> obs_time;
    std::vector<int32_t> obs_id;
    std::vector<int64_t> obs_val;
    std::vector<double> obs_ysum;
};

// launchers implemented in the kernel TUs; all asynchronous on `stream`; return cudaGetLastError()
cudaError_t launch_sim_weight(const ModelHost& m, int sim_precision, int items, int fused, const SimLaunch& a, cudaStream_t stream);
// co-resident CTAs of the fused step kernel for this model / geometry on the current device (0 = unavailable)
int sim_fused_capacity(const ModelHost& m, int sim_precision, int items);
int builtin_model_id(const dpomp_model_desc& d);  // 0 = generic rate table, > 0 = hand-specialised predefined model
cudaError_t launch_resample(int items, const ResampleLaunch& a, cudaStream_t stream);
cudaError_t launch_gather_filters(int32_t* dst, const int32_t* src, const int64_t* dst_slots_dev,
                                  const int64_t* src_slots_dev, int n, long long filter_stride_words, cudaStream_t stream);
cudaError_t launch_pack_filters(int32_t* dst_packed, const int32_t* pop, const int64_t* slots_dev, int n,
                                long long filter_stride_words, int unpack, cudaStream_t stream);
cudaError_t launch_search_hook(int rs_type, const double* cw_dev, long long n, const double* u_dev, long long n_out,
                               int64_t* out_dev, cudaStream_t stream);
int sim_kernel_supported(int n_comp, int n_events);

}  // namespace dpomp
