# DiscretePOMPGPU.jl -- the reference-side binding a maintainer would add to DiscretePOMP.jl to route the
# particle-filter hot path through libdpomp.so (include/dpomp.h).  UNTESTED in this repository: there is no Julia
# toolchain in the build image (SURVEY.md F3); the Python host in discretepomp.jl_b200/ replays the same call
# sequence and is what the test-suite exercises.
#
# Usage (inside module DiscretePOMP, after the includes of src/DiscretePOMP.jl:72-106):
#     include("DiscretePOMPGPU.jl")
# Exported signatures (src/DiscretePOMP.jl:59-67) do not change; only the bodies of get_log_pdf_fn /
# estimate_likelihood and the three partial_log_likelihood! call sites of run_pibis are replaced.

const LIBDPOMP = get(ENV, "LIBDPOMP", "libdpomp")
const DPOMP_MAX_C, DPOMP_MAX_E, DPOMP_MAX_V = 8, 8, 8

# struct dpomp_model_desc (include/dpomp.h); NTuple-of-NTuple fields are row-major like the C arrays
struct DpompModelDesc
    n_compartments::Int32; n_events::Int32; n_params::Int32; t0_index::Int32
    rate_par::NTuple{8,Int32}
    rate_f1::NTuple{64,Int32}; rate_k1::NTuple{8,Int32}
    rate_f2::NTuple{64,Int32}; rate_k2::NTuple{8,Int32}
    rate_has_den::NTuple{8,Int32}
    rate_dn::NTuple{64,Int32}; rate_kd::NTuple{8,Int32}
    trans::NTuple{64,Int32}
    initial_condition::NTuple{8,Int64}
    obs_sigma::Float64
    obs_xmask::NTuple{8,Int32}
    n_obs_vals::Int32
    obs_ymask::NTuple{8,Int32}
    n_obs::Int32
    obs_time::Ptr{Float64}; obs_id::Ptr{Int32}; obs_val::Ptr{Int64}
end

dpomp_check(rc) = rc == 0 ? nothing :
    error("libdpomp: ", unsafe_string(ccall((:dpomp_last_error, LIBDPOMP), Cstring, ())))

pad(v, n, T) = ntuple(i -> i <= length(v) ? T(v[i]) : zero(T), n)
rows(m::AbstractMatrix, T) = ntuple(k -> (e = (k - 1) ÷ 8 + 1; c = (k - 1) % 8 + 1;
                                          e <= size(m, 1) && c <= size(m, 2) ? T(m[e, c]) : zero(T)), 64)

# The rate table (par, f1, k1, f2, k2, has_den, dn, kd) and the observation table (sigma, xmask, ymask) are fitted by
# probing model.rate_function / model.obs_model exactly as discretepomp.jl_b200/rate_table.py does (fit_rate_table and
# fit_obs_table below are line-for-line ports of compile_rate_table / compile_obs_table and raise when a closure is
# not representable -- there is no CPU fallback).
function dpomp_model(model::HiddenMarkovModel)
    ic = model.fn_initial_condition()
    E, C = model.n_events, length(ic)
    tm = vcat([reshape(model.fn_transition(e), 1, C) for e in 1:E]...)      # rows = events (note: no transpose needed)
    rt = fit_rate_table(model.rate_function, E, length(model.prior), C)
    ot = fit_obs_table(model.obs_model, C, length(model.obs_data[1].val), length(model.prior))
    times = Float64[y.time for y in model.obs_data]
    ids = Int32[y.obs_id for y in model.obs_data]
    vals = Int64[y.val[v] for y in model.obs_data for v in 1:length(y.val)]  # row t = y[t].val
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve times ids vals begin
        desc = DpompModelDesc(C, E, length(model.prior), model.t0_index,
            pad(rt.par, 8, Int32), rows(rt.f1, Int32), pad(rt.k1, 8, Int32), rows(rt.f2, Int32), pad(rt.k2, 8, Int32),
            pad(rt.has_den, 8, Int32), rows(rt.dn, Int32), pad(rt.kd, 8, Int32), rows(tm, Int32), pad(ic, 8, Int64),
            ot.sigma, pad(ot.xmask, 8, Int32), length(model.obs_data[1].val), pad(ot.ymask, 8, Int32),
            length(times), pointer(times), pointer(ids), pointer(vals))
        dpomp_check(ccall((:dpomp_model_create, LIBDPOMP), Cint, (Ref{DpompModelDesc}, Ref{Ptr{Cvoid}}), desc, handle))
    end
    return handle[]       # the library copies the observations; release with dpomp_model_destroy
end

mutable struct DpompPF
    handle::Ptr{Cvoid}
    n_batch::Int
end
function DpompPF(mdl_handle, n_particles, n_batch, rs_type; seed = rand(UInt64), device = -1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dpomp_check(ccall((:dpomp_pf_create, LIBDPOMP), Cint, (Ptr{Cvoid}, Int64, Int32, Int32, UInt64, Int32, Ref{Ptr{Cvoid}}),
                      mdl_handle, n_particles, n_batch, rs_type, seed, device, h))
    pf = DpompPF(h[], n_batch)
    finalizer(p -> ccall((:dpomp_pf_destroy, LIBDPOMP), Cint, (Ptr{Cvoid},), p.handle), pf)
    return pf
end

## replaces get_log_pdf_fn (src/hmm_particle_filter.jl:87-101): same signature, same closure type
function get_log_pdf_fn(mdl::HiddenMarkovModel, p::Int64 = C_DF_PF_P, rs_type::Int64 = 1; essc::Float64 = C_DF_ESS_CRIT)
    pf = DpompPF(dpomp_model(mdl), p, 1, rs_type in (2, 3) ? rs_type : 1)
    function comp_log_pdf(parameters::Array{Float64, 1})
        out = Ref{Float64}(0.0)
        GC.@preserve parameters dpomp_check(ccall((:dpomp_pf_loglik, LIBDPOMP), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int32, Ref{Float64}), pf.handle, parameters, 1, out))
        return out[]
    end
    return comp_log_pdf
end

## replaces the loop `for p in eachindex(pop): gx[p] = partial_log_likelihood!(pop[p], ...)` (src/hmm_ibis.jl:53-56):
## theta is the n_theta x outer_p matrix of run_pibis, already column-major = one theta vector per filter
function partial_log_likelihood_batch!(pf::DpompPF, theta::Array{Float64, 2}, ymin::Int64, ymax::Int64)
    gx = Array{Float64, 1}(undef, size(theta, 2))
    GC.@preserve theta gx dpomp_check(ccall((:dpomp_pf_partial, LIBDPOMP), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}), pf.handle, theta, size(theta, 2), ymin, ymax, gx))
    return gx
end

## `pop2[p] .= pop[nidx[p]]` (src/hmm_ibis.jl:74) and `pop[p] .= pop_f` (:108)
permute_filters!(pf::DpompPF, nidx::Array{Int64, 1}) =
    dpomp_check(ccall((:dpomp_pf_permute, LIBDPOMP), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int32), pf.handle, nidx, length(nidx)))
copy_filters!(dst::DpompPF, src::DpompPF, dst_slots::Array{Int64, 1}, src_slots::Array{Int64, 1}) =
    dpomp_check(ccall((:dpomp_pf_copy_filters, LIBDPOMP), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int32),
                      dst.handle, src.handle, dst_slots, src_slots, length(dst_slots)))

## replaces rs_systematic (src/hmm_resample.jl:44-62): same return type (1-based Vector{Int64})
function rs_systematic(w::Array{Float64, 1})
    out = Array{Int64, 1}(undef, length(w)); u = [rand()]
    dpomp_check(ccall((:dpomp_resample_indices, LIBDPOMP), Cint,
        (Int32, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int32), 1, 0, w, length(w), u, 1, length(w), out, -1))
    return out
end
