# make_reference_kats.jl -- dumps FUNCTION-LEVEL known-answer vectors from the UNMODIFIED DiscretePOMP.jl package into
# tests/golden/ref_kats/ref_kats.json, which tests/test_reference_kats.py checks the oracle AND the CUDA path against.
# This is the route from "parity unpinned" (no Julia in the build image, SURVEY.md F3) to pinned parity:
#
#     julia --project=/path/to/DiscretePOMP.jl baseline/make_reference_kats.jl [tests/golden/ref_kats]
#
# The reference draws from Julia's global RNG inside its functions, so every randomised vector is produced with a seeded
# shim: seed, record the draws the function is about to consume (same call shape: scalar rand() or rand(n)), re-seed, call
# the function.  Nothing here depends on the Julia version's RNG algorithm: the recorded draws travel with the answers.
# Anchors: test/runtests.jl:7-52 (what the package's own test-suite exercises), src/hmm_resample.jl, src/hmm_pf_resample.jl,
# src/hmm_cmn.jl:4-10, src/hmm_examples.jl:59-67,103-168, src/hmm_particle_filter.jl:4-6, src/cmn.jl:91-99.
import DiscretePOMP
import Random
using Printf
const D = DiscretePOMP

outdir = length(ARGS) > 0 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "ref_kats")
mkpath(outdir)

fnum(x::Float64) = isfinite(x) ? @sprintf("%.17g", x) : (isnan(x) ? "\"nan\"" : (x > 0 ? "\"inf\"" : "\"-inf\""))
fnum(x::Integer) = string(x)
jarr(v) = "[" * join((fnum(x) for x in v), ",") * "]"
jstr(s) = "\"" * s * "\""
jobj(pairs) = "{" * join((jstr(k) * ":" * v for (k, v) in pairs), ",") * "}"

records = Dict{String, Vector{String}}()
add!(key, rec) = push!(get!(records, key, String[]), rec)

# ---- outer resamplers on RAW weights (src/hmm_resample.jl:4-20, 44-62, 66-83) ------------------------------------------
weight_sets = Vector{Vector{Float64}}()
push!(weight_sets, [1.0, 1.0, 1.0, 1.0])
push!(weight_sets, [0.0, 0.0, 1.0, 0.0])
push!(weight_sets, [3.0, 1.0])
push!(weight_sets, [0.25, 0.25, 0.25, 0.25, 0.0, 0.0, 1.0e-300, 2.0])           # ties, zeros, tiny
Random.seed!(101)
for n in (7, 64, 1000, 4097)
    push!(weight_sets, rand(n))
    push!(weight_sets, exp.(-20.0 .* rand(n)))                                   # many orders of magnitude
end
for (k, w) in enumerate(weight_sets), s in (1, 2, 3)
    n = length(w)
    Random.seed!(1000 * k + s); u = [rand()]
    Random.seed!(1000 * k + s); idx = D.rs_systematic(copy(w))
    add!("rs_systematic", jobj(["w" => jarr(w), "u" => jarr(u), "idx" => jarr(idx)]))
    Random.seed!(2000 * k + s); u = rand(n)                                      # rs_stratified calls rand(length(w))
    Random.seed!(2000 * k + s); idx = D.rs_stratified(copy(w))
    add!("rs_stratified", jobj(["w" => jarr(w), "u" => jarr(u), "idx" => jarr(idx)]))
    if n <= 1000                                                                 # O(N^2)
        Random.seed!(3000 * k + s); u = [rand() for _ in 1:n]                    # one scalar rand() per offspring
        Random.seed!(3000 * k + s); idx = D.rs_multinomial(copy(w))
        add!("rs_multinomial", jobj(["w" => jarr(w), "u" => jarr(u), "idx" => jarr(idx)]))
    end
end

# ---- in-filter systematic resampler on CUMULATIVE weights (src/hmm_pf_resample.jl:24-42) -------------------------------
for (k, w) in enumerate(weight_sets), s in (1, 2)
    n = length(w)
    cw = cumsum(w)
    old_p = reshape(collect(Int64, 1:n), n, 1)                                   # row j holds j: the copied rows ARE the ancestors
    m_pop = zeros(Int64, n, 1)
    Random.seed!(4000 * k + s); u = [rand()]
    Random.seed!(4000 * k + s); D.rsp_systematic(m_pop, old_p, cw)
    add!("rsp_systematic", jobj(["cw" => jarr(cw), "u" => jarr(u), "idx" => jarr(m_pop[:, 1])]))
end

# ---- choose_event (src/hmm_cmn.jl:4-10) --------------------------------------------------------------------------------
Random.seed!(7)
for k in 1:200
    E = rand(1:6)
    rates = rand(E) .* (rand(E) .> 0.25)                                         # some zero rates
    sum(rates) == 0.0 && (rates[1] = 1.0)
    cum = cumsum(rates)
    Random.seed!(5000 + k); u = rand()
    Random.seed!(5000 + k); ev = D.choose_event(copy(cum))
    add!("choose_event", jobj(["cum" => jarr(cum), "u" => fnum(u), "event" => fnum(ev)]))
end

# ---- Gaussian observation model gom2 (src/hmm_examples.jl:59-67) -------------------------------------------------------
for (sigma, seq) in ((2.0, 2), (1.0, 2), (2.0, 3), (0.5, 1), (3.7, 2))
    g = D.partial_gaussian_obs_model(sigma; seq = seq)
    for (yv, xv) in (([0, 18, 0], [83, 18, 0]), ([0, 65, 7], [30, 71, 0]), ([5, 0, 3], [100, 1, 0]), ([0, 1000, 4], [1, 2, 3000]))
        y = D.Observation(20.0, 1, 1.0, Int64.(yv))
        add!("gom2", jobj(["sigma" => fnum(sigma), "seq" => fnum(seq), "y" => jarr(yv), "x" => jarr(xv),
                           "value" => fnum(g(y, Int64.(xv), [0.1, 0.1, 0.1]))]))
    end
end

# ---- predefined rate functions on a grid of states (src/hmm_examples.jl:103-168) ---------------------------------------
models = [("SI", [100, 1]), ("SIR", [100, 1, 0]), ("SIS", [100, 1]), ("SEI", [100, 0, 1]), ("SEIR", [100, 0, 1, 0]),
          ("SEIS", [100, 0, 1]), ("LOTKA", [70, 70]), ("ROSSMAC", [100, 1, 400, 0])]
Random.seed!(11)
for (name, ic) in models, freq_dep in (false, true)
    (freq_dep && name in ("LOTKA", "ROSSMAC")) && continue
    m = D.generate_model(name, Int64.(ic); freq_dep = freq_dep)
    m === nothing && continue
    E = size(m.m_transition, 1)
    for k in 1:12
        theta = rand(max(E, 3)) .* [0.01, 0.5, 0.3, 0.2, 0.1, 0.05][1:max(E, 3)]
        x = Int64[rand(1:300) for _ in ic]
        out = zeros(Float64, E)
        m.rate_function(out, theta, x)
        add!("rates", jobj(["model" => jstr(name), "freq_dep" => (freq_dep ? "1" : "0"), "ic" => jarr(ic), "theta" => jarr(theta),
                            "x" => jarr(x), "rates" => jarr(out), "trans" => jarr(vec(permutedims(m.m_transition)))]))
    end
end

# ---- compute_ess / compute_is_mu_covar! (src/hmm_particle_filter.jl:4-6, src/cmn.jl:91-99) -----------------------------
Random.seed!(13)
for n in (5, 200, 4000)
    w = rand(n); theta = rand(2, n)
    mu = zeros(2); cv = zeros(2, 2)
    D.compute_is_mu_covar!(mu, cv, theta, w)
    add!("moments", jobj(["w" => jarr(w), "theta" => jarr(vec(theta)), "ess" => fnum(D.compute_ess(w)), "mu" => jarr(mu), "cv" => jarr(vec(cv))]))
end

# ---- a particle filter value that does not depend on the RNG: theta = 0 means no events, every particle keeps the initial
#      condition, so partial_log_likelihood! is sum_t log(exp(gom2(y_t, ic))) (src/hmm_particle_filter.jl:39-76) --------------
let
    model = D.generate_model("SIS", [100, 1])
    y = D.get_observations(joinpath(@__DIR__, "..", "tests", "golden", "pooley.csv"))
    mdl = D.get_private_model(model, y)
    for np in (1, 8, 200)
        ll = D.estimate_likelihood(mdl, [0.0, 0.0], np, 2, D.rsp_systematic, 0.3)
        add!("pf_zero_rate", jobj(["model" => jstr("SIS"), "ic" => jarr([100, 1]), "np" => fnum(np), "loglik" => fnum(ll)]))
    end
end

open(joinpath(outdir, "ref_kats.json"), "w") do io
    parts = [jstr(k) * ":[" * join(v, ",") * "]" for (k, v) in sort(collect(records); by = first)]
    push!(parts, jstr("julia_version") * ":" * jstr(string(VERSION)))
    print(io, "{" * join(parts, ",\n") * "}\n")
end
println("wrote ", joinpath(outdir, "ref_kats.json"), " (", sum(length(v) for v in values(records)), " vectors, Julia ", VERSION, ")")
