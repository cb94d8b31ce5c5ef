# Times the ACTUAL reference (mjb3/DiscretePOMP.jl, unmodified, single-threaded Julia) on the fixtures of this repo, for the
# metric of bench.py: particle-observation steps / s of one bootstrap particle-filter log-likelihood evaluation
# (get_particle_filter_lpdf, src/hmm_utils.jl:281-284 -> estimate_likelihood, src/hmm_particle_filter.jl:79-84).
#
#   julia --project=/path/to/DiscretePOMP.jl baseline/run_reference.jl [repo_root] [np_c2]
#
# Julia is not installed in the build image nor on the GPU box (SURVEY.md F3), so this script has never been run here;
# bench.py --impl reference times the C restatement of the same algorithm (oracle/) instead and says so in its JSON line.
# Run it wherever Julia and the package are available to put the real reference number next to bench.py's.
using DiscretePOMP
import Random

root = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..")
np_c2 = length(ARGS) >= 2 ? parse(Int64, ARGS[2]) : 2^16   # the full C2 size is 2^20; the reference needs ~minutes for it

function time_lpdf(name, model, y, theta, np; reps = 3)
    f = get_particle_filter_lpdf(model, y; np = np)
    f(theta)                                  # compile
    best = Inf
    ll = 0.0
    for _ in 1:reps
        t0 = time_ns()
        ll = f(theta)
        best = min(best, (time_ns() - t0) / 1e9)
    end
    println("{\"impl\": \"reference-julia\", \"config\": \"", name, "\", \"particles\": ", np, ", \"observations\": ", length(y),
            ", \"seconds\": ", best, ", \"value\": ", np * length(y) / best, ", \"unit\": \"particle-observation steps/s\", \"loglik\": ", ll,
            ", \"threads\": 1}")
end

Random.seed!(1)
# C1: SIS on data/pooley.csv (test/runtests.jl:12-18), default particle count
y1 = get_observations(joinpath(root, "tests", "golden", "pooley.csv"))
time_lpdf("C1 SIS pooley", generate_model("SIS", [100, 1]), y1, [0.003, 0.1], 200; reps = 20)
# C2: SIR [100,1,0], theta = (0.003, 0.1), 100 observations (tests/golden/sir_c2.csv), systematic resampling every step
y2 = get_observations(joinpath(root, "tests", "golden", "sir_c2.csv"))
time_lpdf("C2 SIR synthetic", generate_model("SIR", [100, 1, 0]), y2, [0.003, 0.1], np_c2)
