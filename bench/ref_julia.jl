# bench/ref_julia.jl — times the REAL reference on BASELINE config 3 for readers who have Julia.
#
# UNEXECUTED in this build: the build image and the GPU boxes have no Julia (SURVEY.md §8c/d), so the CPU arm of
# bench.py times oracle/refcpu.c, a C restatement of the same algorithm, instead.  Run this from a checkout of
# thanasibakis/gcdyn.jl with its pinned Manifest (Julia 1.10.2, OrdinaryDiffEq 6.72.0):
#
#     julia --project=/path/to/gcdyn.jl -t 1 bench/ref_julia.jl [n_trees] [n_psets]
#
# It draws the config-3 workload with the reference's own simulator (Julia's RNG stream differs from the NumPy/
# xoshiro streams of this repository, so the trees are statistically — not bitwise — the same), evaluates
# `loglikelihood(model, trees, present_time)` for every parameter set serially, as the reference does
# (src/multitype_branching_process.jl:250-259), and prints one JSON line in bench.py's format.
using gcdyn, Random, Printf

n_trees = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 1000
n_psets = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 4          # the full 256 take hours on one thread
K, present_time = 20, 4.0
type_space = collect(range(-2, 2, K))
base = SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, type_space)   # γ-scalar constructor, types.jl:97-103
Random.seed!(3)
trees = rand_tree(base, present_time, type_space[10], n_trees)
jitter() = exp(0.2 * randn())
models = [SigmoidalBranchingProcess(base.λ_xscale * jitter(), base.λ_xshift + 0.2 * randn(), base.λ_yscale * jitter(),
                                    base.λ_yshift * jitter(), base.μ * jitter(), base.δ * jitter(), base.Γ, base.ρ, base.σ,
                                    type_space) for _ in 1:n_psets]
loglikelihood(models[1], trees[1:min(10, end)], present_time)             # compile
t = @elapsed lls = [loglikelihood(m, trees, present_time) for m in models]
@printf("{\"impl\": \"reference-julia\", \"metric\": \"tree_loglik_evals_per_sec\", \"value\": %.6g, \"unit\": \"(tree,pset) evals/s\", \"threads\": %d, \"n_trees\": %d, \"n_psets\": %d, \"seconds\": %.3f, \"reltol\": 1e-3, \"abstol\": 1e-3}\n",
        n_trees * n_psets / t, Threads.nthreads(), n_trees, n_psets, t)
