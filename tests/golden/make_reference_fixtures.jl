# Generates REFERENCE-held parity fixtures: tests/golden/ref_<name>.txt from tests/golden/ref_inputs/<name>.txt.
#
#   python tests/golden/export_for_julia.py                                        # writes the inputs (already committed)
#   julia --project=/path/to/gcdyn.jl tests/golden/make_reference_fixtures.jl       # Julia 1.10.2 + the reference's Manifest
#
# UNEXECUTED in the build image (no Julia there).  Every value written here comes from the real gcdyn.jl:
#   loglikelihood(model, tree, present_time; reltol=1e-12, abstol=1e-12)   (src/multitype_branching_process.jl:110-248)
#   gcdyn.naive_loglikelihood / stadler_appx_loglikelihood / stadler_appx_unconditioned_loglikelihood (src/extra_likelihoods.jl)
# on trees rebuilt node for node from the flat export (post-order, parent positions), with every parameter set expressed
# as a DiscreteBranchingProcess (a λ table: the canonical form the C ABI takes for all three model types).
#
# Output format (consumed by tests/test_reference_fixtures.py):
#   gcdyn_ref 1 / name / julia VERSION / P T / then one line per (method, pset): "<method> <p> v_1 ... v_T"
#   with method ∈ {ode, naive, stadler, stadler_uncond}; a method that throws for a case is written as "<method> <p> error".

using gcdyn
using Printf

const EVENT_SYMBOLS = gcdyn.EVENTS          # src/types.jl:6 — the flat event code is the 0-based index into this tuple

struct Case
    name::String
    present_time::Float64
    type_space::Vector{Float64}
    models::Vector{Any}
    trees::Vector{Any}
end

function read_case(path)
    lines = readlines(path)
    i = Ref(1)
    nxt() = (l = split(lines[i[]]); i[] += 1; l)
    @assert nxt() == ["gcdyn_flat", "1"]
    name = String(nxt()[2])
    T = parse(Float64, nxt()[2])
    K = parse(Int, nxt()[2])
    ts = parse.(Float64, nxt()[2:end])
    P = parse(Int, nxt()[2])
    rows = [parse.(Float64, nxt()[3:end]) for _ in 1:P]          # μ δ ρ σ λ_1..λ_K
    shared = parse(Int, nxt()[2]) == 1
    Gs = [permutedims(reshape(parse.(Float64, nxt()[3:end]), K, K)) for _ in 1:(shared ? 1 : P)]   # row-major text → Matrix
    models = Any[]
    for p in 1:P
        r = rows[p]
        Γ = shared ? Gs[1] : Gs[p]
        push!(models, DiscreteBranchingProcess(r[5:end], r[1], r[2], Γ, r[3], r[4], ts))
    end
    ntrees = parse(Int, nxt()[2])
    trees = Any[]
    for t in 1:ntrees
        w = nxt()
        @assert w[1] == "tree"
        n = parse(Int, w[3])
        root = TreeNode(:root, parse(Float64, w[5]), ts[parse(Int, w[4]) + 1])
        ev = Vector{Symbol}(undef, n); ty = Vector{Float64}(undef, n); tm = Vector{Float64}(undef, n); par = Vector{Int}(undef, n)
        for k in 1:n
            a = nxt()
            ev[k] = EVENT_SYMBOLS[parse(Int, a[1]) + 1]
            tyi = parse(Int, a[2])
            ty[k] = tyi >= 0 ? ts[tyi + 1] : NaN
            tm[k] = parse(Float64, a[3])
            par[k] = parse(Int, a[4])
        end
        nodes = [TreeNode(ev[k], tm[k], ty[k]) for k in 1:n]
        # children are attached in increasing post-order position = the order of the `children` vector
        # (PostOrderTraversal visits children in vector order, src/treenode.jl:159-175)
        for k in 1:n
            attach!(par[k] < 0 ? root : nodes[par[k] + 1], nodes[k])
        end
        push!(trees, root)
    end
    return Case(name, T, ts, models, trees)
end

function try_eval(f)
    try
        return f()
    catch e
        return e
    end
end

function write_method(io, method, p, vals)
    if any(v -> v isa Exception, vals)
        println(io, method, " ", p - 1, " error")
    else
        println(io, method, " ", p - 1, " ", join((@sprintf("%.17g", v) for v in vals), " "))
    end
end

function main()
    here = @__DIR__
    indir = joinpath(here, "ref_inputs")
    for f in sort(readdir(indir))
        endswith(f, ".txt") || continue
        c = read_case(joinpath(indir, f))
        open(joinpath(here, "ref_" * c.name * ".txt"), "w") do io
            println(io, "gcdyn_ref 1")
            println(io, "name ", c.name)
            println(io, "julia ", VERSION)
            println(io, "sizes ", length(c.models), " ", length(c.trees))
            for (p, m) in enumerate(c.models)
                T = c.present_time
                write_method(io, "ode", p, [try_eval(() -> loglikelihood(m, t, T; reltol=1e-12, abstol=1e-12)) for t in c.trees])
                write_method(io, "naive", p, [try_eval(() -> gcdyn.naive_loglikelihood(m, t)) for t in c.trees])
                write_method(io, "stadler", p, [try_eval(() -> gcdyn.stadler_appx_loglikelihood(m, t, T)) for t in c.trees])
                write_method(io, "stadler_uncond", p, [try_eval(() -> gcdyn.stadler_appx_unconditioned_loglikelihood(m, t, T)) for t in c.trees])
            end
        end
        println("wrote ref_", c.name, ".txt: ", length(c.models), " psets x ", length(c.trees), " trees")
    end
end

main()
