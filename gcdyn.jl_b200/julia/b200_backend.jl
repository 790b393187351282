# b200_backend.jl — Julia side of the drop-in.  NOT EXECUTED in the build environment (no Julia there);
# exercised only through the Python ctypes mirror in ../gcdyn, which drives the same C ABI.
#
# Use: in gcdyn.jl's `src/gcdyn.jl`, after `include("extra_likelihoods.jl")`, add
#     include("b200_backend.jl")
# with `ENV["GCDYN_B200_LIB"]` pointing at `libgcdyn_b200.so`.  The methods below are more specific than
# nothing in the reference — they REPLACE `StatsAPI.loglikelihood(model, tree(s), present_time; …)`
# (src/multitype_branching_process.jl:110-117, 250-257) and the three alternates
# (src/extra_likelihoods.jl:10, 45, 115) by redefinition, keeping names, argument meaning, keyword
# arguments and error behaviour.  Types, traversals, rate functions and the simulator stay as they are.

const LIBGCDYN = get(ENV, "GCDYN_B200_LIB", "libgcdyn_b200.so")

const GCDYN_METHOD_ODE, GCDYN_METHOD_NAIVE, GCDYN_METHOD_STADLER, GCDYN_METHOD_STADLER_UNCOND = Cint(0), Cint(1), Cint(2), Cint(3)
const GCDYN_STATUS_SELF_LOOP, GCDYN_STATUS_SOLVER, GCDYN_STATUS_DOMAIN = Int32(1), Int32(2), Int32(4)

# ---- mirrors of the C structs in include/gcdyn_b200.h (field order and types must match) -----------------

struct GcdynForestDesc
    n_trees::Int64
    n_nodes::Int64
    n_types::Int32
    reserved::Int32
    tree_off::Ptr{Int64}
    event::Ptr{UInt8}
    type_idx::Ptr{Int32}
    btype_idx::Ptr{Int32}
    t_start::Ptr{Float64}
    t_end::Ptr{Float64}
    parent::Ptr{Int32}
    child_off::Ptr{Int64}
    child_idx::Ptr{Int32}
    live::Ptr{UInt8}
    root_type_idx::Ptr{Int32}
    root_time::Ptr{Float64}
    self_loop::Ptr{UInt8}
end

struct GcdynParams
    n_psets::Int32
    n_types::Int32
    response::Int32
    gamma_pset_stride::Int32
    lambda::Ptr{Float64}
    xscale::Ptr{Float64}
    xshift::Ptr{Float64}
    yscale::Ptr{Float64}
    yshift::Ptr{Float64}
    type_space::Ptr{Float64}
    mu::Ptr{Float64}
    delta::Ptr{Float64}
    rho::Ptr{Float64}
    sigma::Ptr{Float64}
    Gamma::Ptr{Float64}
end

# max_cells: the reference's `maxiters` — cap on the adaptive cells of a trajectory (0 = none)
struct GcdynOpts
    steps::Int32
    order::Int32
    tol::Float64
    flags::Int32
    max_cells::Int32
end

# ---- handles -----------------------------------------------------------------------------------------------

mutable struct B200Context
    handle::Ptr{Cvoid}
    function B200Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:gcdyn_ctx_create, LIBGCDYN), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        rc == 0 || error("libgcdyn_b200: " * unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[])
        finalizer(c -> ccall((:gcdyn_ctx_destroy, LIBGCDYN), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

const DEFAULT_CONTEXT = Ref{Union{B200Context, Nothing}}(nothing)
b200_context() = something(DEFAULT_CONTEXT[], (DEFAULT_CONTEXT[] = B200Context(0)))

b200_check(ctx, rc) = rc == 0 || error("libgcdyn_b200 ($rc): " *
    unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), ctx.handle)))

# ---- flattener: TreeNode trees -> post-order struct-of-arrays ---------------------------------------------

"Struct-of-arrays of a forest, nodes in `PostOrderTraversal(tree.children[1])` order (treenode.jl:159-175)."
struct FlatForest
    n_types::Int
    tree_off::Vector{Int64}
    event::Vector{UInt8}
    type_idx::Vector{Int32}
    btype_idx::Vector{Int32}
    t_start::Vector{Float64}
    t_end::Vector{Float64}
    parent::Vector{Int32}
    child_off::Vector{Int64}
    child_idx::Vector{Int32}
    live::Vector{UInt8}
    root_type_idx::Vector{Int32}
    root_time::Vector{Float64}
    self_loop::Vector{UInt8}
    present_time::Float64          # the present time the leaves were validated against (NaN: not an ODE/Stadler flattening)
end

const EVENT_CODE = Dict(e => UInt8(i - 1) for (i, e) in enumerate(EVENTS))     # types.jl:6, 0-based codes

index_of(type_space, t) = (i = findfirst(==(t), type_space); isnothing(i) ? Int32(-1) : Int32(i - 1))   # mbp.jl:17

function flatten(trees::Vector{<:TreeNode}, model::AbstractBranchingProcess, present_time; method = :ode)
    ts = model.type_space
    f = FlatForest(length(ts), Int64[0], UInt8[], Int32[], Int32[], Float64[], Float64[], Int32[], Int64[0], Int32[],
                   UInt8[], Int32[], Float64[], UInt8[], method == :naive ? NaN : Float64(present_time))
    for tree in trees
        if method == :ode                                   # pass A of the reference, mbp.jl:124-164
            for leaf in LeafTraversal(tree)
                if leaf.event == :sampled_survival
                    leaf.time != present_time && throw(ArgumentError("Sampled survival events must all be at the present time."))
                elseif leaf.event == :sampled_death
                    leaf.time == present_time && throw(ArgumentError("Sampled death events must not be at the present time."))
                else
                    throw(ArgumentError("Leaf event must be either `:sampled_survival` or `:sampled_death`."))
                end
            end
        end
        nodes = collect(PostOrderTraversal(tree.children[1]))   # BoundsError on a stub root, like the reference
        pos = IdDict(n => Int32(i - 1) for (i, n) in enumerate(nodes))
        base = length(f.event)
        looped = false
        dead = Base.IdSet{Any}()
        for n in nodes
            ev = EVENT_CODE[n.event]
            bx, tx = index_of(ts, n.up.type), index_of(ts, n.type)
            if !looped && method == :ode                    # pass B checks, mbp.jl:166-194, in the same order
                if n.event == :sampled_survival || n.event == :sampled_death
                    n.event == :sampled_death && bx < 0 && throw(ArgumentError("The type must be in the type space of the model."))
                    isempty(n.children) || throw(KeyError(n))
                elseif n.event == :birth
                    model isa DiscreteBranchingProcess && bx < 0 && throw(ArgumentError("The type must be in the type space of the model."))
                    length(n.children) < 2 && throw(BoundsError(n.children, 2))
                elseif n.event == :type_change
                    if n.type == n.up.type
                        looped = true
                    elseif bx < 0 || tx < 0
                        throw(ArgumentError("The types must be in the type space of the model."))
                    end
                else
                    throw(ArgumentError("Unknown event type $(n.event)"))
                end
                !looped && bx < 0 && throw(ArgumentError("The type must be in the type space of the model."))
            elseif !looped
                bx < 0 && throw(ArgumentError("The type must be in the type space of the model."))
                ok = method == :naive ? (:birth, :sampled_death, :type_change, :sampled_survival, :unsampled_survival) :
                                        (:birth, :sampled_death, :type_change, :sampled_survival)
                n.event in ok || throw(ArgumentError("Tree contains incompatible event $(n.event)"))
                n.event == :type_change && n.type != n.up.type && tx < 0 && throw(ArgumentError("The types must be in the type space of the model."))
            end
            push!(f.event, ev); push!(f.type_idx, tx); push!(f.btype_idx, bx)
            push!(f.t_start, n.up.time); push!(f.t_end, n.time)
            push!(f.parent, get(pos, n.up, Int32(-1)))
            for c in n.children
                push!(f.child_idx, pos[c])
            end
            push!(f.child_off, length(f.child_idx))
            # children the ODE likelihood computes but never uses (mbp.jl:173-191)
            n.event == :birth && foreach(c -> push!(dead, c), n.children[3:end])
            n.event == :type_change && foreach(c -> push!(dead, c), n.children[2:end])
        end
        flags = ones(UInt8, length(nodes))
        if !isempty(dead)
            for i in length(nodes):-1:1                     # parents before children
                par = f.parent[base + i]
                (nodes[i] in dead || (par >= 0 && flags[par + 1] == 0)) && (flags[i] = 0)
            end
        end
        append!(f.live, flags)
        push!(f.root_type_idx, index_of(ts, tree.type)); push!(f.root_time, tree.time)
        push!(f.self_loop, looped ? 0x01 : 0x00)
        push!(f.tree_off, length(f.event))
    end
    return f
end

mutable struct B200Forest
    ctx::B200Context
    flat::FlatForest
    handle::Ptr{Cvoid}
end

function B200Forest(flat::FlatForest; ctx = b200_context())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve flat begin
        desc = GcdynForestDesc(length(flat.tree_off) - 1, length(flat.event), flat.n_types, 0,
            pointer(flat.tree_off), pointer(flat.event), pointer(flat.type_idx), pointer(flat.btype_idx),
            pointer(flat.t_start), pointer(flat.t_end), pointer(flat.parent), pointer(flat.child_off),
            pointer(flat.child_idx), pointer(flat.live), pointer(flat.root_type_idx), pointer(flat.root_time),
            pointer(flat.self_loop))
        b200_check(ctx, ccall((:gcdyn_forest_create, LIBGCDYN), Cint, (Ptr{Cvoid}, Ref{GcdynForestDesc}, Ref{Ptr{Cvoid}}),
                              ctx.handle, desc, h))
    end
    forest = B200Forest(ctx, flat, h[])
    finalizer(x -> ccall((:gcdyn_forest_destroy, LIBGCDYN), Cvoid, (Ptr{Cvoid},), x.handle), forest)
    return forest
end

# ---- parameters: a vector of models -> canonical arrays (λ per type evaluated with the reference's own λ) -----

function pack_params(models::Vector{<:AbstractBranchingProcess})
    ts = models[1].type_space
    all(m -> m.type_space == ts, models) || throw(ArgumentError("all models of a batch must share one type_space"))
    K, P = length(ts), length(models)
    lambda = Float64[Float64(λ(m, t)) for m in models for t in ts]             # pset-major [P*K]
    mu, delta = Float64[m.μ for m in models], Float64[m.δ for m in models]
    rho, sigma = Float64[m.ρ for m in models], Float64[m.σ for m in models]
    shared = all(m -> m.Γ == models[1].Γ, models)
    Gamma = shared ? vec(copy(models[1].Γ)) : reduce(vcat, (vec(m.Γ) for m in models))   # column-major as stored
    return (; K, P, lambda, mu, delta, rho, sigma, Gamma, stride = shared ? 0 : K * K)
end

function b200_evaluate(models::Vector{<:AbstractBranchingProcess}, forest::B200Forest, present_time, method::Cint;
                       per_tree = false, tol = 0.0, maxiters = 0)
    # the leaves were validated against the forest's own present time (mbp.jl:125-134): another one would extrapolate silently
    if method != GCDYN_METHOD_NAIVE && !isnan(forest.flat.present_time) && Float64(present_time) != forest.flat.present_time
        throw(ArgumentError("Sampled survival events must all be at the present time."))
    end
    pp = pack_params(models)
    T = length(forest.flat.tree_off) - 1
    out = Vector{Float64}(undef, pp.P)
    rows = per_tree ? Matrix{Float64}(undef, T, pp.P) : Matrix{Float64}(undef, 0, 0)     # column p = pset p
    status = Matrix{Int32}(undef, T, pp.P)
    GC.@preserve pp out rows status begin
        par = GcdynParams(pp.P, pp.K, 0, pp.stride, pointer(pp.lambda), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL,
                          pointer(pp.mu), pointer(pp.delta), pointer(pp.rho), pointer(pp.sigma), pointer(pp.Gamma))
        opts = GcdynOpts(0, 0, tol, 0, Int32(clamp(maxiters, 0, typemax(Int32))))
        b200_check(forest.ctx, ccall((:gcdyn_loglik_batch, LIBGCDYN), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ref{GcdynParams}, Cdouble, Cint, Ref{GcdynOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            forest.ctx.handle, forest.handle, par, Float64(present_time), method, opts, out,
            per_tree ? pointer(rows) : C_NULL, pointer(status)))
    end
    bits = reduce(|, status; init = Int32(0))
    bits & GCDYN_STATUS_DOMAIN != 0 && throw(DomainError(NaN, "log or sqrt of a negative number while evaluating the likelihood"))
    bits & GCDYN_STATUS_SELF_LOOP != 0 && @warn "Self-loop encountered at a type change event in the tree. Density will evaluate to zero."
    bits & GCDYN_STATUS_SOLVER != 0 && @warn "Extinction probabilities could not be integrated. Density will evaluate to zero."
    return out, rows, status
end

# ---- the reference's entry points, re-pointed at the library ---------------------------------------------------

function StatsAPI.loglikelihood(model::AbstractBranchingProcess, trees::Vector{TreeNode{T}}, present_time;
                                reltol = 1e-3, abstol = 1e-3, maxiters = 1e5) where T
    forest = B200Forest(flatten(trees, model, present_time))
    tol = min(reltol, abstol) < 1e-11 ? min(reltol, abstol) : 0.0          # the library is never looser than 1e-11
    return b200_evaluate([model], forest, present_time, GCDYN_METHOD_ODE; tol, maxiters = floor(Int, min(maxiters, 2.0^31 - 1)))[1][1]
end

StatsAPI.loglikelihood(model::AbstractBranchingProcess, tree::TreeNode, present_time; kw...) =
    StatsAPI.loglikelihood(model, [tree], present_time; kw...)

"Batched entry (new): one value per model; `forest` may be reused across calls (upload once per dataset)."
StatsAPI.loglikelihood(models::Vector{<:AbstractBranchingProcess}, forest::B200Forest, present_time) =
    b200_evaluate(models, forest, present_time, GCDYN_METHOD_ODE)[1]
StatsAPI.loglikelihood(models::Vector{<:AbstractBranchingProcess}, trees::Vector{<:TreeNode}, present_time) =
    StatsAPI.loglikelihood(models, B200Forest(flatten(trees, models[1], present_time)), present_time)

"""
    loglikelihood_and_gradient(models::Vector{<:SigmoidalBranchingProcess}, forest, present_time) -> (ll[P], grad[6, P])

Exact derivatives of every model's log-likelihood with respect to its autodiff-able fields
`(λ_xscale, λ_xshift, λ_yscale, λ_yshift, μ, δ)` — the quantities the reference lets ForwardDiff differentiate through its
`T<:Real` structs (src/types.jl:92, src/multitype_branching_process.jl:118-122) — from `gcdyn_loglik_grad_batch`
(forward sensitivities of the trajectory on the GPU).  An `rrule`/`frule` for `loglikelihood` is one line on top of this.
`ConstantBranchingProcess` models give `(λ, μ, δ)` (3 rows).
"""
function loglikelihood_and_gradient(models::Vector{<:AbstractBranchingProcess}, forest::B200Forest, present_time)
    m1 = models[1]
    sigmoid_resp = m1 isa SigmoidalBranchingProcess
    (sigmoid_resp || m1 isa ConstantBranchingProcess) || throw(ArgumentError("gradients need a parametric birth response"))
    all(m -> typeof(m) == typeof(m1), models) || throw(ArgumentError("all models of a batch must be of one type"))
    P, K = length(models), length(m1.type_space)
    mu, delta = Float64[m.μ for m in models], Float64[m.δ for m in models]
    rho, sigma = Float64[m.ρ for m in models], Float64[m.σ for m in models]
    shared = all(m -> m.Γ == m1.Γ, models)
    Gamma = shared ? vec(copy(m1.Γ)) : reduce(vcat, (vec(m.Γ) for m in models))
    ts = Vector{Float64}(m1.type_space)
    xs = sigmoid_resp ? Float64[m.λ_xscale for m in models] : Float64[]
    xsh = sigmoid_resp ? Float64[m.λ_xshift for m in models] : Float64[]
    ys = sigmoid_resp ? Float64[m.λ_yscale for m in models] : Float64[]
    ysh = sigmoid_resp ? Float64[m.λ_yshift for m in models] : Float64[]
    lam = sigmoid_resp ? Float64[] : Float64[m.λ for m in models]
    out = Vector{Float64}(undef, P)
    grad = Matrix{Float64}(undef, 6, P)
    ndirs = Ref{Int32}(0)
    GC.@preserve mu delta rho sigma Gamma ts xs xsh ys ysh lam out grad begin
        par = sigmoid_resp ?
            GcdynParams(P, K, 2, shared ? 0 : K * K, C_NULL, pointer(xs), pointer(xsh), pointer(ys), pointer(ysh), pointer(ts),
                        pointer(mu), pointer(delta), pointer(rho), pointer(sigma), pointer(Gamma)) :
            GcdynParams(P, K, 1, shared ? 0 : K * K, pointer(lam), C_NULL, C_NULL, C_NULL, C_NULL, pointer(ts),
                        pointer(mu), pointer(delta), pointer(rho), pointer(sigma), pointer(Gamma))
        opts = GcdynOpts(0, 0, 0.0, 0, 0)
        b200_check(forest.ctx, ccall((:gcdyn_loglik_grad_batch, LIBGCDYN), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ref{GcdynParams}, Cdouble, Ref{GcdynOpts}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
            forest.ctx.handle, forest.handle, par, Float64(present_time), opts, out, grad, ndirs))
    end
    n = Int(ndirs[])
    return out, reshape(vec(grad)[1:n * P], n, P)
end

naive_loglikelihood(model::AbstractBranchingProcess, tree::TreeNode) =
    b200_evaluate([model], B200Forest(flatten([tree], model, 0.0; method = :naive)), 0.0, GCDYN_METHOD_NAIVE)[1][1]
stadler_appx_loglikelihood(model::AbstractBranchingProcess, tree::TreeNode, present_time) =
    b200_evaluate([model], B200Forest(flatten([tree], model, present_time; method = :stadler)), present_time, GCDYN_METHOD_STADLER)[1][1]
stadler_appx_unconditioned_loglikelihood(model::AbstractBranchingProcess, tree::TreeNode, present_time) =
    b200_evaluate([model], B200Forest(flatten([tree], model, present_time; method = :stadler)), present_time, GCDYN_METHOD_STADLER_UNCOND)[1][1]

# ---- multi-GPU: one process per GPU, trees sharded; the [P] sums are combined by the library's peer-memory kernel ----

"`gcdyn_comm_*`: `allgather_handles(handle::Vector{UInt8}) -> Vector{UInt8}` (world × 64 bytes, rank order) is supplied
by the caller, e.g. `h -> MPI.Allgather(h, comm)`.  `row_bytes > 0` adds the row window of the trajectory sharding
(`gcdyn_comm_rows_create/_connect`): sharded evaluations of ≥ 512 parameter sets then integrate only this rank's slice of
the sets and exchange the coefficient rows over peer memory; `2 * P * (Jst * 8 + 32)` bytes with
`Jst = 2K + 4 + 129K + K²` covers every grid degree."
mutable struct B200Comm
    handle::Ptr{Cvoid}
    function B200Comm(ctx::B200Context, rank::Integer, world::Integer, max_psets::Integer, allgather_handles; row_bytes::Integer = 0)
        mine = Vector{UInt8}(undef, 64)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:gcdyn_comm_create, LIBGCDYN), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
                   ctx.handle, rank, world, max_psets, mine, out)
        rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), ctx.handle)))
        if world > 1
            all = allgather_handles(mine)
            rc = ccall((:gcdyn_comm_connect, LIBGCDYN), Cint, (Ptr{Cvoid}, Ptr{UInt8}), out[], all)
            rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), ctx.handle)))
        end
        if row_bytes > 0
            rc = ccall((:gcdyn_comm_rows_create, LIBGCDYN), Cint, (Ptr{Cvoid}, Int64, Ptr{UInt8}), out[], row_bytes, mine)
            rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), ctx.handle)))
            if world > 1
                all = allgather_handles(mine)
                rc = ccall((:gcdyn_comm_rows_connect, LIBGCDYN), Cint, (Ptr{Cvoid}, Ptr{UInt8}), out[], all)
                rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), ctx.handle)))
            end
        end
        c = new(out[])
        finalizer(x -> ccall((:gcdyn_comm_destroy, LIBGCDYN), Cvoid, (Ptr{Cvoid},), x.handle), c)
        return c
    end
end

"Sharded batched entry: `forest` holds this rank's trees, the result is the sum over all ranks (collective)."
function StatsAPI.loglikelihood(models::Vector{<:AbstractBranchingProcess}, forest::B200Forest, present_time, comm::B200Comm)
    pp = pack_params(models)
    out = Vector{Float64}(undef, pp.P)
    GC.@preserve pp out begin
        par = GcdynParams(pp.P, pp.K, 0, pp.stride, pointer(pp.lambda), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL,
                          pointer(pp.mu), pointer(pp.delta), pointer(pp.rho), pointer(pp.sigma), pointer(pp.Gamma))
        rc = ccall((:gcdyn_loglik_batch_sharded, LIBGCDYN), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ref{GcdynParams}, Cdouble, Cint, Ref{GcdynOpts}, Ptr{Cvoid},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                   forest.ctx.handle, forest.handle, par, Float64(present_time), GCDYN_METHOD_ODE, GcdynOpts(0, 0, 0.0, 0, 0),
                   comm.handle, out, C_NULL, C_NULL)
    end
    rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), forest.ctx.handle)))
    return out
end

# ---- device-resident random-walk Metropolis (gcdyn_mh_*) -----------------------------------------------------------

struct GcdynMhDesc
    n_chains::Int32
    n_types::Int32
    type_space::Ptr{Float64}
    Gamma::Ptr{Float64}
    rho::Float64
    sigma::Float64
    theta0::Ptr{Float64}
    lower::Ptr{Float64}
    upper::Ptr{Float64}
    step::Float64
    seed::UInt64
end

"C chains over θ = (log xscale, xshift, log yscale, log yshift, log μ, log δ) of a `SigmoidalBranchingProcess`;
returns (chain[6, C, iterations], loglik[C, iterations], accepted[C])."
function b200_metropolis(template::SigmoidalBranchingProcess, forest::B200Forest, present_time, theta0::Matrix{Float64},
                         iterations::Integer; step = 0.05, lower, upper, seed = 0, comm = nothing)
    C = size(theta0, 2)                                   # column per chain: Julia [6, C] = C [C][6]
    ts = collect(Float64, template.type_space); G = vec(collect(Float64, template.Γ))
    lo = collect(Float64, lower); up = collect(Float64, upper)
    mh = Ref{Ptr{Cvoid}}(C_NULL)
    chain = Array{Float64}(undef, 6, C, iterations); ll = Matrix{Float64}(undef, C, iterations); acc = Vector{Int64}(undef, C)
    GC.@preserve ts G lo up theta0 begin
        d = GcdynMhDesc(C, length(ts), pointer(ts), pointer(G), template.ρ, template.σ, pointer(theta0), pointer(lo), pointer(up),
                        step, UInt64(seed))
        rc = ccall((:gcdyn_mh_create, LIBGCDYN), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{GcdynMhDesc}, Cdouble, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                   forest.ctx.handle, forest.handle, d, Float64(present_time), comm === nothing ? C_NULL : comm.handle, mh)
        rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), forest.ctx.handle)))
    end
    try
        rc = ccall((:gcdyn_mh_run, LIBGCDYN), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), mh[], iterations, chain, ll)
        rc == 0 || error(unsafe_string(ccall((:gcdyn_last_error, LIBGCDYN), Cstring, (Ptr{Cvoid},), forest.ctx.handle)))
        ccall((:gcdyn_mh_state, LIBGCDYN), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
              mh[], C_NULL, C_NULL, acc, C_NULL)
    finally
        ccall((:gcdyn_mh_destroy, LIBGCDYN), Cvoid, (Ptr{Cvoid},), mh[])
    end
    return chain, ll, acc
end
