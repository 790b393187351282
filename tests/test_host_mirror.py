"""CPU tests of the host-side mirror of the Julia API (types.jl, treenode.jl, rates, simulator) and of
the flattener, whose integer arrays must be bit-exact with an independent restatement (oracle.flatten)."""
import math

import numpy as np
import pytest

import gcdyn
import oracle
from gcdyn import (ArgumentError, BoundsError, ConstantBranchingProcess, DimensionMismatch,
                   DiscreteBranchingProcess, SigmoidalBranchingProcess, TreeNode)


def test_events_and_treenode():
    assert gcdyn.EVENTS == oracle.EVENTS and gcdyn.EVENTS[4] == "type_change"   # types.jl:6
    with pytest.raises(ArgumentError):
        TreeNode("speciation", 0, 1)
    n = TreeNode("root", 0, 3)
    assert isinstance(n.time, float) and n.type == 3 and n.children == [] and n.up is None


def test_model_validation_rules():
    # types.jl:78-90
    with pytest.raises(ArgumentError):
        ConstantBranchingProcess(1, 1, 1, 1.5, 0, [1, 2])
    with pytest.raises(ArgumentError):
        ConstantBranchingProcess(1, 1, 1, 1, -0.1, [1, 2])
    G = np.array([[-1.0, 1.0], [0.5, -0.5]])
    with pytest.raises(DimensionMismatch):
        ConstantBranchingProcess(1, 1, 1, G, 1, 0, [1, 2, 3])
    with pytest.raises(ArgumentError):
        ConstantBranchingProcess(1, 1, 1, np.array([[-1.0, 0.5], [0.5, -0.5]]), 1, 0, [1, 2])   # row sums
    with pytest.raises(ArgumentError):
        ConstantBranchingProcess(1, 1, 1, np.array([[1.0, -1.0], [0.5, -0.5]]), 1, 0, [1, 2])   # positive diagonal
    m = ConstantBranchingProcess(2, 1, 3, 1, 0, range(1, 4))      # integers promoted, range accepted
    assert isinstance(m.λ, float) and m.δ == 1.0 and m.type_space.dtype == np.float64
    assert np.allclose(m.Γ, [[-3, 1.5, 1.5], [1.5, -3, 1.5], [1.5, 1.5, -3]])   # types.jl:149-155
    with pytest.raises(ArgumentError):          # K = 1 through the γ constructor: 0/0 rows, as the reference
        ConstantBranchingProcess(2, 1, 3, 1, 0, [1])
    with pytest.raises(AttributeError):
        m.λ = 3.0
    s = SigmoidalBranchingProcess(1, 0, 2, 0.5, 1, 1, 0.3, 0.5, [0.0, 1.0])
    assert gcdyn.λ(s, 0.0) == pytest.approx(2 * 0.5 + 0.5)
    d = DiscreteBranchingProcess([1.0, 2.0], 1, 1, 1, 0, [5, 7])
    assert gcdyn.λ(d, 7) == 2.0
    with pytest.raises(ArgumentError):
        gcdyn.λ(d, 6)
    with pytest.raises(ArgumentError):
        gcdyn.μ(m, 9)
    assert gcdyn.γ(m, 1) == 3.0 and gcdyn.γ(m, 1, 2) == 1.5 and gcdyn.γ(m, 2, 2) == 0
    assert gcdyn.type_space_index(m, 2) == 1 and gcdyn.type_space_index(m, 2.5) is None


def _chain():
    r = TreeNode("root", 0, 1)
    a = TreeNode("birth", 1, 1); gcdyn.attach(r, a)
    b = TreeNode("type_change", 1.5, 2); gcdyn.attach(a, b)
    c = TreeNode("sampled_survival", 3, 1); gcdyn.attach(a, c)
    d = TreeNode("sampled_death", 2, 2); gcdyn.attach(b, d)
    return r, a, b, c, d


def test_traversals_and_mutators():
    r, a, b, c, d = _chain()
    assert list(gcdyn.PostOrderTraversal(r)) == [d, b, c, a, r]     # treenode.jl:159-175
    assert list(gcdyn.PreOrderTraversal(r)) == [r, a, b, d, c]
    assert list(gcdyn.LeafTraversal(r)) == [d, c]
    assert len(gcdyn.PostOrderTraversal(r)) == 5
    gcdyn.delete(b)                                                # child re-attached at the END (treenode.jl:38-48)
    assert a.children == [c, d] and d.up is a and b.up is None
    gcdyn.detach(a, c)
    assert a.children == [d] and c.up is None
    # map_types prunes self-loops created by the mapping
    r, a, b, c, d = _chain()
    t2 = gcdyn.map_types(lambda x: 1, r)
    assert [n.event for n in gcdyn.PostOrderTraversal(t2)] == ["sampled_survival", "sampled_death", "birth", "root"]
    assert [n.event for n in gcdyn.PostOrderTraversal(r)][1] == "type_change"     # original untouched
    gcdyn.map_types_inplace(lambda x: 1, r)
    assert [n.event for n in gcdyn.PostOrderTraversal(r)] == ["sampled_survival", "sampled_death", "birth", "root"]


def test_rand_tree_law_and_invariants():
    # runtests.jl:4-18: E[#sampled survivors] = exp(T(λ−μ)) with reject_stubs = false, γ = 0, ρ = σ = 1
    rng = np.random.default_rng(2)
    lam, mu, T = 2.5, 1.1, 2
    model = ConstantBranchingProcess(lam, mu, 0, 1, 1, range(1, 4))
    trees = gcdyn.rand_tree(model, T, 1, 4000, reject_stubs=False, rng=rng)
    obs = [sum(1 for n in gcdyn.LeafTraversal(t) if n.event == "sampled_survival") for t in trees]
    assert np.mean(obs) == pytest.approx(math.exp(T * (lam - mu)), rel=1e-1)
    # structural invariants of accepted, pruned trees
    model = ConstantBranchingProcess(2.0, 1.0, 1.0, 0.5, 0.2, [1, 2, 3])
    for t in gcdyn.rand_tree(model, 3.0, 1, 50, rng=rng):
        assert t.event == "root" and t.time == 0.0 and len(t.children) == 1 and isinstance(t.type, int)
        for n in gcdyn.PostOrderTraversal(t.children[0]):
            if n.event == "birth":
                assert len(n.children) == 2 and n.type == n.up.type
            elif n.event == "type_change":
                assert len(n.children) == 1 and n.type != n.up.type
            elif n.event == "sampled_survival":
                assert not n.children and n.time == 3.0
            elif n.event == "sampled_death":
                assert not n.children and n.time < 3.0
            else:
                raise AssertionError(n.event)
            assert n.time >= n.up.time
    with pytest.raises(ArgumentError):     # max_rejections (mbp.jl:389-391)
        gcdyn.rand_tree(ConstantBranchingProcess(0.01, 50.0, 0, 1, 0, [1, 2]), 5.0, 1, max_rejections=3, rng=rng)
    assert all(len(gcdyn.LeafTraversal(t)) >= 5 for t in gcdyn.rand_tree(model, 3.0, 1, 5, min_leaves=5, rng=rng))


def test_flatten_bit_exact_with_oracle_restatement():
    rng = np.random.default_rng(5)
    ts = np.linspace(-2, 2, 7)
    model = SigmoidalBranchingProcess(1.5, 0.0, 2.5, 0.5, 1.0, 1.0, 0.5, 0.1, ts)
    trees = gcdyn.rand_tree(model, 3.0, ts[3], 40, rng=rng)
    flat = gcdyn.flatten(trees, ts, 3.0)
    ref = oracle.flatten(trees, ts)
    for k, v in ref.items():
        got = getattr(flat, k)
        assert got.dtype == v.dtype and np.array_equal(got, v), k
    assert np.array_equal(flat.event_counts, oracle.event_counts(trees))
    assert flat.live.all() and not flat.self_loop.any()
    # integer-typed trees against a Float64 type space (runtests.jl:7-8)
    m2 = ConstantBranchingProcess(2.5, 1.1, 1.1, 1, 1, range(1, 4))
    trees2 = gcdyn.rand_tree(m2, 2.0, 1, 10, rng=rng)
    f2 = gcdyn.flatten(trees2, m2.type_space, 2.0)
    r2 = oracle.flatten(trees2, m2.type_space)
    assert all(np.array_equal(getattr(f2, k), v) for k, v in r2.items())
    # post-order property: every child precedes its parent; CSR children reproduce `children` order
    for t in range(flat.n_trees):
        lo, hi = flat.tree_off[t], flat.tree_off[t + 1]
        par = flat.parent[lo:hi]
        assert np.all((par == -1) | (par > np.arange(hi - lo)))
        assert np.sum(par == -1) == 1 and par[-1] == -1


def test_flatten_raises_like_the_reference():
    N = TreeNode
    ts = [1.0, 2.0]
    T = 2.0

    def tree(leaf_event="sampled_survival", leaf_time=T, mid="birth"):
        r = N("root", 0, 1.0)
        b = N(mid, 1.0, 1.0); gcdyn.attach(r, b)
        gcdyn.attach(b, N(leaf_event, leaf_time, 1.0))
        gcdyn.attach(b, N("sampled_survival", T, 1.0))
        return r
    with pytest.raises(ArgumentError, match="present time"):        # mbp.jl:126-128
        gcdyn.flatten(tree(leaf_time=1.5), ts, T)
    with pytest.raises(ArgumentError, match="must not be at the present"):   # mbp.jl:132-134
        gcdyn.flatten(tree("sampled_death", T), ts, T)
    with pytest.raises(ArgumentError, match="Leaf event"):         # mbp.jl:162
        gcdyn.flatten(tree("unsampled_survival"), ts, T)
    with pytest.raises(ArgumentError, match="Leaf event"):         # stub root: the root itself is a leaf
        gcdyn.flatten(N("root", 0, 1.0), ts, T)
    with pytest.raises(ArgumentError, match="Unknown event"):      # mbp.jl:193
        gcdyn.flatten(tree(mid="unsampled_death"), ts, T)
    one_child = N("root", 0, 1.0)
    b = N("birth", 1.0, 1.0); gcdyn.attach(one_child, b); gcdyn.attach(b, N("sampled_survival", T, 1.0))
    with pytest.raises(BoundsError):                               # event.children[2], mbp.jl:178
        gcdyn.flatten(one_child, ts, T)
    bad_type = tree(); bad_type.children[0].children[0].up.type = 9.0
    with pytest.raises(ArgumentError, match="type space"):         # μ(model, up.type), mbp.jl:53-55
        gcdyn.flatten(bad_type, ts, T)
    # alternates: incompatible events (extra_likelihoods.jl:29,75)
    un = tree("unsampled_survival")
    gcdyn.flatten(un, ts, T, "naive")                              # allowed by naive
    with pytest.raises(ArgumentError, match="incompatible"):
        gcdyn.flatten(un, ts, T, "stadler")
    with pytest.raises(BoundsError):
        gcdyn.flatten(N("root", 0, 1.0), ts, T, "naive")
    # a third child of a birth is computed but never used by the reference (mbp.jl:173-179): live = 0
    r = tree()
    extra = N("type_change", 1.2, 2.0); gcdyn.attach(r.children[0], extra)
    gcdyn.attach(extra, N("sampled_survival", T, 2.0))
    f = gcdyn.flatten(r, ts, T)
    assert f.live.tolist() == [1, 1, 0, 0, 1]
    # self loop is recorded, not raised (mbp.jl:181-185)
    r = N("root", 0, 1.0)
    tc = N("type_change", 1.0, 1.0); gcdyn.attach(r, tc); gcdyn.attach(tc, N("sampled_survival", T, 1.0))
    assert gcdyn.flatten(r, ts, T).self_loop.tolist() == [1]


def test_sharding_partition_is_balanced_and_complete():
    from gcdyn.sharding import partition_trees, shard_flat
    rng = np.random.default_rng(8)
    model = ConstantBranchingProcess(2.0, 1.0, 1.0, 0.5, 0.2, [1, 2, 3])
    trees = gcdyn.rand_tree(model, 3.0, 1, 60, rng=rng)
    flat = gcdyn.flatten(trees, model.type_space, 3.0)
    sizes = np.diff(flat.tree_off)
    for R in (1, 2, 4, 8):
        parts = partition_trees(sizes, R)
        assert sorted(np.concatenate(parts).tolist()) == list(range(60))
        loads = [sizes[p].sum() for p in parts]
        assert max(loads) - min(loads) <= sizes.max()
        shards = shard_flat(flat, R)
        assert sum(s.n_nodes for s in shards) == flat.n_nodes
        for ids, s in zip(parts, shards):
            ref = gcdyn.flatten([trees[i] for i in ids], model.type_space, 3.0)
            for k in ("tree_off", "event", "type_idx", "btype_idx", "t_start", "t_end", "parent", "child_off",
                      "child_idx", "root_type_idx", "root_time"):
                assert np.array_equal(getattr(s, k), getattr(ref, k)), k


def test_maxiters_maps_onto_the_cell_cap():
    """`maxiters` of the reference's solver call (mbp.jl:116) becomes `gcdyn_opts.max_cells`; the default 1e5 and anything
    that does not fit an int32 mean "no cap beyond the table capacity"."""
    from gcdyn.likelihood import _opts
    assert _opts(maxiters=None).max_cells == 0
    assert _opts(maxiters=7).max_cells == 7
    assert _opts(maxiters=1e5).max_cells == 100000
    assert _opts(maxiters=0.5).max_cells == 1
    assert _opts(maxiters=float("inf")).max_cells == 0 and _opts(maxiters=1e12).max_cells == 0
    # tolerances looser than the library's own are not forwarded, tighter ones are
    assert _opts(reltol=1e-3, abstol=1e-3).tol == 0.0 and _opts(reltol=1e-12, abstol=1e-3).tol == 1e-12


def test_packed_params_struct_is_cached_until_an_array_is_replaced():
    import numpy as np
    import gcdyn
    from gcdyn.likelihood import pack_params
    ts = np.linspace(-1.0, 1.0, 4)
    models = [gcdyn.SigmoidalBranchingProcess(1.0, 0.0, 2.0 + 0.1 * i, 0.5, 1.0, 1.0, 0.5, 0.2, ts) for i in range(3)]
    ps = pack_params(models)
    a = ps.struct()
    assert ps.struct() is a                                   # same arrays attached: same struct, no pointer rebuilding
    ps.mu[1] = 3.0                                            # contents may change freely
    assert ps.struct() is a and a.mu[1] == 3.0
    old_mu = ps.mu
    ps.mu = ps.mu.copy()                                      # another array: the struct must point at it
    b = ps.struct()
    assert b is not a
    ps.mu[0] = 9.0
    assert b.mu[0] == 9.0 and old_mu[0] != 9.0
