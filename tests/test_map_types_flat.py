"""`map_types` on flattened arrays against the pointer-tree path of the host mirror
(`map_types` restating treenode.jl:60-85, then `flatten`): every array must come out identical."""
import numpy as np
import pytest

import gcdyn
from gcdyn.flatten import flatten, map_types_flat

FIELDS = ("tree_off", "event", "type_idx", "btype_idx", "t_start", "t_end", "parent", "child_off", "child_idx", "live",
          "root_type_idx", "root_time", "self_loop", "event_counts")


def _forest(seed, K=8, n=12, T=3.0):
    rng = np.random.default_rng(seed)
    ts = np.linspace(-2.0, 2.0, K)
    model = gcdyn.SigmoidalBranchingProcess(1.3, 0.0, 2.4, 0.5, 1.0, 2.5, 0.6, 0.2, ts)     # δ large: many type changes
    return model, ts, gcdyn.rand_tree(model, T, ts[K // 2], n, rng=rng), T


@pytest.mark.parametrize("seed,bins", [(1, 3), (2, 2), (3, 4), (4, 1)])
def test_flat_map_types_equals_pointer_tree_map_types(seed, bins):
    model, ts, trees, T = _forest(seed)
    edges = np.linspace(ts[0], ts[-1] + 1e-9, bins + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    binof = lambda x: int(np.searchsorted(edges, x, side="right") - 1)
    mapping = lambda x: float(centres[binof(x)])
    flat_old = flatten(trees, ts, T)
    index_map = np.array([binof(x) for x in ts])
    for prune in (True, False):
        got = map_types_flat(flat_old, index_map, bins, prune_self_loops=prune)
        want = flatten([gcdyn.map_types(mapping, t, prune_self_loops=prune) for t in trees], centres, T)
        assert got.n_types == want.n_types == bins
        for f in FIELDS:
            assert np.array_equal(getattr(got, f), getattr(want, f)), (f, prune)
        if prune:
            assert not got.self_loop.any()
            assert got.n_nodes <= flat_old.n_nodes
    assert map_types_flat(flat_old, index_map, bins).n_nodes < flat_old.n_nodes or bins == len(ts)


def test_identity_map_is_a_no_op_and_bad_map_is_rejected():
    model, ts, trees, T = _forest(7)
    flat = flatten(trees, ts, T)
    same = map_types_flat(flat, np.arange(len(ts)), len(ts))
    for f in FIELDS:
        assert np.array_equal(getattr(same, f), getattr(flat, f)), f
    with pytest.raises(gcdyn.ArgumentError):
        map_types_flat(flat, np.arange(3), 3)


@pytest.mark.gpu
def test_likelihood_of_a_flat_rebinned_forest(ctx):
    """The rebinned forest goes straight to the GPU: same value as rebuilding pointer trees first."""
    model, ts, trees, T = _forest(11)
    bins = 3
    edges = np.linspace(ts[0], ts[-1] + 1e-9, bins + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    binof = lambda x: int(np.searchsorted(edges, x, side="right") - 1)
    coarse = gcdyn.SigmoidalBranchingProcess(1.3, 0.0, 2.4, 0.5, 1.0, 2.5, 0.6, 0.2, centres)
    flat_new = map_types_flat(flatten(trees, ts, T), np.array([binof(x) for x in ts]), bins)
    f1 = gcdyn.Forest(flat_new, None, T, ctx=ctx)
    got = gcdyn.loglikelihood(coarse, f1, T)
    want = gcdyn.loglikelihood(coarse, [gcdyn.map_types(lambda x: float(centres[binof(x)]), t) for t in trees], T)
    assert got == want


def _hand_tree():
    """root(1.0) → tc(→2.0) → tc(→3.0) → birth → [tc(→1.0) → survivor, death]; types on a 3-point space."""
    N = gcdyn.TreeNode
    r = N("root", 0.0, 1.0)
    a = N("type_change", 0.4, 2.0); gcdyn.attach(r, a)
    b = N("type_change", 0.7, 3.0); gcdyn.attach(a, b)
    c = N("birth", 1.0, 3.0); gcdyn.attach(b, c)
    d = N("type_change", 1.5, 1.0); gcdyn.attach(c, d)
    gcdyn.attach(d, N("sampled_survival", 2.0, 1.0))
    gcdyn.attach(c, N("sampled_death", 1.8, 3.0))
    return r


@pytest.mark.parametrize("bins", [
    {1.0: 1.0, 2.0: 1.0, 3.0: 3.0},      # the TOP node becomes a self-loop: its child moves up to the root
    {1.0: 1.0, 2.0: 3.0, 3.0: 3.0},      # the second type change of a chain disappears
    {1.0: 2.0, 2.0: 2.0, 3.0: 2.0},      # everything collapses: every type change is pruned, child order changes
])
def test_hand_built_chains_of_type_changes(bins):
    ts_old = np.array([1.0, 2.0, 3.0])
    ts_new = np.array(sorted(set(bins.values())))
    T = 2.0
    tree = _hand_tree()
    flat = flatten([tree], ts_old, T)
    imap = np.array([int(np.nonzero(ts_new == bins[x])[0][0]) for x in ts_old])
    for prune in (True, False):
        got = map_types_flat(flat, imap, len(ts_new), prune_self_loops=prune)
        want = flatten([gcdyn.map_types(lambda x: bins[x], tree, prune_self_loops=prune)], ts_new, T)
        for f in FIELDS:
            assert np.array_equal(getattr(got, f), getattr(want, f)), (f, prune, bins)
    pruned = map_types_flat(flat, imap, len(ts_new))
    assert pruned.n_nodes == flat.n_nodes - sum(1 for a, b in ((1.0, 2.0), (2.0, 3.0), (3.0, 1.0)) if bins[a] == bins[b])
