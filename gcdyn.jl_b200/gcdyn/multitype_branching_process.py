"""Rate functions and the forward simulator (reference: src/multitype_branching_process.jl).

The likelihood itself (`loglikelihood`, mbp.jl:110-259) is NOT computed here: it
is flattened and handed to the CUDA library through the C ABI (see
`likelihood.py`).  There is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np

from .types import (ArgumentError, AbstractBranchingProcess, ConstantBranchingProcess,
                    DiscreteBranchingProcess, SigmoidalBranchingProcess, TreeNode)
from .treenode import LeafTraversal, PostOrderTraversal, attach, delete

__all__ = ["expit", "sigmoid", "type_space_index", "λ", "μ", "γ",
           "lambda_", "mu", "gamma", "rand_tree", "mutate", "sample_child"]


def expit(x):
    """``1 / (1 + exp(-x))`` (mbp.jl:4)."""
    try:
        e = math.exp(-x)
    except OverflowError:  # Julia's exp saturates to Inf
        e = math.inf
    return 1 / (1 + e)


def sigmoid(x, xscale, xshift, yscale, yshift):
    """``yscale * expit(xscale * (x - xshift)) + yshift`` (mbp.jl:7)."""
    return yscale * expit(xscale * (x - xshift)) + yshift


def type_space_index(model: AbstractBranchingProcess, type):
    """`findfirst(==(type), model.type_space)` (mbp.jl:16-18), memoised per model.

    Returns the 0-based index (Julia: 1-based) or ``None``.
    """
    cache = model._index_cache
    try:
        return cache[type]
    except KeyError:
        pass
    except TypeError:  # unhashable type value
        cache = None
    hits = np.nonzero(model.type_space == type)[0]
    idx = int(hits[0]) if len(hits) else None
    if cache is not None:
        cache[type] = idx
    return idx


def λ(model, type):
    """Birth rate of `type` (mbp.jl:27-43)."""
    if isinstance(model, ConstantBranchingProcess):
        return model.λ
    if isinstance(model, DiscreteBranchingProcess):
        i = type_space_index(model, type)
        if i is None:
            raise ArgumentError("The type must be in the type space of the model.")
        return float(model.λ[i])
    if isinstance(model, SigmoidalBranchingProcess):
        return sigmoid(type, model.λ_xscale, model.λ_xshift, model.λ_yscale, model.λ_yshift)
    raise TypeError("λ: not a branching process")


def μ(model, type):
    """Death rate (mbp.jl:52-58); type-independent but membership-checked."""
    if type_space_index(model, type) is None:
        raise ArgumentError("The type must be in the type space of the model.")
    return model.μ


_MISSING = object()


def γ(model, from_type, to_type=_MISSING):
    """`γ(model, type)` — total rate out of `type` (mbp.jl:67-75);
    `γ(model, from_type, to_type)` — one transition rate (mbp.jl:84-97; integer 0 on a self-loop)."""
    if to_type is _MISSING:
        i = type_space_index(model, from_type)
        if i is None:
            raise ArgumentError("The type must be in the type space of the model.")
        return model.δ * -model.Γ[i, i]
    if from_type == to_type:
        return 0
    i = type_space_index(model, from_type)
    j = type_space_index(model, to_type)
    if i is None or j is None:
        raise ArgumentError("The types must be in the type space of the model.")
    return model.δ * model.Γ[i, j]


lambda_, mu, gamma = λ, μ, γ  # ASCII aliases


# --------------------------------------------------------------------------- simulator

def _weighted_index(rng, weights):
    """StatsBase `sample(Weights(w))`: one uniform draw scaled by the total, then a
    cumulative walk (StatsBase 0.34.2, pinned in Manifest.toml:1502-1506)."""
    n = len(weights)
    t = rng.random() * sum(weights)
    i = 0
    cw = weights[0]
    while cw < t and i < n - 1:
        i += 1
        cw += weights[i]
    return i


def mutate(node: TreeNode, model, rng=None):
    """`mutate!(node, model)` (mbp.jl:411-420): new type ∝ Γ[i,:]/(−Γ[i,i]) with self weight 0."""
    rng = np.random.default_rng() if rng is None else rng
    i = type_space_index(model, node.type)
    if i is None:
        raise ArgumentError("The type of the node must be in the type space of the mutator.")
    with np.errstate(divide="ignore", invalid="ignore"):
        probs = model.Γ[i, :] / -model.Γ[i, i]
    probs = probs.tolist()
    probs[i] = 0
    new = model.type_space[_weighted_index(rng, probs)]
    # `node.type = …` converts the Float64 draw to the node's T (InexactError if impossible).
    if isinstance(node.type, (int, np.integer)) and not isinstance(node.type, bool):
        if new != int(new):
            raise ArgumentError(f"InexactError: cannot convert {new} to an integer type")
        new = int(new)
    else:
        new = float(new)
    node.type = new


def sample_child(parent: TreeNode, model, present_time, rng=None) -> TreeNode:
    """`sample_child!(parent, model, present_time)` (mbp.jl:432-461)."""
    rng = np.random.default_rng() if rng is None else rng
    if len(parent.children) == 2:
        raise ArgumentError("Can only have 2 children max")
    λx, μx, γx = λ(model, parent.type), μ(model, parent.type), γ(model, parent.type)
    waiting_time = rng.exponential(1 / (λx + μx + γx))
    if parent.time + waiting_time > present_time:
        k = _weighted_index(rng, [model.ρ, 1 - model.ρ])
        child = TreeNode(("sampled_survival", "unsampled_survival")[k], present_time, parent.type)
    else:
        k = _weighted_index(rng, [λx, μx, γx])
        event = ("birth", "unsampled_death", "type_change")[k]
        if event == "unsampled_death" and rng.random() < model.σ:
            event = "sampled_death"
        child = TreeNode(event, parent.time + waiting_time, parent.type)
        if event == "type_change":
            mutate(child, model, rng)
    attach(parent, child)
    return child


def rand_tree(model, present_time, init_type, n=None, *, reject_stubs=True, min_leaves=0,
              max_rejections=500, rng=None):
    """`rand_tree(model, present_time, init_type[, n]; reject_stubs, min_leaves, max_rejections)`
    (mbp.jl:318-402).  `rng` is a `numpy.random.Generator` (Julia's global RNG stream
    cannot be reproduced; only the law is the same)."""
    rng = np.random.default_rng() if rng is None else rng
    if n is not None:
        return [rand_tree(model, present_time, init_type, reject_stubs=reject_stubs,
                          min_leaves=min_leaves, max_rejections=max_rejections, rng=rng)
                for _ in range(n)]
    num_rejections = 0
    while True:
        root = TreeNode("root", 0, init_type)
        needs_children = [root]
        # Evolve a fully-observed tree (LIFO order, mbp.jl:346-356).
        while needs_children:
            node = needs_children.pop()
            child = sample_child(node, model, present_time, rng)
            if child.event == "birth":
                needs_children.append(child)
                needs_children.append(child)
            elif child.event == "type_change":
                needs_children.append(child)
        # Flag every node with a sampled descendant (mbp.jl:359-372).
        marked = set()
        for leaf in LeafTraversal(root):
            if leaf.event not in ("sampled_survival", "sampled_death"):
                continue
            node = leaf
            while node is not None and id(node) not in marked:
                marked.add(id(node))
                node = node.up
        # Prune unsampled subtrees; a birth left with one child is spliced out (mbp.jl:375-383).
        for node in PostOrderTraversal(root):
            if node.event in ("birth", "root"):
                node.children[:] = [c for c in node.children if id(c) in marked]
                if node.event == "birth" and len(node.children) == 1:
                    delete(node)
        if (reject_stubs and len(root.children) == 0) or len(LeafTraversal(root)) < min_leaves:
            num_rejections += 1
            if num_rejections > max_rejections:
                raise ArgumentError(f"{max_rejections} trees rejected.")
            continue
        return root
