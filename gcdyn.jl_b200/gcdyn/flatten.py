"""Tree flattener: pointer trees -> post-order struct-of-arrays (the C-ABI input format).

Order and indexing follow the reference exactly:
  * nodes of a tree are `PostOrderTraversal(tree.children[1])` (treenode.jl:159-175, mbp.jl:166),
    i.e. the root itself is NOT a node record — it is per-tree metadata;
  * a node's type index is `findfirst(==(type), type_space) - 1` (mbp.jl:17), -1 if absent;
  * event codes are positions in `EVENTS` (types.jl:6).

Validation mirrors the order in which the reference raises (mbp.jl:124-218,
extra_likelihoods.jl:10-34,45-102).  These arrays are integer/bit-exact data: the
tests compare them with an independent restatement in `oracle/`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .types import ArgumentError, BoundsError, EVENT_CODE, TreeNode
from .treenode import LeafTraversal, PostOrderTraversal

__all__ = ["FlatForest", "flatten", "map_types_flat", "METHODS"]

METHODS = ("ode", "naive", "stadler", "stadler_unconditioned")

_EV_ROOT, _EV_BIRTH, _EV_SDEATH, _EV_UDEATH, _EV_TC, _EV_SSURV, _EV_USURV = range(7)


@dataclass
class FlatForest:
    """Flattened forest.  All per-node arrays have length `n_nodes` and are laid out
    tree after tree (`tree_off` is the CSR over nodes); positions stored in `parent`
    and `child_idx` are relative to the start of the node's own tree."""
    n_types: int
    tree_off: np.ndarray          # int64  [T+1]
    event: np.ndarray             # uint8  [N]   index into EVENTS
    type_idx: np.ndarray          # int32  [N]   node's own type (new type for type_change); -1 if not in type_space
    btype_idx: np.ndarray         # int32  [N]   type of the branch above the node (= up.type)
    t_start: np.ndarray           # float64[N]   up.time
    t_end: np.ndarray             # float64[N]   time
    parent: np.ndarray            # int32  [N]   position of `up` in the tree, -1 if `up` is the root
    child_off: np.ndarray         # int64  [N+1] CSR over children
    child_idx: np.ndarray         # int32  [sum children] positions, in `children` order
    live: np.ndarray              # uint8  [N]   0 = inside a subtree the reference computes but never uses
    root_type_idx: np.ndarray     # int32  [T]
    root_time: np.ndarray         # float64[T]
    self_loop: np.ndarray         # uint8  [T]   1 = reference returns -Inf for this tree (mbp.jl:181-185)
    event_counts: np.ndarray = field(default=None)  # int64 [T,7] per-tree event histogram

    @property
    def n_trees(self):
        return len(self.tree_off) - 1

    @property
    def n_nodes(self):
        return int(self.tree_off[-1])


def _index_lookup(type_space):
    ts = np.asarray(type_space, dtype=np.float64)
    cache = {}

    def lookup(t):
        try:
            return cache[t]
        except KeyError:
            hits = np.nonzero(ts == t)[0]
            cache[t] = i = int(hits[0]) if len(hits) else -1
            return i
        except TypeError:
            hits = np.nonzero(ts == t)[0]
            return int(hits[0]) if len(hits) else -1
    return lookup


def _check_leaves_ode(tree, present_time):
    # mbp.jl:124-164 — pass A runs over the leaves of the WHOLE tree, before anything else.
    for leaf in LeafTraversal(tree):
        if leaf.event == "sampled_survival":
            if leaf.time != present_time:
                raise ArgumentError("Sampled survival events must all be at the present time.")
        elif leaf.event == "sampled_death":
            if leaf.time == present_time:
                raise ArgumentError("Sampled death events must not be at the present time.")
        else:
            raise ArgumentError("Leaf event must be either `:sampled_survival` or `:sampled_death`.")


def flatten(trees, type_space, present_time=None, method="ode", *, discrete_lambda=False):
    """Flatten `trees` (a `TreeNode` or a list of them) for `method`.

    `discrete_lambda` = the model is a `DiscreteBranchingProcess`, whose `λ(model, type)`
    raises for a type outside the type space where the other two models do not (mbp.jl:31-43).
    """
    if method not in METHODS:
        raise ValueError(f"method must be one of {METHODS}")
    if isinstance(trees, TreeNode):
        trees = [trees]
    lookup = _index_lookup(type_space)
    K = len(type_space)

    tree_off = [0]
    event, type_idx, btype_idx, t_start, t_end, parent, live = [], [], [], [], [], [], []
    child_cnt, child_idx = [], []
    root_type_idx, root_time, self_loop = [], [], []
    code = EVENT_CODE

    for tree in trees:
        if method == "ode":
            _check_leaves_ode(tree, present_time)
        if len(tree.children) == 0:
            raise BoundsError("attempt to access 0-element Vector{TreeNode} at index [1]")
        nodes = PostOrderTraversal(tree.children[0])._collect()
        pos = {id(n): i for i, n in enumerate(nodes)}
        base = len(event)
        looped = 0
        dead = set()   # ids of nodes whose subtree the ODE likelihood ignores
        for i, n in enumerate(nodes):
            up = n.up
            ev = code[n.event]
            bx = lookup(up.type)
            tx = lookup(n.type)
            if not looped:
                # ---- per-node checks, in the order the reference performs them
                if method == "ode":
                    if ev == _EV_SSURV or ev == _EV_SDEATH:
                        if ev == _EV_SDEATH and bx < 0:
                            raise ArgumentError("The type must be in the type space of the model.")
                        if n.children:
                            raise KeyError("p_end has no entry for a sampled node that is not a leaf")
                    elif ev == _EV_BIRTH:
                        if discrete_lambda and bx < 0:
                            raise ArgumentError("The type must be in the type space of the model.")
                        if len(n.children) < 2:
                            raise BoundsError("birth event with fewer than 2 children")
                    elif ev == _EV_TC:
                        if n.type == up.type:
                            looped = 1   # `@warn` + `return -Inf`: nothing after this node is examined
                        elif bx < 0 or tx < 0:
                            raise ArgumentError("The types must be in the type space of the model.")
                    else:
                        raise ArgumentError(f"Unknown event type {n.event}")
                    if not looped and bx < 0:   # μ(model, parent_type) inside dp_logq_dt! (mbp.jl:290)
                        raise ArgumentError("The type must be in the type space of the model.")
                else:
                    if bx < 0:   # μ(model, node.up.type), extra_likelihoods.jl:14,49,119
                        raise ArgumentError("The type must be in the type space of the model.")
                    ok = (_EV_BIRTH, _EV_SDEATH, _EV_TC, _EV_SSURV, _EV_USURV) if method == "naive" \
                        else (_EV_BIRTH, _EV_SDEATH, _EV_TC, _EV_SSURV)
                    if ev not in ok:
                        raise ArgumentError(f"Tree contains incompatible event {n.event}")
                    if ev == _EV_TC and n.type != up.type and tx < 0:
                        raise ArgumentError("The types must be in the type space of the model.")
            event.append(ev)
            type_idx.append(tx)
            btype_idx.append(bx)
            t_start.append(up.time)
            t_end.append(n.time)
            parent.append(pos.get(id(up), -1))
            child_cnt.append(len(n.children))
            for c in n.children:
                child_idx.append(pos[id(c)])
            if ev == _EV_BIRTH:
                dead.update(id(c) for c in n.children[2:])
            elif ev == _EV_TC:
                dead.update(id(c) for c in n.children[1:])
        if dead:
            # liveness propagates downwards: walk in reverse post-order (parents first)
            flags = [1] * len(nodes)
            for i in range(len(nodes) - 1, -1, -1):
                n = nodes[i]
                if id(n) in dead or (parent[base + i] >= 0 and flags[parent[base + i]] == 0):
                    flags[i] = 0
            live.extend(flags)
        else:
            live.extend([1] * len(nodes))
        rt = lookup(tree.type)
        if rt < 0 and method in ("ode", "stadler") and not looped:
            raise ArgumentError("The type of the root must be in the type space of the model.")
        root_type_idx.append(rt)
        root_time.append(tree.time)
        self_loop.append(looped)
        tree_off.append(len(event))

    t_start = np.asarray(t_start, dtype=np.float64)
    t_end = np.asarray(t_end, dtype=np.float64)
    root_time_a = np.asarray(root_time, dtype=np.float64)
    if method in ("ode", "stadler", "stadler_unconditioned") and len(t_end):
        T = float(present_time)
        lo = min(t_start.min(), t_end.min(), root_time_a.min())
        hi = max(t_start.max(), t_end.max(), root_time_a.max())
        if not (lo >= 0.0 and hi <= T):   # also catches NaN
            raise ArgumentError("All node times must lie in [0, present_time].")
    ev_a = np.asarray(event, dtype=np.uint8)
    tree_off_a = np.asarray(tree_off, dtype=np.int64)
    child_off = np.zeros(len(event) + 1, dtype=np.int64)
    np.cumsum(np.asarray(child_cnt, dtype=np.int64), out=child_off[1:])
    counts = np.zeros((len(tree_off) - 1, 7), dtype=np.int64)
    if len(ev_a):
        tree_id = np.repeat(np.arange(len(tree_off) - 1), np.diff(tree_off_a))
        np.add.at(counts, (tree_id, ev_a), 1)
    return FlatForest(
        n_types=K, tree_off=tree_off_a, event=ev_a,
        type_idx=np.asarray(type_idx, dtype=np.int32), btype_idx=np.asarray(btype_idx, dtype=np.int32),
        t_start=t_start, t_end=t_end, parent=np.asarray(parent, dtype=np.int32),
        child_off=child_off, child_idx=np.asarray(child_idx, dtype=np.int32),
        live=np.asarray(live, dtype=np.uint8), root_type_idx=np.asarray(root_type_idx, dtype=np.int32),
        root_time=root_time_a, self_loop=np.asarray(self_loop, dtype=np.uint8), event_counts=counts)


def map_types_flat(flat: FlatForest, index_map, n_types_new, *, prune_self_loops=True) -> FlatForest:
    """`map_types(mapping, tree; prune_self_loops)` (treenode.jl:60-85) on flattened arrays.

    `index_map[k]` is the position, in the NEW type space, of the image of old type `k` under the mapping
    (e.g. the bin of a discretised affinity); a type outside the old space (index -1) stays -1.  The result equals
    `flatten(map_types(mapping, tree), new_type_space)` array for array — including the reference's child order:
    `delete!` (treenode.jl:38-48) re-attaches the children of a pruned self-loop type change at the END of its
    parent's children, and the nodes are examined in the pre-order of the ORIGINAL tree (`PreOrderTraversal`
    snapshots the node list before iterating, treenode.jl:177-194).  No pointer trees are rebuilt.  The per-node
    validation of `flatten` is not repeated (the forest was valid for the old type space; a mapping cannot
    invalidate events or times)."""
    imap = np.asarray(index_map, dtype=np.int64)
    if imap.shape != (flat.n_types,):
        raise ArgumentError("index_map must have one entry per old type")

    def remap(a):
        a = np.asarray(a, dtype=np.int64)
        return np.where(a >= 0, imap[np.clip(a, 0, None)], -1).astype(np.int32)

    tnew_all, bnew_all, rnew = remap(flat.type_idx), remap(flat.btype_idx), remap(flat.root_type_idx)
    o_tree_off = [0]
    o_event, o_type, o_btype, o_ts, o_te, o_parent, o_live, o_ccnt, o_cidx, o_loop = [], [], [], [], [], [], [], [], [], []
    for t in range(flat.n_trees):
        a, b = int(flat.tree_off[t]), int(flat.tree_off[t + 1])
        n = b - a
        ev, tn, bn = flat.event[a:b], tnew_all[a:b], bnew_all[a:b]
        te = flat.t_end[a:b]
        coff = flat.child_off[a:b + 1]
        children = [list(flat.child_idx[coff[i]:coff[i + 1]]) for i in range(n)]
        up = [int(x) for x in flat.parent[a:b]]
        root_children = [n - 1]                       # the top node closes the post-order
        if prune_self_loops:
            order, stack = [], [n - 1]
            while stack:                               # pre-order of the original tree
                v = stack.pop()
                order.append(v)
                stack.extend(reversed(children[v]))
            for v in order:
                if ev[v] == _EV_TC and tn[v] == bn[v]:
                    w = up[v]
                    lst = root_children if w < 0 else children[w]
                    lst.remove(v)
                    for c in children[v]:
                        up[c] = w
                        lst.append(c)
                    children[v] = []
                    up[v] = -2                          # gone
        if not root_children:
            raise BoundsError("attempt to access 0-element Vector{TreeNode} at index [1]")
        # post-order of the new first subtree
        post, stack = [], [(root_children[0], 0)]
        while stack:
            v, k = stack.pop()
            if k < len(children[v]):
                stack.append((v, k + 1))
                stack.append((children[v][k], 0))
            else:
                post.append(v)
        newpos = {v: i for i, v in enumerate(post)}
        dead = set()
        looped = 0
        for v in post:
            w = up[v]
            o_event.append(int(ev[v]))
            o_type.append(int(tn[v]))
            o_btype.append(int(rnew[t]) if w < 0 else int(tn[w]))
            o_ts.append(float(flat.root_time[t]) if w < 0 else float(te[w]))
            o_te.append(float(te[v]))
            o_parent.append(-1 if w < 0 else newpos[w])
            o_ccnt.append(len(children[v]))
            o_cidx.extend(newpos[c] for c in children[v])
            if ev[v] == _EV_BIRTH:
                dead.update(children[v][2:])
            elif ev[v] == _EV_TC:
                dead.update(children[v][1:])
                if tn[v] == (rnew[t] if w < 0 else tn[w]):
                    looped = 1
        flags = [1] * len(post)
        if dead:
            for i in range(len(post) - 1, -1, -1):    # parents first
                v = post[i]
                w = up[v]
                if v in dead or (w >= 0 and flags[newpos[w]] == 0):
                    flags[i] = 0
        o_live.extend(flags)
        o_loop.append(looped)
        o_tree_off.append(len(o_event))
    ev_a = np.asarray(o_event, dtype=np.uint8)
    tree_off_a = np.asarray(o_tree_off, dtype=np.int64)
    child_off = np.zeros(len(o_event) + 1, dtype=np.int64)
    np.cumsum(np.asarray(o_ccnt, dtype=np.int64), out=child_off[1:])
    counts = np.zeros((flat.n_trees, 7), dtype=np.int64)
    if len(ev_a):
        tree_id = np.repeat(np.arange(flat.n_trees), np.diff(tree_off_a))
        np.add.at(counts, (tree_id, ev_a), 1)
    return FlatForest(
        n_types=int(n_types_new), tree_off=tree_off_a, event=ev_a,
        type_idx=np.asarray(o_type, dtype=np.int32), btype_idx=np.asarray(o_btype, dtype=np.int32),
        t_start=np.asarray(o_ts, dtype=np.float64), t_end=np.asarray(o_te, dtype=np.float64),
        parent=np.asarray(o_parent, dtype=np.int32), child_off=child_off,
        child_idx=np.asarray(o_cidx, dtype=np.int32), live=np.asarray(o_live, dtype=np.uint8),
        root_type_idx=rnew, root_time=np.asarray(flat.root_time, dtype=np.float64).copy(),
        self_loop=np.asarray(o_loop, dtype=np.uint8), event_counts=counts)
