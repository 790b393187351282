"""Tree container operations and traversals (reference: src/treenode.jl).

Julia's `attach!`, `detach!`, `delete!`, `map_types!` lose the `!` here:
`attach`, `detach`, `delete`, `map_types_inplace` (aliases `attach_b`... are not
provided; the mapping is documented in INTEGRATION.md).
"""
from __future__ import annotations

from .types import TreeNode

__all__ = ["attach", "detach", "delete", "map_types", "map_types_inplace",
           "PostOrderTraversal", "PreOrderTraversal", "LeafTraversal"]


def attach(parent: TreeNode, child: TreeNode) -> None:
    """`attach!(parent, child)` (treenode.jl:10-15): append to `children`, set `up`."""
    parent.children.append(child)
    child.up = parent


def detach(parent: TreeNode, child: TreeNode) -> None:
    """`detach!(parent, child)` (treenode.jl:24-29): remove by identity, clear `up`."""
    parent.children[:] = [c for c in parent.children if c is not child]
    child.up = None


def delete(node: TreeNode) -> None:
    """`delete!(node)` (treenode.jl:38-48): splice a node out.

    Its children are re-attached at the END of the grandparent's child list, so
    sibling order can change — the flattener depends on this being reproduced.
    """
    parent = node.up
    children = list(node.children)
    detach(parent, node)
    for child in children:
        detach(node, child)
        attach(parent, child)


class _TreeTraversal:
    """Eager traversal (treenode.jl:122-234): the node list is collected when
    iteration starts, so mutating the tree while iterating is safe, as in Julia."""

    __slots__ = ("root",)

    def __init__(self, root):
        self.root = root

    def _collect(self):
        raise NotImplementedError

    def __iter__(self):
        if self.root is None:
            return iter(())
        return iter(self._collect())

    def __len__(self):  # treenode.jl:225-234
        return len(self._collect()) if self.root is not None else 0


class PostOrderTraversal(_TreeTraversal):
    """Children (in `children` order) before the node itself (treenode.jl:159-175)."""

    def _collect(self):
        out = []
        # Iterative form of the reference's recursive `collect_nodes!`.
        stack = [(self.root, 0)]
        while stack:
            node, i = stack.pop()
            if i < len(node.children):
                stack.append((node, i + 1))
                stack.append((node.children[i], 0))
            else:
                out.append(node)
        return out


class PreOrderTraversal(_TreeTraversal):
    """Node before its children (treenode.jl:177-194)."""

    def _collect(self):
        out = []
        stack = [self.root]
        while stack:
            node = stack.pop()
            out.append(node)
            stack.extend(reversed(node.children))
        return out


class LeafTraversal(_TreeTraversal):
    """Nodes with no children, left to right (treenode.jl:196-215)."""

    def _collect(self):
        out = []
        stack = [self.root]
        while stack:
            node = stack.pop()
            if len(node.children) == 0:
                out.append(node)
            else:
                stack.extend(reversed(node.children))
        return out


def map_types(mapping, tree, *, prune_self_loops=True):
    """`map_types(mapping, tree; prune_self_loops=true)` (treenode.jl:60-85)."""
    new_tree = TreeNode(tree.event, tree.time, mapping(tree.type))
    stack = [(new_tree, tree)]
    while stack:
        new_node, old_node = stack.pop()
        pairs = []
        for old_child in old_node.children:
            new_child = TreeNode(old_child.event, old_child.time, mapping(old_child.type))
            attach(new_node, new_child)
            pairs.append((new_child, old_child))
        stack.extend(pairs)
    for node in PreOrderTraversal(new_tree):
        if prune_self_loops and node.event == "type_change" and node.type == node.up.type:
            delete(node)
    return new_tree


def map_types_inplace(mapping, tree, *, prune_self_loops=True):
    """`map_types!(mapping, tree; prune_self_loops=true)` (treenode.jl:98-108)."""
    for node in PreOrderTraversal(tree):
        node.type = mapping(node.type)
        if prune_self_loops and node.event == "type_change" and node.type == node.up.type:
            delete(node)
