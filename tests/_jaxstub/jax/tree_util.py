"""Minimal pytree helpers (dicts / lists / tuples of tensors; dict keys in sorted order, as JAX flattens them)."""
import torch


def tree_leaves(tree):
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out += tree_leaves(tree[k])
        return out
    if isinstance(tree, (list, tuple)):
        out = []
        for v in tree:
            out += tree_leaves(v)
        return out
    return [tree]


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: tree_map(f, tree[k], *[r[k] for r in rest]) for k in tree}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return f(tree, *rest)


def tree_unflatten_like(tree, leaves):
    it = iter(leaves)

    def rec(t):
        if isinstance(t, dict):
            vals = {k: rec(t[k]) for k in sorted(t)}
            return {k: vals[k] for k in t}
        if isinstance(t, (list, tuple)):
            return type(t)(rec(v) for v in t)
        return next(it)
    return rec(tree)
