"""jax.flatten_util.ravel_pytree: leaves in key-sorted order, each raveled row-major, concatenated."""
import torch

from .tree_util import tree_leaves, tree_unflatten_like


def ravel_pytree(tree):
    leaves = tree_leaves(tree)
    shapes = [tuple(l.shape) for l in leaves]
    flat = torch.cat([l.reshape(-1) for l in leaves]) if leaves else torch.zeros(0, dtype=torch.float64)

    def unravel(vec):
        out, o = [], 0
        for s in shapes:
            n = 1
            for d in s:
                n *= d
            out.append(vec[o:o + n].reshape(s))
            o += n
        return tree_unflatten_like(tree, out)
    return flat, unravel
