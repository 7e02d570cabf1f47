"""``jax.tree_util`` / ``jax.tree`` stand-in for dict / tuple / namedtuple / list pytrees. TEST INFRASTRUCTURE ONLY."""


def _is_leaf(x):
    return not isinstance(x, (dict, tuple, list))


def tree_map(f, tree, *rest):
    if _is_leaf(tree):
        return f(tree, *rest)
    if isinstance(tree, dict):
        return {k: tree_map(f, tree[k], *[r[k] for r in rest]) for k in tree}
    vals = [tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)]
    if hasattr(tree, "_fields"):
        return type(tree)(*vals)
    return type(tree)(vals)


map = tree_map  # noqa: A001


def tree_leaves(tree):
    if _is_leaf(tree):
        return [tree]
    items = tree.values() if isinstance(tree, dict) else tree
    out = []
    for t in items:
        out.extend(tree_leaves(t))
    return out


leaves = tree_leaves


def tree_flatten(tree):
    lv = tree_leaves(tree)

    def rebuild(new):
        it = iter(new)
        return tree_map(lambda _: next(it), tree)
    return lv, rebuild
