"""``jax.lax`` stand-in (scan only). TEST INFRASTRUCTURE ONLY."""


def scan(f, init, xs, length=None):
    carry = init
    n = length if xs is None else len(xs)
    ys = []
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, ys
