"""CPU test of the algorithm behind k_accu_seq (pgure-svt_b200/csrc/tile_eval.cuh): a sequential FP64 running sum — what
arma::accu does with two accumulators (pguresvt.hpp:139) — emulated exactly by integer arithmetic whose per-element
operations compose ASSOCIATIVELY, so that a parallel scan can reproduce the sequential result bit for bit.  This is a plain
numpy/Python restatement of the device logic: per binade, `fl(acc + a) = (A + q + round) * 2^(e-52)`; a run of elements is
the function `A -> A + c1` (no tie) or `A -> roundup_even(A + c1) + c2` (a tie makes the result even whatever came before)."""
import struct

import numpy as np
import pytest


def bits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def from_bits(b):
    return struct.unpack("<d", struct.pack("<q", b))[0]


def element_fn(a, ebias):
    """(c1, tie, c2) of adding `a` to an accumulator with biased exponent `ebias`, or None if the emulation does not apply."""
    b = bits(a)
    if b == 0:
        return (0, 0, 0)
    eb = (b >> 52) & 0x7FF
    sh = ebias - eb
    if b < 0 or eb == 0 or sh < 1:
        return None
    if sh >= 64:
        return (0, 0, 0)
    m = (b & ((1 << 52) - 1)) | (1 << 52)
    rem, half = m & ((1 << sh) - 1), 1 << (sh - 1)
    return ((m >> sh) + (1 if rem > half else 0), 1 if rem == half else 0, 0)


def compose(f, g):
    """f first, then g (as_compose in tile_eval.cuh)."""
    fc1, ft, fc2 = f
    gc1, gt, gc2 = g
    if not gt:
        return (fc1, 1, fc2 + gc1) if ft else (fc1 + gc1, 0, 0)
    if ft:
        return (fc1, 1, ((fc2 + gc1 + 1) & ~1) + gc2)
    return (fc1 + gc1, 1, gc2)


def apply(f, A):
    c1, t, c2 = f
    return ((A + c1 + 1) & ~1) + c2 if t else A + c1


def emulated_sum(values, chunk=64):
    """Sequential sum of `values`, chunk by chunk: inside a chunk the elements' functions are composed in a TREE (any
    bracketing gives the same function: associativity) and applied once; where the preconditions fail, real adds."""
    acc = 0.0
    i = 0
    n = len(values)
    while i < n:
        b = bits(acc)
        ebias = (b >> 52) & 0x7FF
        ok = b > 0 and 0 < ebias < 0x7FF
        A0 = (b & ((1 << 52) - 1)) | (1 << 52)
        fns = []
        j = i
        while ok and j < min(i + chunk, n):
            f = element_fn(values[j], ebias)
            if f is None:
                break
            fns.append(f)
            j += 1
        if fns:
            # tree reduction, deliberately not left to right
            level = fns
            while len(level) > 1:
                nxt = [compose(level[k], level[k + 1]) for k in range(0, len(level) - 1, 2)]
                if len(level) % 2:
                    nxt.append(level[-1])
                level = nxt
            A = apply(level[0], A0)
            if A < (1 << 53):
                acc = from_bits((ebias << 52) | (A & ((1 << 52) - 1)))
                i = j
                continue
            # a binade crossing inside the run: shrink it (the device redoes the offending thread's elements with real adds)
            if len(fns) > 1:
                chunk_try = max(1, len(fns) // 2)
                sub = fns[:chunk_try]
                level = sub
                while len(level) > 1:
                    nxt = [compose(level[k], level[k + 1]) for k in range(0, len(level) - 1, 2)]
                    if len(level) % 2:
                        nxt.append(level[-1])
                    level = nxt
                A = apply(level[0], A0)
                if A < (1 << 53):
                    acc = from_bits((ebias << 52) | (A & ((1 << 52) - 1)))
                    i += chunk_try
                    continue
        acc = acc + values[i]  # real FP64 add
        i += 1
    return acc


@pytest.mark.parametrize("case", ["uniform", "u16_over_max", "ties", "sparse", "ones"])
def test_integer_emulation_equals_the_sequential_sum(case):
    rng = np.random.RandomState(5)
    n = 20000
    if case == "uniform":
        v = rng.rand(n)
    elif case == "u16_over_max":
        v = rng.randint(0, 65536, n).astype(np.float64) / 65535.0
    elif case == "ties":
        v = rng.randint(0, 2 ** 45, n).astype(np.float64) / 2.0 ** 45  # plenty of exact half-ulp ties once the sum is large
    elif case == "sparse":
        v = np.zeros(n)
        v[rng.randint(0, n, 200)] = rng.rand(200)
    else:
        v = np.ones(n)
    want = 0.0
    for x in v:
        want += x
    assert want == float(np.cumsum(v)[-1])  # (np.cumsum adds sequentially: the tests' stand-in for Armadillo's loop)
    assert emulated_sum([float(x) for x in v]) == want


def test_composition_is_associative():
    rng = np.random.RandomState(7)
    for _ in range(2000):
        f, g, h = [(int(rng.randint(0, 1 << 20)), int(rng.randint(0, 2)), 0) for _ in range(3)]
        f = (f[0], f[1], int(rng.randint(0, 1 << 20)) if f[1] else 0)
        assert compose(compose(f, g), h) == compose(f, compose(g, h))
        A = int(rng.randint(1 << 30, 1 << 31))
        assert apply(compose(f, g), A) == apply(g, apply(f, A))
