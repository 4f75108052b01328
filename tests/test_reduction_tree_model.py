"""Model of the matrix reduction over ranks (``mellon_b200/csrc/mb_reduce.cu``: ``canonical_nodes``, ``push_merge``, the
local and the global phase of ``mb_gemm_tn_cells``), restated in Python and run for EVERY world size 1..32.

The GPU box offers 1, 2, 4 and 8 ranks; the algorithm is specified for any number.  The claim it makes: whatever the
number of ranks, the additions performed are exactly those of the fixed pairwise tree over the 32 chunk leaves
(((l0 + l1) + (l2 + l3)) + ...), with empty leaves (chunks beyond the last row) skipped as exact zeros.  The model
records every addition symbolically, so "same bits" becomes "same expression".
"""

import numpy as np
import pytest

NCHUNK = 32   # MB_NCHUNK (mb_common.cuh)


def canonical_nodes(lo, hi):
    """Maximal aligned subtrees covering the leaves [lo, hi), left to right (mb_reduce.cu: canonical_nodes)."""
    out = []
    while lo < hi:
        size = 1
        while lo % (2 * size) == 0 and lo + 2 * size <= hi:
            size *= 2
        out.append((lo, size))
        lo += size
    return out


def push_merge(stack, node):
    """Push; while the two topmost nodes are siblings replace them by their parent (mb_reduce.cu: push_merge).
    A node is (start, size, value); value None = every leaf beyond the last row."""
    stack.append(node)
    live = 0
    while len(stack) >= 2:
        (a0, asz, av), (b0, bsz, bv) = stack[-2], stack[-1]
        if not (asz == bsz and a0 % (2 * asz) == 0 and b0 == a0 + asz):
            break
        if av is not None and bv is not None:
            val = ("+", av, bv)
        else:
            val = av if av is not None else bv
        stack[-2:] = [(a0, 2 * asz, val)]
    live = sum(1 for n in stack if n[2] is not None)
    return live


def fixed_tree(leaves):
    """The definition: pairwise tree over the NCHUNK leaves, None = exact zero that is never added."""
    level = list(leaves)
    while len(level) > 1:
        nxt = []
        for i in range(0, len(level), 2):
            a, b = level[i], level[i + 1]
            nxt.append(("+", a, b) if a is not None and b is not None else (a if a is not None else b))
        level = nxt
    return level[0]


def rank_leaves(rank, world):
    return rank * NCHUNK // world, (rank + 1) * NCHUNK // world


def reduce_on_ranks(world, n_rows):
    """Every rank's result of the two phases of mb_gemm_tn_cells, and the largest number of buffers a rank held."""
    cr = max(1, -(-n_rows // NCHUNK))
    empty = lambda c: c * cr >= n_rows
    own, peak = [], 0
    for rank in range(world):                       # local phase: each own node is the tree over its leaves
        lo, hi = rank_leaves(rank, world)
        nodes = []
        for start, size in canonical_nodes(lo, hi):
            st = []
            for c in range(start, start + size):
                live = push_merge(st, (c, 1, None if empty(c) else f"l{c}"))
                peak = max(peak, live + sum(1 for n in nodes if n[2] is not None))
            assert len(st) == 1
            nodes.append(st[0])
        own.append(nodes)
    results = []
    for me in range(world):                         # global phase: walk every rank's nodes in order; the owner broadcasts
        st = []
        held = sum(1 for n in own[me] if n[2] is not None)
        for rho in range(world):
            lo, hi = rank_leaves(rho, world)
            theirs = canonical_nodes(lo, hi)
            assert [(s, z) for s, z, _ in own[rho]] == theirs   # every rank derives the same node list
            for k, (start, size) in enumerate(theirs):
                val = own[rho][k][2]
                if rho != me and val is not None:
                    held += 1                        # an incoming buffer from the pool
                before = sum(1 for n in st if n[2] is not None)
                push_merge(st, (start, size, val))
                after = sum(1 for n in st if n[2] is not None)
                held -= max(0, before + (1 if val is not None else 0) - after)   # merged buffers go back to the pool
                peak = max(peak, held)
        assert len(st) == 1 and st[0][:2] == (0, NCHUNK)
        results.append(st[0][2])
    return results, peak


@pytest.mark.parametrize("world", list(range(1, NCHUNK + 1)))
def test_every_world_size_performs_the_additions_of_the_fixed_tree(world):
    for n_rows in (1_000_000, 100, 33, 32, 31, 17, 5, 1):
        cr = max(1, -(-n_rows // NCHUNK))
        want = fixed_tree([None if c * cr >= n_rows else f"l{c}" for c in range(NCHUNK)])
        results, peak = reduce_on_ranks(world, n_rows)
        for got in results:
            assert got == want
        # the scratch of mb_gemm_tn_cells: own nodes + 8 buffers
        lo, hi = rank_leaves(0, world)
        most_nodes = max(len(canonical_nodes(*rank_leaves(r, world))) for r in range(world))
        assert peak <= most_nodes + 8


def test_canonical_nodes_are_aligned_power_of_two_blocks():
    for lo in range(NCHUNK):
        for hi in range(lo, NCHUNK + 1):
            nodes = canonical_nodes(lo, hi)
            assert sum(z for _, z in nodes) == hi - lo
            pos = lo
            for start, size in nodes:
                assert start == pos and size & (size - 1) == 0 and start % size == 0
                pos += size
            assert len(nodes) <= 2 * 5 - 1 + 1     # at most ~2 log2(32) blocks


def test_rank_blocks_tile_the_leaves_for_every_world_size():
    """mb_row_block: rank r owns leaves [r*32/W, (r+1)*32/W) - contiguous, disjoint, covering (ranks beyond 32 own none)."""
    for world in range(1, 40):
        pos = 0
        for r in range(world):
            lo, hi = rank_leaves(r, world)
            assert lo == pos and hi >= lo
            pos = hi
        assert pos == NCHUNK


def test_the_model_sums_like_numpy_in_tree_order():
    """Evaluating the symbolic expression on numbers gives the bits of the explicit pairwise tree."""
    rng = np.random.default_rng(0)
    vals = {f"l{c}": rng.standard_normal() * 10.0 ** rng.integers(-8, 8) for c in range(NCHUNK)}

    def ev(e):
        return vals[e] if isinstance(e, str) else ev(e[1]) + ev(e[2])

    level = [vals[f"l{c}"] for c in range(NCHUNK)]
    while len(level) > 1:
        level = [level[i] + level[i + 1] for i in range(0, len(level), 2)]
    for world in (1, 3, 5, 7, 8, 12, 32):
        results, _ = reduce_on_ranks(world, 10_000)
        assert all(ev(r) == level[0] for r in results)
