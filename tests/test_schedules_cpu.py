"""CPU models of the two device-side schedules that have no reference counterpart, kept in step with the CUDA
code (csrc/svd.cu `jacobi_block_kernel`, csrc/panel_xch.cuh): the properties the kernels rely on are checked
on the model, for every cluster size / block size the drivers can choose.

* block Jacobi: every column pair is rotated exactly once per sweep; within an inner round the pairs are
  disjoint; the owner / tile-half a block is pushed to at the end of a stage is the CTA that works on it in
  the next stage, and no two blocks are pushed into the same tile half.
* two-level barrier: the leaves partition the CTAs and the counters reach their targets exactly when every
  CTA has arrived."""
import itertools
import math

import pytest


def rr_player(i, r, np_):
    return 0 if i == 0 else ((i - 1 + r) % (np_ - 1)) + 1


def rr_player_fast(i, r, np_):
    if i == 0:
        return 0
    x = i - 1 + r
    if x >= np_ - 1:
        x -= np_ - 1
    return x + 1


def owner_of(X, stage, C):
    nblk = 2 * C
    if stage < 0:
        return X >> 1, X & 1
    i = 0
    if X != 0:
        i = ((X - 1 - stage) % (nblk - 1)) + 1
    partner = rr_player(nblk - 1 - i, stage, nblk)
    return (i if i < C else nblk - 1 - i), (0 if X < partner else 1)


def stage_blocks(c, stage, C):
    nblk = 2 * C
    if stage < 0:
        return 2 * c, 2 * c + 1
    I, J = rr_player(c, stage, nblk), rr_player(nblk - 1 - c, stage, nblk)
    return (I, J) if I < J else (J, I)


@pytest.mark.parametrize("C", [1, 2, 4, 8])
def test_rr_player_fast_matches(C):
    for np_ in range(2, 40, 2):
        for i in range(np_):
            for r in range(np_ - 1):
                assert rr_player_fast(i, r, np_) == rr_player(i, r, np_)


@pytest.mark.parametrize("C,bs,l", [(8, 14, 210), (8, 7, 110), (8, 4, 60), (8, 16, 256), (8, 32, 512), (4, 1, 7), (2, 1, 4),
                                    (2, 3, 11), (1, 2, 4), (8, 13, 201)])
def test_block_jacobi_sweep_covers_every_pair_once(C, bs, l):
    nblk = 2 * C
    assert nblk * bs >= l
    seen = {}

    def rotate(p, q, tag, busy):
        assert p != q and p < l and q < l
        assert p not in busy and q not in busy, "pairs of an inner round must be disjoint"
        busy.update((p, q))
        key = (min(p, q), max(p, q))
        assert key not in seen, (key, seen.get(key), tag)
        seen[key] = tag

    ncols = lambda blk: max(0, min(bs, l - blk * bs))            # noqa: E731
    # (a) intra-block stage
    npb = (bs + 1) // 2 * 2
    for rnd in range(npb - 1):
        busy = set()
        for c in range(C):
            for slot in range(npb):
                half = 1 if slot >= npb // 2 else 0
                t = slot - half * (npb // 2)
                p, q = rr_player_fast(t, rnd, npb), rr_player_fast(npb - 1 - t, rnd, npb)
                if p > q:
                    p, q = q, p
                blk = 2 * c + half
                if q < ncols(blk):
                    rotate(blk * bs + p, blk * bs + q, ("a", rnd), busy)
    # (b) cross stages
    for br in range(nblk - 1):
        blocks = [stage_blocks(c, br, C) for c in range(C)]
        assert sorted(itertools.chain(*blocks)) == list(range(nblk))     # every block exactly once per stage
        for r in range(bs):
            busy = set()
            for c in range(C):
                I, J = blocks[c]
                for w in range(bs):
                    jq = (w + r) % bs
                    if w < ncols(I) and jq < ncols(J):
                        rotate(I * bs + w, J * bs + jq, ("b", br, r), busy)
    assert len(seen) == l * (l - 1) // 2


@pytest.mark.parametrize("C", [1, 2, 4, 8])
def test_block_jacobi_pushes_reach_the_next_owner(C):
    nblk = 2 * C
    stages = [-1] + list(range(nblk - 1))
    for s_idx, stage in enumerate(stages):
        nxt = stages[(s_idx + 1) % len(stages)]
        targets = set()
        for c in range(C):
            for X in stage_blocks(c, stage, C):
                cta, half = owner_of(X, nxt, C)
                assert stage_blocks(cta, nxt, C)[half] == X          # the receiver works on X, in that half
                assert (cta, half) not in targets                    # nobody else pushes into the same half
                targets.add((cta, half))
        assert len(targets) == nblk


def pbar_leaves(G):
    nl = 1
    while nl * nl < G:
        nl += 1
    return min(nl, 16)


@pytest.mark.parametrize("G", [1, 2, 3, 7, 16, 40, 79, 148, 256])
def test_two_level_barrier_counts(G):
    nl = pbar_leaves(G)
    assert 1 <= nl <= 16 and (G <= 256)
    sizes = [(G - leaf + nl - 1) // nl for leaf in range(nl)]
    assert sum(sizes) == G and all(s >= 1 for s in sizes)            # the leaves partition the CTAs
    leaf_cnt, root = [0] * nl, 0
    for epoch in (1, 2, 3):
        order = list(range(G))
        order = order[epoch:] + order[:epoch]                        # any arrival order
        for k, b in enumerate(order):
            leaf = b % nl
            leaf_cnt[leaf] += 1
            if leaf_cnt[leaf] == epoch * sizes[leaf]:
                root += 1
            assert (root >= epoch * nl) == (k == G - 1)              # released exactly by the last arrival
    assert nl <= math.isqrt(G - 1) + 1 if G > 1 else nl == 1
