"""A model of the LU-SGS pencil wavefront's scheduling (aither_b200/csrc/lusgs_pencil.cuh), run on
the CPU: pencils handed out by an atomic ticket in anti-diagonal order to R resident thread blocks,
each walking its pencil in local planes q = i + jl + kl; a plane's halo needs the boundary lines
the two pencils behind have POSTED (a plane is posted one iteration after it was finished, the last
one when the pencil ends). The model checks what the design argues:

  * no deadlock, whatever the number of resident thread blocks (1 ... more than there are pencils)
    and whatever the block's shape (clipped pencils included);
  * every cell is solved after its three neighbours behind it -- the reference's hyperplane order
    (src/utility.cpp:377-398) is one of the orders this admits, so the numbers are the same.

It is a model of the scheme, not of the CUDA code; the code itself is checked against the
reference's dumps on the GPU (tests/test_gpu_golden.py and friends).
"""
import itertools

import pytest

TJ, TK = 12, 8


def simulate(ni, nj, nk, resident, tj_full=TJ, tk_full=TK):
    nbj, nbk = -(-nj // tj_full), -(-nk // tk_full)
    order = [(s - y, y) for s in range(nbj + nbk - 1)
             for y in range(max(0, s - (nbj - 1)), min(s, nbk - 1) + 1)]
    ext = {(x, y): (min(tj_full, nj - x * tj_full), min(tk_full, nk - y * tk_full)) for x, y in order}
    finished = {p: -1 for p in order}   # last plane finished
    posted = {p: -1 for p in order}     # last plane whose boundary lines are in the mailbox
    solved_at = {}                      # cell -> round it was solved in
    ticket = 0
    running = [None] * resident         # (pencil, next plane)
    rounds = 0
    while True:
        progressed = False
        posts = []
        for r in range(resident):
            if running[r] is None:
                if ticket < len(order):
                    running[r] = (order[ticket], 0)
                    ticket += 1
                    progressed = True
                continue
            (x, y), q = running[r]
            tj, tk = ext[(x, y)]
            n_steps = ni + tj + tk - 2
            # the halo of plane q: lines (0, kS) need the pencil behind in j, lines (jS, 0) the one in k
            ok = True
            if x > 0:
                need = [q + ext[(x - 1, y)][0] - 1 for ks in range(tk) if 0 <= q - ks < ni]
                ok = ok and all(posted[(x - 1, y)] >= p for p in need)
            if y > 0:
                need = [q + ext[(x, y - 1)][1] - 1 for js in range(tj) if 0 <= q - js < ni]
                ok = ok and all(posted[(x, y - 1)] >= p for p in need)
            if not ok:
                continue
            for js, ks in itertools.product(range(tj), range(tk)):
                i = q - js - ks
                if 0 <= i < ni:
                    solved_at[(i, x * tj_full + js, y * tk_full + ks)] = rounds
            # the plane finished in the previous iteration is posted during this one
            if q > 0:
                posts.append(((x, y), q - 1))
            finished[(x, y)] = q
            progressed = True
            if q + 1 == n_steps:
                posts.append(((x, y), q))   # post(nSteps - 1) after the loop
                running[r] = None
            else:
                running[r] = ((x, y), q + 1)
        for p, q in posts:                   # visible to the others from the next round on
            posted[p] = max(posted[p], q)
        rounds += 1
        if ticket == len(order) and all(s is None for s in running):
            return solved_at, rounds
        assert progressed, "deadlock: %r" % (running,)


@pytest.mark.parametrize("shape", [(9, 12, 8), (7, 30, 20), (5, 25, 17), (16, 13, 40), (6, 48, 9)])
@pytest.mark.parametrize("resident", [1, 2, 3, 7, 64])
def test_no_deadlock_and_dependencies_respected(shape, resident):
    ni, nj, nk = shape
    solved_at, rounds = simulate(ni, nj, nk, resident)
    assert len(solved_at) == ni * nj * nk
    for (i, j, k), t in solved_at.items():
        for nb in ((i - 1, j, k), (i, j - 1, k), (i, j, k - 1)):
            if min(nb) >= 0:
                assert solved_at[nb] < t, ((i, j, k), nb)


def test_enough_resident_blocks_reach_the_critical_path():
    """with a thread block per pencil the sweep takes (planes of one pencil) + (lag per pencil step)
    rounds: the length DESIGN.md quotes, independent of the number of pencils beside the path"""
    ni, nj, nk = 40, 4 * TJ, 3 * TK
    _, rounds = simulate(ni, nj, nk, resident=64)
    n_steps = ni + TJ + TK - 2
    # a pencil starts when the pencil behind has posted plane tj - 1 (k: tk - 1): one round to draw
    # the ticket, the plane itself, one iteration until it is posted, one round until it is seen
    lag_j, lag_k = TJ + 2, TK + 2
    assert rounds <= 1 + n_steps + 3 * lag_j + 2 * lag_k + 4
