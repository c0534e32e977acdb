"""Executable model of hnsw_search's visited table (csrc/hnsw_search.cu: visited_insert_warp): an
open-addressing table of 4-slot groups, private to one warp, filled by a warp-collective insert
WITHOUT atomics. The model replays the kernel's round protocol lane by lane and checks it against
a Python set: a key is reported fresh exactly once, no key is stored twice, and "a group that still
has an empty slot and does not hold the key proves the key absent" — the invariant every lookup
relies on — holds after any sequence of batches, including groups that overflow into their
neighbours. (The CUDA code itself is checked on the GPU by the walk-identity tests.)"""
import numpy as np
import pytest

EMPTY = 0


def home_group(row, n_groups):                       # scn::home_group
    return ((int(row) * 2654435761) & 0xFFFFFFFF) * n_groups >> 32


def insert_batch(tab, n_groups, rows):
    """One warp batch (<= 32 distinct rows, one per lane). Returns the per-lane `fresh` flags."""
    lanes = len(rows)
    assert lanes <= 32 and len(set(rows)) == lanes   # scn_graph_upload drops repeated neighbours
    key = [r + 1 for r in rows]
    g = [home_group(r, n_groups) for r in rows]
    pending, fresh = [True] * lanes, [False] * lanes
    for _round in range(n_groups + 1):
        if not any(pending):
            break
        snap = tab.copy()                            # every lane loads its group before anyone stores
        e = [4] * lanes
        for l in range(lanes):
            if not pending[l]:
                continue
            grp = snap[g[l] * 4:g[l] * 4 + 4]
            if key[l] in grp:
                pending[l] = False                   # visited
            else:
                empties = [i for i in range(4) if grp[i] == EMPTY]
                e[l] = empties[0] if empties else 4
                if empties:                          # slots fill in order: the empties are a suffix
                    assert empties == list(range(empties[0], 4))
        gi = [g[l] if (pending[l] and e[l] < 4) else None for l in range(lanes)]
        for l in range(lanes):
            if not pending[l]:
                continue
            rank = sum(1 for j in range(l) if gi[j] is not None and gi[j] == g[l])
            slot = e[l] + rank
            if slot < 4:
                assert tab[g[l] * 4 + slot] == EMPTY
                tab[g[l] * 4 + slot] = key[l]         # plain store, nobody waits for it
                pending[l] = False
                fresh[l] = True
            else:
                g[l] = (g[l] + 1) % n_groups          # this group is (or has just become) full
    assert not any(pending)
    return fresh


def contains(tab, n_groups, row):
    g, key = home_group(row, n_groups), row + 1
    for _ in range(n_groups):
        grp = tab[g * 4:g * 4 + 4]
        if key in grp:
            return True
        if EMPTY in grp:
            return False                              # the absence proof
        g = (g + 1) % n_groups
    return False


@pytest.mark.parametrize("n_groups,n_rows,fill", [(64, 5000, 0.85), (256, 1_000_000, 0.8), (16, 200, 0.875), (1024, 3000, 0.5)])
def test_warp_collective_insert_behaves_like_a_set(n_groups, n_rows, fill):
    rng = np.random.default_rng(n_groups)
    tab = np.zeros(n_groups * 4, np.int64)
    seen = set()
    while len(seen) < fill * n_groups * 4 - 32:       # the kernel keeps the table below 7/8 full
        rows = [int(x) for x in rng.choice(n_rows, size=int(rng.integers(1, 33)), replace=False)]
        fresh = insert_batch(tab, n_groups, rows)
        for r, f in zip(rows, fresh):
            assert f == (r not in seen)
            seen.add(r)
    stored = tab[tab != EMPTY]
    assert len(stored) == len(set(stored.tolist())) == len(seen)
    for r in list(seen)[:500]:
        assert contains(tab, n_groups, r)
    for r in rng.choice(n_rows, size=500):
        assert contains(tab, n_groups, int(r)) == (int(r) in seen)


def test_a_batch_that_crowds_one_group_spills_in_lane_order():
    # 12 rows with the same home group: lanes 0-3 take its four slots, the others move on round by round
    n_groups = 8
    rows, r = [], 0
    while len(rows) < 12:
        if home_group(r, n_groups) == 3:
            rows.append(r)
        r += 1
    tab = np.zeros(n_groups * 4, np.int64)
    assert insert_batch(tab, n_groups, rows) == [True] * 12
    assert tab[12:16].tolist() == [x + 1 for x in rows[:4]]
    assert tab[16:20].tolist() == [x + 1 for x in rows[4:8]] and tab[20:24].tolist() == [x + 1 for x in rows[8:12]]
    assert insert_batch(tab, n_groups, rows[::-1]) == [False] * 12
