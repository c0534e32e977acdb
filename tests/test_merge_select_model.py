"""Executable model of merge_candidates' radix-select path (csrc/flat_tensor.cu): among thousands of
(score, row) keys find the score V of rank k'' with four 8-bit passes over the order-preserving
score bits, move the keys with score <= V into a 256-entry buffer, sort only those. The model checks
the selection against a full sort: same first k'' keys, same tau (score of the key of rank k''),
and the fall-back conditions (fewer valid keys than k'' + 1; more ties than the buffer holds)."""
import numpy as np
import pytest

NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def f32_ord(x):                                       # scn::f32_ord
    x = np.asarray(x, np.float32) + np.float32(0)
    u = x.view(np.uint32)
    o = np.where(u & 0x80000000, ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    return np.where(np.isnan(x), np.uint32(0xFFFFFFFF), o)


def select(keys, kpp, cap=256):
    """Returns (sorted buffer or None when the kernel falls through to the full sort, V)."""
    valid = keys[keys != NONE]
    prefix, mask, rank = 0, 0, kpp
    take_all = False
    for shift in (24, 16, 8, 0):
        o = (valid >> np.uint64(32)).astype(np.uint64)
        sel = o[(o & np.uint64(mask)) == np.uint64(prefix)]
        hist = np.bincount(((sel >> np.uint64(shift)) & np.uint64(255)).astype(np.int64), minlength=256)
        total = int(hist.sum())
        if rank >= total:
            assert shift == 24                        # only the first pass can run out of keys
            take_all = True
            break
        cum = np.cumsum(hist)
        b = int(np.searchsorted(cum, rank, side="right"))
        rank -= int(cum[b - 1]) if b else 0
        prefix |= b << shift
        mask |= 255 << shift
    vmax = 0xFFFFFFFF if take_all else prefix
    buf = valid[(valid >> np.uint64(32)) <= np.uint64(vmax)]
    if len(buf) > cap:
        return None, vmax
    return np.sort(buf), vmax


@pytest.mark.parametrize("n,kpp", [(4736, 32), (2368, 32), (9472, 64), (600, 128)])
def test_selection_equals_the_head_of_a_full_sort(n, kpp):
    rng = np.random.default_rng(n + kpp)
    score = rng.standard_normal(n).astype(np.float32) * 30
    score[rng.integers(0, n, 40)] = score[0]          # a few exact ties
    rows = rng.permutation(1 << 20)[:n].astype(np.uint64)
    keys = (f32_ord(score).astype(np.uint64) << np.uint64(32)) | rows
    keys[rng.integers(0, n, n // 10)] = NONE          # unfilled list slots
    full = np.sort(keys[keys != NONE])
    buf, v = select(keys, kpp)
    assert buf is not None and len(buf) >= kpp + 1
    assert np.array_equal(buf[:kpp + 1], full[:kpp + 1])
    assert v == int(full[kpp] >> np.uint64(32))       # tau: the score of the first key left out


def test_fewer_valid_keys_than_the_cut_takes_them_all():
    keys = np.full(2048, NONE, np.uint64)
    keys[:20] = (f32_ord(np.arange(20, dtype=np.float32)).astype(np.uint64) << np.uint64(32)) | np.arange(20, dtype=np.uint64)
    buf, v = select(keys, 32)
    assert v == 0xFFFFFFFF and len(buf) == 20 and np.array_equal(buf, np.sort(keys[:20]))


def test_mass_ties_at_the_boundary_fall_through_to_the_full_sort():
    n = 4000
    score = np.full(n, 1.5, np.float32)
    score[:10] = 0.25
    keys = (f32_ord(score).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    buf, v = select(keys, 32)
    assert buf is None and v == int(f32_ord(np.float32(1.5)))
