"""Host-side arithmetic the kernels rely on, restated in Python (no GPU, no library)."""
import random


def _magic(d):
    """bsvd_capi.cu set_tile_div: ceil(2^32 / d), 0 for d == 1."""
    return 0 if d <= 1 else ((1 << 32) + d - 1) // d


def _tile_div(n, magic):
    """conv_tc.cuh tile_div: __umulhi(n, magic), or n when magic == 0."""
    return n if magic == 0 else (n * magic) >> 32


def test_reciprocal_tile_division_is_exact_below_the_checked_bound():
    """decode_tile replaces n / d by umulhi(n, ceil(2^32 / d)); plan_stage only plans a launch when
    n * d < 2^32 for each of its three dividends.  Exhaustive over the divisors that occur (tile counts per
    row / column / N up to 4096) at the multiples of d and their neighbours, plus random dividends."""
    rng = random.Random(0)
    for d in list(range(1, 4097)) + [5000, 21600, 65535, 65536, 1 << 20]:
        m = _magic(d)
        assert m < (1 << 32)
        lim = ((1 << 32) - 1) // d                      # largest n with n * d < 2^32
        cand = {0, 1, d - 1, d, d + 1, lim, lim - 1}
        for k in (1, 2, 3, 7, 1000, lim // d if d > 1 else 5):
            cand.update({k * d - 1, k * d, k * d + 1})
        cand.update(rng.randrange(0, lim + 1) for _ in range(64))
        for n in cand:
            if 0 <= n <= lim:
                assert _tile_div(n, m) == n // d, (n, d)


def test_reciprocal_tile_division_breaks_beyond_the_bound():
    """The bound is not decorative: past n * d < 2^32 the shortcut does go wrong (which is why plan_stage
    refuses such launches instead of trusting it)."""
    d = 3
    m = _magic(d)
    wrong = [n for n in range((1 << 32) - 64, 1 << 32) if _tile_div(n, m) != n // d]
    assert wrong
