"""Host-side check of the arithmetic behind the far path's packed accumulators (csrc/vlg_pass2.cuh: pack, add as one
signed 64-bit integer, unpack with the borrow undone; headroom from the counted number of contributions).  numpy int64
wraps like the device's 64-bit atomics, so this is the same arithmetic -- no GPU needed."""
import numpy as np


def _pack(a0, a1):
    # V = (a1 << 32) + a0 as a signed 64-bit sum: low word a0, high word a1 minus the borrow of a negative a0
    hi = (a1.astype(np.int64) + (a0.astype(np.int64) >> 31)) & 0xFFFFFFFF
    return ((hi << 32) | (a0.astype(np.int64) & 0xFFFFFFFF)).astype(np.int64)


def _unpack(S):
    lo = (S & 0xFFFFFFFF).astype(np.uint32).astype(np.int32).astype(np.int64)   # low lane read back as a signed 32-bit integer
    hi = ((S - lo) >> 32).astype(np.int64)                                      # ... and subtracted before the high lane is taken
    return lo, hi


def _ceil_log2(cnt):
    return 0 if cnt <= 1 else int(cnt - 1).bit_length()


def test_pack_is_the_signed_sum_of_its_lanes():
    rng = np.random.default_rng(0)
    a0 = rng.integers(-2**31, 2**31, 10000, dtype=np.int64)
    a1 = rng.integers(-2**31, 2**31, 10000, dtype=np.int64)
    with np.errstate(over="ignore"):
        want = (a1 << 32) + a0          # wraps exactly like the device's 64-bit add
    assert np.array_equal(_pack(a0, a1), want)


def test_sums_of_packed_words_unpack_to_the_sums_of_the_lanes():
    """cnt contributions, each lane |a| <= 2^(30-h) with h = ceil(log2 cnt): whatever the signs, the lane sums stay inside
    (-2^31, 2^31), the packed sum (in ANY order: integer adds commute) unpacks to them exactly."""
    rng = np.random.default_rng(1)
    for cnt in (1, 2, 3, 17, 64, 100, 512):
        h = _ceil_log2(cnt)
        assert (1 << h) >= cnt
        bound = 1 << (30 - h)
        for trial in range(50):
            a0 = rng.integers(-bound, bound + 1, cnt, dtype=np.int64)
            a1 = rng.integers(-bound, bound + 1, cnt, dtype=np.int64)
            if trial == 0:
                a0[:] = bound; a1[:] = -bound           # worst case: every contribution at the bound
            if trial == 1:
                a0[:] = -bound; a1[:] = bound
            V = _pack(a0, a1)
            with np.errstate(over="ignore"):
                S = np.int64(0)
                for v in rng.permutation(V):
                    S = np.int64(S + v)
            lo, hi = _unpack(np.array([S]))
            assert abs(int(a0.sum())) < 2**31 and abs(int(a1.sum())) < 2**31
            assert int(lo[0]) == int(a0.sum()) and int(hi[0]) == int(a1.sum())


def test_resolution_of_a_packed_contribution():
    """A contribution v (|v| < 2^eg, eg = exponent of the group's largest gradient) is stored as round(v * 2^(30 - eg - h)):
    the rounding error is at most 2^-(31 - h) of 2^eg, i.e. 2^-(30 - h) of the largest gradient (which is >= 2^(eg-1)) --
    2^-23 for the h = 7 of BASELINE config 5, below the 1e-5 parity bar by four orders of magnitude per contribution."""
    rng = np.random.default_rng(2)
    for eg in (-20, -3, 0, 5):
        for h in (0, 4, 7, 9):
            v = (rng.random(1000, dtype=np.float32) * 2 - 1) * np.float32(2.0 ** eg) * np.float32(0.999)
            a = np.rint(v.astype(np.float64) * 2.0 ** (30 - eg - h))
            assert np.all(np.abs(a) <= 2 ** (30 - h))
            back = a * 2.0 ** -(30 - eg - h)
            assert np.max(np.abs(back - v.astype(np.float64))) <= 2.0 ** (eg - 31 + h)
