"""CPU: the algorithm of resample_plan_kernel (thesis_b200/csrc/k_resample.cu) restated with Python integers.

The reference's running sum c_i = fl(c_{i-1} + v_i) (main.py:57,61-62) is a chain of dependent float64 adds.  The
kernel scans windows of 4,096 elements as exact INTEGER prefix sums in units of the ulp of the carry's binade and ends a
window at the first element that is a tie, is not a positive normal number, or takes the sum out of the binade; that
element is added with a genuine float64 add and the next window starts behind it.  This test replays exactly those
rules (same bit manipulations) and demands np.cumsum's sequential result bit for bit -- on weights that span many
orders of magnitude, on constructed ties, zeros, denormals and non-finite values.  The kernel itself is compared with
the oracle on the GPU (tests/test_gpu_parity.py); this guards the arithmetic of the window rule without one."""
import struct

import numpy as np
import pytest

MANT = (1 << 52) - 1


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def dbl(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def windowed_cumsum(v, window=4096, seq=64):
    """(running sum, number of windows, number of windows ended early)"""
    n = len(v)
    out = np.empty(n)
    carry, pos, windows, stops = 0.0, 0, 0, 0
    while pos < n:
        cb = bits(carry)
        kexp = (cb >> 52) & 0x7FF
        fast_ok = not (cb >> 63) and 54 <= kexp < 0x7FF
        if pos == 0 or not fast_ok:                       # sequential stretch
            for i in range(pos, min(pos + seq, n)):
                carry = float(np.float64(carry) + np.float64(v[i]))
                out[i] = carry
            pos = min(pos + seq, n)
            continue
        windows += 1
        s_in = (cb & MANT) | (1 << 52)
        acc, stop = 0, None
        end = min(pos + window, n)
        for i in range(pos, end):
            x = float(v[i])
            a = 0
            if x != 0.0:
                vb = bits(x)
                ve = (vb >> 52) & 0x7FF
                if (vb >> 63) or ve == 0x7FF or ve == 0:
                    stop = i
                    break
                m = (vb & MANT) | (1 << 52)
                sft = kexp - ve
                if sft < 1:
                    stop = i
                    break
                if sft <= 54:
                    r, half = m & ((1 << sft) - 1), 1 << (sft - 1)
                    a = m >> sft
                    if r > half:
                        a += 1
                    elif r == half:
                        stop = i
                        break
            if s_in + acc + a >= (1 << 53):
                stop = i
                break
            acc += a
            out[i] = dbl((kexp << 52) | ((s_in + acc) & MANT))
        if stop is None:
            carry = dbl((kexp << 52) | ((s_in + acc) & MANT))
            pos = end
        else:
            stops += 1
            cprev = dbl((kexp << 52) | ((s_in + acc) & MANT))
            carry = float(np.float64(cprev) + np.float64(v[stop]))
            out[stop] = carry
            pos = stop + 1
    return out, windows, stops


def same(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("seed", range(4))
def test_heavy_tailed_weights(seed):
    rng = np.random.default_rng(seed)
    v = np.exp(rng.normal(15.0, 4.0, 20000))              # seven orders of magnitude, like the filter's weights
    got, windows, stops = windowed_cumsum(v)
    assert same(got, np.cumsum(v))
    assert stops >= 8 and windows <= 5 + stops + len(v) // 4096 + 1


def test_equal_weights_and_ties():
    v = np.full(10000, 1.0)                                # integers: every add exact, crossings at powers of two
    assert same(windowed_cumsum(v)[0], np.cumsum(v))
    # constructed ties: the carry sits in [2^60, 2^61), ulp 2^8; v = k * 2^8 + 2^7 is exactly half an ulp off
    v = np.concatenate(([2.0 ** 60], np.full(64, 3.0), [5 * 2.0 ** 8 + 2.0 ** 7, 2.0 ** 7, 3 * 2.0 ** 7, 2.0 ** 8 + 2.0 ** 7] * 50))
    got, _, stops = windowed_cumsum(v)
    assert same(got, np.cumsum(v)) and stops >= 100


def test_zeros_denormals_and_non_finite():
    rng = np.random.default_rng(9)
    v = rng.random(6000)
    v[:100] = 0.0                                          # leading zeros: the carry stays 0
    v[500:520] = 0.0
    v[700] = 5e-324                                        # denormal
    v[900] = 1e300                                         # jumps many binades at once
    assert same(windowed_cumsum(v)[0], np.cumsum(v))
    v[3000] = np.inf
    with np.errstate(invalid="ignore"):
        want = np.cumsum(v)
        got = windowed_cumsum(v)[0]
    assert same(got[:3000], want[:3000]) and np.all(np.isinf(got[3000:]))
    v[3000] = 1.0
    v[4000] = -0.5                                         # a negative weight (cannot happen after main.py:54-55) takes a real add
    assert same(windowed_cumsum(v)[0], np.cumsum(v))


def test_tiny_after_huge():
    v = np.concatenate(([1.0], np.full(70, 1e-3), [1e18], np.full(5000, 1.0), [3.0] * 10))   # adds that round away entirely
    assert same(windowed_cumsum(v)[0], np.cumsum(v))
