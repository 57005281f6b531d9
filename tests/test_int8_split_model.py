"""CPU model of the arithmetic behind the int8-tensor-pipe posterior kernel (bopl_b200/csrc/posterior_ozaki.cu): the error-free
digit split, the accumulator bounds, the k-permutation shared by the two operand images and the fragment order of the diagonal
blocks — restated in NumPy integer arithmetic so that the invariants the CUDA code relies on are checked without a GPU.
(The kernel itself is checked against the oracle in tests/test_gpu_int8_posterior.py.)"""
import numpy as np
import pytest


def digits_sequential(t, S):
    """ozaki_prepare (host): least-significant balanced base-256 digit first."""
    t = np.asarray(t, dtype=np.int64).copy()
    out = np.zeros(t.shape + (S,), dtype=np.int64)
    for s in range(S - 1, -1, -1):
        dg = ((t + 128) & 255) - 128
        t = (t - dg) >> 8
        out[..., s] = dg
    assert np.all(t == 0)
    return out


def digits_bias(t, S):
    """Compute warps / oz_digits_kernel (device): add 0x80 to every byte (one 64-bit add does the carries), read the bytes, flip the top bit."""
    bias = sum(0x80 << (8 * b) for b in range(S))
    u = (np.asarray(t, dtype=np.int64) + bias).astype(np.uint64)
    out = np.zeros(np.shape(t) + (S,), dtype=np.int64)
    for s in range(S):
        byte = ((u >> np.uint64(8 * (S - 1 - s))) & np.uint64(255)).astype(np.int64)
        out[..., s] = byte - 128                                # byte = digit + 128; the kernel's XOR 0x80 is this subtraction in two's complement
    return out


@pytest.mark.parametrize("S", [6, 7])
def test_digit_split_is_exact_and_both_forms_agree(S):
    rs = np.random.RandomState(S)
    lim = int(0.995 * 2 ** (8 * S - 1))                      # |x| / scale <= 0.995 (OZ_FILL)  ->  |t| <= 0.995 2^(8S-1) < 127 (256^S - 1) / 255
    t = rs.randint(-lim, lim + 1, size=20000, dtype=np.int64)
    t[:6] = [0, 1, -1, lim, -lim, 127]
    d1, d2 = digits_sequential(t, S), digits_bias(t, S)
    assert np.array_equal(d1, d2)
    assert d1.min() >= -128 and d1.max() <= 127
    w = np.array([256 ** (S - 1 - s) for s in range(S)], dtype=object)
    assert all(int(sum(int(d1[i, s]) * w[s] for s in range(S))) == int(t[i]) for i in range(2000))


def test_magic_number_rounding_matches_llrint():
    """6 digits: fma(v, 2^47 / vs, 1.5 * 2^52) leaves round-to-nearest-even(v 2^47 / vs) in the low mantissa bits."""
    rs = np.random.RandomState(1)
    vs = 4.0
    v = rs.uniform(-0.995 * vs, 0.995 * vs, size=50000)
    z = v * (2.0 ** 47 / vs) + 6755399441055744.0            # the product is exact (power-of-two scale): one rounding, as in the FMA
    t = z.view(np.int64) - np.int64(0x4338000000000000)
    assert np.array_equal(t, np.rint(v * (2.0 ** 47 / vs)).astype(np.int64))
    assert np.abs(t).max() <= 2 ** 47


def test_k_permutation_is_a_bijection_and_thread_rows_are_contiguous():
    """Row 8 nt + 2 t + e of a 64-deep block sits at position 16 t + 2 nt + e in BOTH operand images, so the sixteen values a
    thread of the fragment layout owns (t fixed) are one 16-byte row of a digit plane."""
    pos = {}
    for k in range(64):
        kp = 16 * ((k % 8) // 2) + 2 * (k // 8) + (k % 2)    # ozaki_prepare / oz_digits_kernel
        pos[k] = kp
    assert sorted(pos.values()) == list(range(64))
    for t in range(4):
        mine = sorted(pos[8 * nt + 2 * t + e] for nt in range(8) for e in range(2))
        assert mine == list(range(16 * t, 16 * t + 16))       # compute warps: img + (t >> 1) * VHALF + (t & 1) * 16 rows ...


@pytest.mark.parametrize("S", [6, 7])
def test_accumulated_digit_products_reproduce_the_dot_product(S):
    """sum_g 2^(-14-8g) acc_g with acc_g = sum_k sum_{a+s=g} dV_a dL_s equals (L/ls).(V/vs) up to the dropped pairs a+s >= S;
    every accumulator fits in s32 for n <= 16384."""
    rs = np.random.RandomState(10 + S)
    n = 16384
    L = rs.standard_normal(n) * np.exp(rs.uniform(-6, 0, size=n))
    V = rs.standard_normal(n)
    f, e = np.frexp(np.abs(L).max())
    ls = 2.0 ** (e if f <= 0.995 else e + 1)                  # ozaki_prepare: smallest power of two with max|L| <= 0.995 ls
    sigma = np.abs(V).max()
    vs = 2.0 ** np.ceil(np.log2(sigma / 0.995))
    tl = np.rint(np.ldexp(L / ls, 8 * S - 1)).astype(np.int64)
    tv = np.rint(np.ldexp(V / vs, 8 * S - 1)).astype(np.int64)
    dl, dv = digits_sequential(tl, S), digits_sequential(tv, S)
    acc = np.zeros(S, dtype=np.int64)
    for a in range(S):
        for s in range(S - a):
            acc[a + s] += int(np.dot(dv[:, a], dl[:, s]))
    assert np.abs(acc).max() < 2 ** 31
    got = sum(float(acc[g]) * 2.0 ** (-14 - 8 * g) for g in range(S))
    want = float(np.dot(L / ls, V / vs))
    # dropped pairs: (S-1 ... ) weights 2^(-14-8g), g >= S, at most S-1 pairs per g, |digit| <= 128; plus the roundings of tl, tv
    bound = n * (2.0 ** (-8 * S + 1) + 2.0 ** (-8 * S + 1)) + n * (S - 1) * 2.0 ** (14 - 14 - 8 * S) * 1.01
    assert abs(got - want) <= bound, (abs(got - want), bound)
    # exact statement: with the fixed-point operands themselves there is no error but the dropped pairs
    exact = sum(int(tl[i]) * int(tv[i]) for i in range(n))
    kept = sum(int(acc[g]) << (8 * (2 * S - 2 - g)) for g in range(S))
    dropped = exact - kept
    assert abs(dropped) <= n * sum((2 * S - 1 - g) * 128 * 128 * 256 ** (2 * S - 2 - g) for g in range(S, 2 * S - 1))


def test_integer_horner_groups_fit_the_magic_conversion():
    """Weights 0..2 and 3..5 (3..4, 5..6 for seven digits) are combined in 64-bit integers before the 1.5 * 2^52 conversion:
    the groups stay below 2^51 for the largest accumulators n = 16384 allows."""
    amax = 7 * 16384 * 128 * 128                              # |acc_g| <= (g+1) n 2^14
    assert amax < 2 ** 31
    assert (amax << 16) + (amax << 8) + amax < 2 ** 51


def test_diagonal_fragment_order_covers_the_lower_triangle_once():
    """ozaki_pack_stage / oz_stage_kernel: G = 1, 0; kt = 0..4G+3; u = 0..3 with kt <= 4G+u -> block (nt = 4G+u, kt): the 36
    lower-triangular 8x8 blocks of a 64x64 diagonal block, each once, every n-tile's k-tiles in ascending order (in-place solve)."""
    order = []
    for G in (1, 0):
        for kt in range(4 * G + 4):
            for u in range(4):
                if kt <= 4 * G + u:
                    order.append((4 * G + u, kt))
    assert len(order) == 36 and len(set(order)) == 36
    assert set(order) == {(nt, kt) for nt in range(8) for kt in range(nt + 1)}
    for nt in range(8):
        kts = [kt for (a, kt) in order if a == nt]
        assert kts == sorted(kts)
    # in place: n-tiles 4..7 are finished (and overwritten) before 0..3, which only read k-tiles 0..3
    first_low = min(i for i, (nt, _) in enumerate(order) if nt < 4)
    assert all(nt >= 4 for nt, _ in order[:first_low]) and all(nt < 4 for nt, _ in order[first_low:])
