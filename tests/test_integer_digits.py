"""Host-side check of the fixed-point digit scheme behind structureFactorI8Kernel (csrc/kspace_tc.cu): the float -> digit
conversion done inside an FMA, the range of every digit, the exact reconstruction, the weight groups the MMAs accumulate
into and the bound on what is dropped. Pure numpy (float64 holds the FMA intermediates exactly); no GPU, no oracle."""
import numpy as np

MAGIC = np.float32(12615808.0)        # 2^23 + 0x408080
MAGIC_LO = np.float32(8388736.0)      # 2^23 + 128
RANGE = 4160000.0


def fma32(x, y, z):
    """fl32(x*y + z) for float32 inputs: the product of two float32 is exact in float64, the sum is rounded once."""
    exact = np.float64(x) * np.float64(y) + np.float64(z)       # 48-bit product + 24-bit addend of similar scale: exact here
    return exact.astype(np.float32)


def digits(x, y, nd):
    """Signed base-256 digits (D2, D1, D0[, D-1]) of x*y as the kernel forms them."""
    t = fma32(x, y, MAGIC)
    m = t.view(np.uint32) & 0xFFFFFF
    d2 = ((m >> 16) & 0xFF).astype(np.int64) - 64
    d1 = ((m >> 8) & 0xFF).astype(np.int64) - 128
    d0 = (m & 0xFF).astype(np.int64) - 128
    out = [d2, d1, d0]
    if nd == 4:
        r = np.float64(x) * np.float64(y) - (np.float64(t) - np.float64(MAGIC))        # the exact FMA residual
        assert np.all(np.abs(r) <= 0.5)
        t2 = fma32(np.minimum(r, 0.4975).astype(np.float32), np.float32(256.0), MAGIC_LO)
        out.append((t2.view(np.uint32) & 0xFF).astype(np.int64) - 128)
    return out


def test_digits_reconstruct_the_fixed_point_value():
    rng = np.random.default_rng(0)
    x = (rng.uniform(-1, 1, 200000) * RANGE).astype(np.float32)
    x[:4] = np.float32([RANGE, -RANGE, 0.0, 0.5])
    y = rng.uniform(-1, 1, x.size).astype(np.float32)
    y[:4] = np.float32([1.0, 1.0, 1.0, 1.0])
    exact = np.float64(x) * np.float64(y)
    d2, d1, d0 = digits(x, y, 3)
    assert d2.min() >= -64 and d2.max() <= 63
    for d in (d1, d0):
        assert d.min() >= -128 and d.max() <= 127
    v3 = d2 * 65536 + d1 * 256 + d0
    assert np.all(v3 == np.rint(exact))                       # three digits = round-to-nearest-even integer of the product
    d2b, d1b, d0b, dm = digits(x, y, 4)
    assert np.array_equal(d2, d2b) and np.array_equal(d1, d1b) and np.array_equal(d0, d0b)
    assert dm.min() >= -128 and dm.max() <= 127
    v4 = v3 + dm / 256.0
    assert np.abs(v4 - exact).max() <= 1.0 / 512 + 0.0025 + 1e-9      # half a unit of the fourth digit (+ the clamp at +1/2)


def test_weight_groups_hold_the_product_up_to_the_dropped_terms():
    rng = np.random.default_rng(1)
    n = 4096
    a = (rng.uniform(-1, 1, n) * RANGE).astype(np.float32)
    z = (rng.uniform(-1, 1, n) * RANGE).astype(np.float32)
    one = np.ones(n, np.float32)
    for nd in (3, 4):
        A, Z = digits(a, one, nd), digits(z, one, nd)
        groups = [0, 0, 0, 0]                                  # weights 2^32, 2^24, 2^16, 2^8
        dropped = 0.0
        for i in range(nd):
            for k in range(nd):
                prod = A[i] * Z[k]
                if i + k <= 3:
                    groups[i + k] = groups[i + k] + prod
                else:
                    dropped = dropped + prod.astype(np.float64) * 2.0 ** (32 - 8 * (i + k))
        # what one MMA accumulates per atom stays far inside int32 for 40,960 atoms
        assert max(int(np.abs(g).max()) for g in groups) * 40960 < 2 ** 31
        total = sum(g.astype(object) * (1 << (32 - 8 * w)) for w, g in enumerate(groups))     # exact Python integers
        va = sum(d.astype(np.float64) * 2.0 ** (16 - 8 * i) for i, d in enumerate(A))
        vz = sum(d.astype(np.float64) * 2.0 ** (16 - 8 * i) for i, d in enumerate(Z))
        full = va * vz                                         # product of the digit values, exact enough in float64
        err = np.abs(np.array([float(t) for t in total]) + dropped - full)
        assert err.max() <= 0.05                               # float64 rounding of ~1e13-sized numbers, nothing else
        assert np.abs(dropped).max() <= 3 * 128 * 128 + 2 * 128 * 128 / 256 + 1        # weights below 2^8: <= 3 x 2^14 (+ tails)
        # three digits carry the operand to 1/2, four digits to 1/512, of 4.16e6
        assert np.abs(va - np.float64(a)).max() <= (0.5 if nd == 3 else 1.0 / 512 + 0.0025) + 1e-9
