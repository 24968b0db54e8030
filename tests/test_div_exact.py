"""The kernels' division by a divisor known in advance (csrc/step_math.cuh, div_by_const): reciprocal times
dividend, one FMA correction step, an exact residual check, IEEE division as the fallback.  The host build of
the very same function is driven here against true division, bit for bit — random dividends over the whole
exponent range, dividends that make the quotient land next to powers of two and on ties, the divisors the
golden cases use (0.3, 0.7 and their squares) and adversarial ones (mantissa all ones, 1 + ulp, 3, 10, 1e-3)."""
import math
import random
import struct

import numpy as np


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


DIVISORS = [0.3, 0.7, 0.3 * 0.3, 0.7 * 0.7, 3.0, 10.0, 1e-3, 1.0 + 2.0 ** -52, 2.0 - 2.0 ** -52, 1.5, 0.1, 7.0 / 3.0,
            math.pi, 1e5, 123456.789, 2.0 ** -20 * 1.2345]


def test_division_by_constant_is_ieee_division(csim):
    rng = random.Random(123)
    fast = total = 0
    for d in DIVISORS:
        xs = [0.0, -0.0, 1.0, -1.0, d, -d, 3.0 * d, d * (1.0 + 2.0 ** -52), d * (2.0 - 2.0 ** -52), 5e-324, 1e-310, -1e-310,
              1e-300, 1e300, 1.7e308, math.inf, -math.inf, 2.0 ** -1000, 2.0 ** 900]
        xs += [rng.uniform(-10.0, 10.0) for _ in range(20000)]
        xs += [math.ldexp(rng.uniform(1.0, 2.0), rng.randint(-1070, 1020)) * rng.choice((-1, 1)) for _ in range(20000)]
        # quotients next to a power of two and exact products k*d (exact quotients, halves: ties after rounding)
        for k in range(1, 4000):
            xs.append(k * d)
            xs.append((2.0 ** rng.randint(-30, 30)) * d * (1.0 + rng.choice((-1, 1)) * 2.0 ** -52 * rng.randint(0, 3)))
            xs.append((k + 0.5) * 2.0 ** -52 * d)
        for a in xs:
            got, want = csim.div_by_const(a, d), (a / d if not math.isinf(a) else math.copysign(math.inf, a))
            assert _bits(got) == _bits(want) or (got != got and want != want), (a, d, got, want)
            fast += csim.div_by_const_fast(a, d)
            total += 1
    # the checked fast path is the rule, the IEEE fallback the exception (ties, power-of-two neighbours, extremes)
    assert fast / total > 0.9, (fast, total)


def test_division_by_constant_many_random_pairs(csim):
    rng = np.random.default_rng(7)
    n = 200000
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-12, 12, n)
    d = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-6, 6, n) + 1e-9
    bad = [(x, y) for x, y in zip(a.tolist(), d.tolist()) if _bits(csim.div_by_const(x, y)) != _bits(x / y)]
    assert not bad, bad[:5]
