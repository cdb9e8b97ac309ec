"""CPU restatement of the cold-start path -- TEST INFRASTRUCTURE ONLY (checker for
csrc_host/initial_y.cpp; never imported by the product).

The raw stream comes from numpy's own MT19937 (an independent implementation seeded through
init_by_array, like commons-math3's MersenneTwister(long) = setSeed(new int[]{high, low})); the
derived draws restate commons-math3 3.2's BitsStreamGenerator (nextDouble / nextGaussian / nextInt;
absent from /root/reference, pom.xml:81) and the rest follows, line by line,
RandomUtils.doRandomUnitVector / randomUnitVectorFarFrom (common/src/net/myrrix/common/random/
RandomUtils.java:88-140) and AlternatingLeastSquares.constructInitialY
(online/src/net/myrrix/online/factorizer/als/AlternatingLeastSquares.java:264-335) on Python dicts
in insertion order.

Pinned: the raw stream against the published mt19937ar known-answer vector (init_by_array
{0x123, 0x234, 0x345, 0x456}; the same vector commons-math3's MersenneTwisterTest holds).  Parity
of the Gaussian / unit-vector layer against the reference itself is UNPINNED: the reference holds
no test vector for it and no JVM exists here.
"""
import math

import numpy as np

MT19937AR_FIRST = [1067595299, 955945823, 477289528, 4107218783, 4228976476, 3344332714, 3355579695,
                   227628506, 810200273, 2591290167]


class MersenneTwister:
    def __init__(self, seed):
        if isinstance(seed, (list, tuple)):
            key = [int(s) & 0xFFFFFFFF for s in seed]
        else:
            seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            key = [seed >> 32, seed & 0xFFFFFFFF]
        self._rs = np.random.RandomState(np.array(key, dtype=np.uint32))
        self._buf = []
        self._gauss = float("nan")

    def _next32(self):
        if not self._buf:
            self._buf = list(self._rs.randint(0, 1 << 32, size=1024, dtype=np.uint64))[::-1]
        return int(self._buf.pop())

    def next(self, bits):
        return self._next32() >> (32 - bits)

    def nextDouble(self):
        high = self.next(26) << 26
        low = self.next(26)
        return (high | low) * 2.0 ** -52

    def nextGaussian(self):
        if math.isnan(self._gauss):
            x = self.nextDouble()
            y = self.nextDouble()
            alpha = 2 * math.pi * x
            r = math.sqrt(-2 * math.log(y))
            self._gauss = r * math.sin(alpha)
            return r * math.cos(alpha)
        g, self._gauss = self._gauss, float("nan")
        return g

    def nextInt(self, n):
        if (n & -n) == n:
            return (n * self.next(31)) >> 31
        while True:
            bits = self.next(31)
            val = bits % n
            if bits - val + (n - 1) < (1 << 31):  # Java: `bits - val + (n-1) < 0` after int overflow
                return val


def _dot(x, y):  # SimpleVectorMath.dot
    d = 0.0
    for a, b in zip(x, y):
        d += float(np.float32(a) * np.float32(b))
    return d


def _normalize(v):  # SimpleVectorMath.normalize
    total = 0.0
    for f in v:
        total += float(np.float32(f) * np.float32(f))
    norm = np.float32(math.sqrt(total))
    return (v / norm).astype(np.float32)


def random_unit_vector(dims, random):
    d = [random.nextGaussian() for _ in range(dims)]
    total = 0.0
    for g in d:
        total += g * g
    v = np.array(d, dtype=np.float32)
    return (v / np.float32(math.sqrt(total))).astype(np.float32)


def random_unit_vector_far_from(dims, far_from, random):
    size = len(far_from)
    num_samples = min(100, size)
    while True:
        v = random_unit_vector(dims, random)
        smallest = float("inf")
        for s in range(num_samples):
            other = far_from[s if size == num_samples else random.nextInt(size)]
            dist2 = 2.0 - 2.0 * _dot(v, other)
            if math.isfinite(dist2) and dist2 < smallest:
                smallest = dist2
        if math.isfinite(smallest) and not (dims == 1 and smallest == 0.0):
            if random.nextDouble() < smallest / 4.0:
                return v
        else:
            return v


def construct_initial_y(previousY, column_keys, features, random, max_far_from=100000):
    """previousY: {id: vector} or None; column_keys: RbyColumn's keys in iteration order."""
    if not previousY:
        randomY = {}
    else:
        old = len(next(iter(previousY.values())))
        if old > features:
            randomY = {k: _normalize(np.array(v[:features], np.float32)) for k, v in previousY.items()}
        elif old < features:
            randomY = {}
            for k, v in previousY.items():
                w = np.zeros(features, np.float32)
                w[:old] = v
                for i in range(old, features):
                    w[i] = np.float32(random.nextGaussian())
                randomY[k] = _normalize(w)
        else:
            randomY = previousY
    recent = []
    for v in randomY.values():
        if len(recent) >= max_far_from:
            break
        recent.append(v)
    for key in column_keys:
        if key not in randomY:
            v = random_unit_vector_far_from(features, recent, random)
            randomY[key] = v
            if len(recent) < max_far_from:
                recent.append(v)
    return randomY
