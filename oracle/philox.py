"""Oracle: the counter-based variates of the CUDA kernels, restated on the host.  TEST INFRASTRUCTURE ONLY.

The reference draws its noise from numpy's / torch's global generators (``bem/datasets/Distributions.py:45,65``), so there
is no reference stream to reproduce bit for bit; what CAN be pinned pointwise is the chain
``Philox words -> lattice uniforms -> reference formula``:

* ``philox4x32`` is Philox4x32-R (Salmon et al. 2011), checked against the Random123 known-answer vectors for R = 7
  (the kernels' default, ``csrc/rng.cuh`` DLPM_PHILOX_ROUNDS) and R = 10 (``tests/test_oracle_golden.py``); ``ROUNDS`` is
  what ``words`` uses -- the tests set it from the library (``dlpm_b200_philox_rounds``);
* ``counters`` is the counter layout of ``dlpm_b200/csrc/rng.cuh`` (position, global sample index, call offset, stream tag);
* ``stable_A_from_words`` / ``normal_from_words`` map the 32-bit words to the exactly representable fp32 lattice points the
  kernels use and then evaluate the REFERENCE formulas in float64 (``oracle/stable.py::kanter_A`` = scipy's CMS branch,
  Box-Muller for N(0,1)), so a GPU draw can be compared value by value with what the reference arithmetic gives on the
  same variates (``tests/test_gpu_noise.py::test_pointwise_*``).
"""
import numpy as np

from . import stable

STREAM_A, STREAM_G, STREAM_Z, STREAM_EPS_A = 0x0A, 0x06, 0x5A, 0xEA
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


ROUNDS = 7  # DLPM_PHILOX_ROUNDS of the default build


def philox4x32(key, c0, c1, c2, c3, rounds=None):
    """Philox4x32-R on arrays of counters; ``key`` = (k0, k1) 32-bit words.  Returns four uint32 arrays."""
    rounds = ROUNDS if rounds is None else rounds
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0), p1 & _MASK, (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1), p0 & _MASK
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def words(seed, stream, offset, sample, pos):
    """The four words a kernel reads at (stream tag, call offset, GLOBAL sample index, position) -- rng.cuh::philox_at."""
    sample = np.asarray(sample, dtype=np.uint64)
    offset = int(offset)
    c3 = np.uint64(stream) | (((sample >> np.uint64(32)) & np.uint64(0xFF)) << np.uint64(8)) | np.uint64(((offset >> 32) & 0xFFFF) << 16)
    key = (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return philox4x32(key, pos, sample & _MASK, np.uint64(offset & 0xFFFFFFFF), c3)


def _lattice(x):
    """fmaf((float)x, 2^-32, 2^-33) in fp32 (one rounding of the int -> float conversion, one of the fma), as float64."""
    xf = np.asarray(x, dtype=np.uint32).astype(np.float32).astype(np.float64)
    return (xf * 2.0 ** -32 + 2.0 ** -33).astype(np.float32).astype(np.float64)


def stable_A_from_words(alpha, xu, xw, clamp_a=None):
    """rng.cuh::stable_A: U = pi u with u on the centred 32-bit lattice (reflected upper half), W = -log(1 - d)."""
    xu = np.asarray(xu, dtype=np.uint32)
    if alpha == 2.0:
        return np.full(xu.shape, 2.0)
    top = (xu >> np.uint32(31)).astype(bool)
    v = _lattice(np.where(top, ~xu, xu))
    # sin(pi u) is evaluated on v = min(u, 1 - u) by the kernel, the other two sines on fp32 u = 1 - v
    u = np.where(top, (1.0 - v).astype(np.float32).astype(np.float64), v)
    d = np.minimum(_lattice(xw), np.float64(np.float32(0.99999994)))
    W = -np.log1p(-d)
    a = alpha / 2.0
    logK = (np.log(np.sin(a * np.pi * u)) - np.log(np.sin(np.pi * v)) / a
            + (1.0 - a) / a * (np.log(np.sin((1.0 - a) * np.pi * u)) - np.log(W)))
    A = 2.0 * np.exp(logK)
    if clamp_a is not None and clamp_a >= 0:
        A = np.clip(A, 0.0, clamp_a)
    return A


def normal_from_words(x, y):
    """rng.cuh::box_muller: radius from the 32-bit lattice uniform of x, angle 2 pi (y >> 9) / 2^23.  Returns (z0, z1)."""
    xf = np.asarray(x, dtype=np.uint32).astype(np.float32)
    u = ((xf + np.float32(0.5)) * np.float32(2.0 ** -32)).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(u))
    ang = 2.0 * np.pi * (np.asarray(y, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) / 2.0 ** 23
    return r * np.cos(ang), r * np.sin(ang)


def sample_A(alpha, seed, offset, samples, stream=STREAM_A, clamp_a=None):
    """One draw per sample (isotropic A): position 0 of the sample's stream, words (x, y)."""
    r = words(seed, stream, offset, samples, 0)
    return stable_A_from_words(alpha, r[0], r[1], clamp_a)


def element_A(alpha, seed, offset, samples, inner, stream=STREAM_A, clamp_a=None):
    """Per-element draws, shape (len(samples), inner), inner % 4 == 0: quad q of a sample uses positions 2q+1 and 2q+2."""
    samples = np.asarray(samples, dtype=np.uint64)[:, None]
    q = np.arange(inner // 4, dtype=np.uint64)[None, :]
    r0, r1 = words(seed, stream, offset, samples, 2 * q + 1), words(seed, stream, offset, samples, 2 * q + 2)
    quads = [stable_A_from_words(alpha, r0[0], r0[1], clamp_a), stable_A_from_words(alpha, r0[2], r0[3], clamp_a),
             stable_A_from_words(alpha, r1[0], r1[1], clamp_a), stable_A_from_words(alpha, r1[2], r1[3], clamp_a)]
    return np.stack(quads, axis=-1).reshape(samples.shape[0], inner)


def normal(seed, offset, samples, inner, stream=STREAM_Z):
    """N(0,1) field, shape (len(samples), inner), inner % 4 == 0: quad q of a sample = position q."""
    samples = np.asarray(samples, dtype=np.uint64)[:, None]
    q = np.arange(inner // 4, dtype=np.uint64)[None, :]
    r = words(seed, stream, offset, samples, q)
    a0, a1 = normal_from_words(r[0], r[1])
    b0, b1 = normal_from_words(r[2], r[3])
    return np.stack([a0, a1, b0, b1], axis=-1).reshape(samples.shape[0], inner)


def normal6_from_words(w0, w1, w2, w3):
    """rng.cuh::normal6 -- the "sextet" scheme of the K1 fills: pair j takes its radius from the top 27 bits of word j
    (u = (k + 1/2) 2^-27) and its 15-bit angle from bits [10 j, 10 j + 10) of word 3 (high part) and the low 5 bits of word j.
    Returns six arrays (z0 .. z5)."""
    w = [np.asarray(x, dtype=np.uint32) for x in (w0, w1, w2)]
    w3 = np.asarray(w3, dtype=np.uint32)
    out = []
    for j in range(3):
        k = (w[j] >> np.uint32(5)).astype(np.float64)
        u = (k + 0.5) * 2.0 ** -27  # exact in fp32: k < 2^27 needs 27 bits -> the kernel's fmaf rounds; restate that rounding
        u = (k.astype(np.float32).astype(np.float64) * 2.0 ** -27 + 2.0 ** -28).astype(np.float32).astype(np.float64)
        r = np.sqrt(-2.0 * np.log(u))
        a = (((w3 >> np.uint32(10 * j)) & np.uint32(0x3FF)).astype(np.uint64) << np.uint64(5)) | (w[j] & np.uint32(31)).astype(np.uint64)
        ang = 2.0 * np.pi * a.astype(np.float64) / 2.0 ** 15
        out += [r * np.cos(ang), r * np.sin(ang)]
    return out


def normal_sextet(seed, offset, samples, inner, stream=STREAM_Z):
    """N(0,1) field of the K1 fills for rows of inner % 384 == 0 elements (rng.cuh, "sextet" scheme): a row is cut into granules of
    96 quads; generator g = granule * 32 + lane draws the Philox blocks at positions 2 g and 2 g + 1 = twelve normals, and quad
    granule * 96 + 32 j + lane holds normals 4 j .. 4 j + 3."""
    assert inner % 384 == 0
    samples = np.asarray(samples, dtype=np.uint64)[:, None]
    n_gen = inner // 12
    g = np.arange(n_gen, dtype=np.uint64)[None, :]
    n12 = normal6_from_words(*words(seed, stream, offset, samples, 2 * g)) + normal6_from_words(*words(seed, stream, offset, samples, 2 * g + 1))
    n12 = np.stack(n12, axis=-1)  # (samples, generators, 12)
    out = np.empty((samples.shape[0], inner // 4, 4))
    gran, lane = np.arange(n_gen) // 32, np.arange(n_gen) % 32
    for j in range(3):
        out[:, gran * 96 + 32 * j + lane, :] = n12[:, :, 4 * j:4 * j + 4]
    return out.reshape(samples.shape[0], inner)


def normal_fill(seed, offset, samples, inner, stream=STREAM_Z):
    """What dlpm_b200_normal / dlpm_b200_sas (isotropic) write: the sextet scheme for rows of inner % 384 == 0 elements, else one
    Philox block per quad (``normal``).  The in-kernel noise of the step kernels always uses ``normal``."""
    return normal_sextet(seed, offset, samples, inner, stream) if inner % 384 == 0 else normal(seed, offset, samples, inner, stream)


def sas_isotropic_fill(alpha, seed, offset, samples, inner, clamp_eps=None):
    """dlpm_b200_sas, isotropic (x_T of the samplers, gen_sas): sqrt(A_b) * G with G = ``normal_fill`` of the G stream."""
    A = sample_A(alpha, seed, offset, samples, stream=STREAM_EPS_A)
    e = np.sqrt(A)[:, None] * normal_fill(seed, offset, samples, inner, stream=STREAM_G)
    if clamp_eps is not None and clamp_eps >= 0:
        e = np.clip(e, -clamp_eps, clamp_eps)
    return e


def sas_isotropic(alpha, seed, offset, samples, inner, clamp_eps=None):
    """gen_sas, isotropic: sqrt(A_b) * G with A_b from the EPS_A stream and G from the G stream of the same call offset."""
    A = sample_A(alpha, seed, offset, samples, stream=STREAM_EPS_A)
    e = np.sqrt(A)[:, None] * normal(seed, offset, samples, inner, stream=STREAM_G)
    if clamp_eps is not None and clamp_eps >= 0:
        e = np.clip(e, -clamp_eps, clamp_eps)
    return e
