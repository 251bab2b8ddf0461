"""Oracle: alpha-stable noise (numpy float64).  TEST INFRASTRUCTURE ONLY.

Restates the arithmetic behind ``gen_skewed_levy`` / ``gen_sas``
(reference ``bem/datasets/Distributions.py:33-73``).  The S(alpha/2, 1) draw in
the reference is ``scipy.stats.levy_stable.rvs`` -- third-party, NOT vendored in
``/root/reference`` and unpinned there (``bem/requirements.txt:10`` says just
``scipy``; 1.18.1 is installed in the build image).  Its published algorithm is
the Chambers-Mallows-Stuck / Nolan transform ``_rvs_Z1`` (scipy
``stats/_levy_stable/__init__.py:429-482``, branch ``otherwise`` because
alpha' = alpha/2 != 1 and beta = 1), followed by ``X = scale * Z + loc`` (S1
parameterisation, alpha' != 1).  ``cms_totally_skewed`` restates that branch on
explicit (TH, W) so it can be compared with scipy on identical variates
(``tests/test_oracle_golden.py``), and ``kanter_A`` is the algebraically reduced
form the CUDA kernel implements.
"""
import numpy as np


def levy_scale(alpha: float) -> float:
    """scale argument at Distributions.py:45,48: 2*cos(pi*alpha/4)**(2/alpha)."""
    return 2.0 * np.cos(np.pi * alpha / 4.0) ** (2.0 / alpha)


def cms_totally_skewed(alpha_half, TH, W):
    """scipy ``_rvs_Z1`` branch ``otherwise`` with beta=1 (scipy :450-463).

    TH ~ Unif(-pi/2, pi/2), W ~ Exp(1).  Returns Z ~ S1(alpha_half, 1, 0, 1).
    """
    a = float(alpha_half)
    TH = np.asarray(TH, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    aTH = a * TH
    cosTH = np.cos(TH)
    tanTH = np.tan(TH)
    val0 = 1.0 * np.tan(np.pi * a / 2)
    th0 = np.arctan(val0) / a
    val3 = W / (cosTH / np.tan(a * (th0 + TH)) + np.sin(TH))
    res3 = val3 * ((np.cos(aTH) + np.sin(aTH) * tanTH
                    - val0 * (np.sin(aTH) - np.cos(aTH) * tanTH)) / W) ** (1.0 / a)
    return res3


def skewed_levy_from_variates(alpha, TH, W, clamp_a=None):
    """A = scale * Z  (Distributions.py:45-50).  float64 in, float32 out like
    ``torch.tensor(..., dtype=torch.float32)`` at :45."""
    if alpha == 2.0:
        return np.full(np.shape(TH), 2.0, dtype=np.float32)  # :40-42
    A = levy_scale(alpha) * cms_totally_skewed(alpha / 2.0, TH, W)
    A = A.astype(np.float32)
    if clamp_a is not None:
        A = np.clip(A, 0.0, np.float32(clamp_a))  # :49-50
    return A


def kanter_A(alpha, U, W):
    """Reduced form used by the CUDA kernel: with a' = alpha/2, U in (0, pi),

        K = sin(a'U) / sin(U)^(1/a') * (sin((1-a')U) / W)^((1-a')/a'),  A = 2 K.

    Identical in law AND pointwise (U = TH + pi/2) to ``skewed_levy_from_variates``;
    E exp(-s K) = exp(-s^a').
    """
    if alpha == 2.0:
        return np.full(np.shape(U), 2.0)
    a = alpha / 2.0
    U = np.asarray(U, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    logK = (np.log(np.sin(a * U)) - np.log(np.sin(U)) / a
            + (1.0 - a) / a * (np.log(np.sin((1.0 - a) * U)) - np.log(W)))
    return 2.0 * np.exp(logK)


def gen_skewed_levy(alpha, size, isotropic=True, clamp_a=None, rng=None):
    """Distributions.py:33-51 with an explicit numpy Generator/RandomState.

    Draw order follows scipy ``_rvs_Z1`` (:472-475): all TH first, then all W.
    Isotropic: one draw per leading index, broadcast to ``size`` (:45-46).
    """
    if alpha > 2.0 or alpha <= 0.0:
        raise Exception("Wrong value of alpha ({}) for skewed levy r.v generation".format(alpha))
    size = tuple(int(s) for s in size)
    if alpha == 2.0:
        return np.full(size, 2.0, dtype=np.float32)
    rng = np.random if rng is None else rng
    n = (size[0],) if isotropic else size
    TH = rng.uniform(-np.pi / 2.0, np.pi / 2.0, size=n)
    W = rng.standard_exponential(size=n)
    A = skewed_levy_from_variates(alpha, TH, W, clamp_a)
    if isotropic:
        A = np.ascontiguousarray(np.broadcast_to(A.reshape((size[0],) + (1,) * (len(size) - 1)), size))
    return A


def gen_sas(alpha, size, a=None, isotropic=True, clamp_eps=None, rng=None, G=None):
    """Distributions.py:57-73: eps = sqrt(A) * G, clamp to +-clamp_eps.

    NOTE (reference quirk): ``clamp_a`` is NOT forwarded to the
    inner A draw (:64)."""
    size = tuple(int(s) for s in size)
    if a is None:
        a = gen_skewed_levy(alpha, size, isotropic=isotropic, rng=rng)
    if G is None:
        rng = np.random if rng is None else rng
        G = rng.standard_normal(size=size).astype(np.float32)
    ret = np.sqrt(a.astype(np.float32)) * G.astype(np.float32)
    if clamp_eps is not None:
        ret = np.clip(ret, -np.float32(clamp_eps), np.float32(clamp_eps))
    return ret.astype(np.float32)
