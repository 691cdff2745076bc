"""Streaming moments (API of reference ``hss/moments/__init__.py``).

The two scalar recurrences are kept for API parity; the FSST normalisation uses their parallel
form (per-column moments + Chan pairwise merges) inside ``if_reassign_kernel`` /
``stats_finalize_kernel`` of ``csrc/fsst.cu``.
"""


def update_mean(m: float, x: float, k: int) -> float:
    """Running mean after seeing ``x`` as the ``k``-th value (reference hss/moments/__init__.py:1-16)."""
    return m + (x - m) / k


def update_variance(x: float, m: float, var: float, k: int) -> float:
    """Welford update of the sum of squared deviations (reference hss/moments/__init__.py:19-36).

    ``m`` is the mean *before* ``x``; like the reference this returns the running M2, i.e. the
    caller divides by ``k - 1`` to get the unbiased variance.
    """
    d = x - m
    m_new = m + d / k
    return var + d * (x - m_new)


__all__ = ["update_mean", "update_variance"]
