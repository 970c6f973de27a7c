"""Post-processing of psi fields (host NumPy; not on the hot path).

The reference only eyeballs pictures of the wettability validation (validation.py:419-425,
wettability_validation.png); these helpers turn the same psi field into numbers so that the
validation sweep over psi_wall and the fp32-vs-fp64 statement on the interface position can be tested."""
import numpy as np


def interface_columns(psi, row):
    """x positions (sub-cell, linear interpolation) where psi crosses 0 along `row`"""
    p = np.asarray(psi)[row]
    s = np.signbit(p)
    idx = np.nonzero(s[:-1] != s[1:])[0]
    return idx + p[idx] / (p[idx] - p[idx + 1])


def droplet_contact_angle(psi, wall_row=0):
    """Contact angle (degrees) of a droplet (psi > 0) sitting on the wall at `wall_row`, from the
    spherical-cap relation tan(theta/2) = 2 h / b with h the cap height above the wall and b its base width,
    both measured on the psi = 0 contour.  x-periodic domains: the droplet must not straddle the seam."""
    psi = np.asarray(psi)
    step = 1 if wall_row == 0 else -1
    base = interface_columns(psi, wall_row)
    if len(base) < 2:
        return float("nan")
    b = base[-1] - base[0]
    xc = int(round(0.5 * (base[0] + base[-1])))
    col = psi[::step, xc] if wall_row == 0 else psi[::-1, xc]
    s = np.signbit(col)
    top = np.nonzero(s[:-1] != s[1:])[0]
    if len(top) == 0:
        return float("nan")
    k = top[0]
    h = k + col[k] / (col[k] - col[k + 1]) + 0.5  # the wall sits half a cell below the first row (half-way bounce-back)
    return float(np.degrees(2.0 * np.arctan2(2.0 * h, b)))


def interface_shift(psi_a, psi_b):
    """largest displacement (cells) of the psi = 0 crossings between two fields, row by row"""
    worst = 0.0
    for r in range(np.asarray(psi_a).shape[0]):
        a, b = interface_columns(psi_a, r), interface_columns(psi_b, r)
        if len(a) != len(b):
            return float("inf")
        if len(a):
            worst = max(worst, float(np.max(np.abs(a - b))))
    return worst
