"""Compartment signal models as zonal functions: per shell a profile of t = g . d, band-limited to even Legendre orders <= 12.

The closed forms are the standard ones the reference synthesises with ``amico/synthesis.py`` (stick / zeppelin / ball, Watson
dispersed NODDI sticks with tortuosity, Van Gelderen cylinder, Murday-Cotts sphere, astrosticks); on the reference's own
500-direction high-resolution scheme they agree with that module to 1.5e-8 (NODDI), 5e-8 (zeppelin), <= 4e-6 (cylinder at
b = 4000, the lmax = 12 truncation) -- ``tests/test_signals_cpu.py`` checks it against committed outputs of the reference.
Each profile callable carries its Legendre coefficients as ``.coeffs`` (what ``generate`` rotates, ``amico/lut.pyx:227-271``).
"""
from __future__ import annotations

import math

import numpy as np
from numpy.polynomial import legendre as npleg
from scipy import optimize, special

from .scheme import GAMMA

LMAX = 12

_GL_X, _GL_W = npleg.leggauss(256)
_PL = np.stack([special.eval_legendre(l, _GL_X) for l in range(0, LMAX + 1, 2)])  # (7, 256)


def _legendre_coeffs(f_vals):
    """Even Legendre coefficients a_l (l=0..12) of a zonal profile sampled on the GL nodes."""
    l = np.arange(0, LMAX + 1, 2)
    return (2 * l + 1) / 2.0 * (_PL * (f_vals * _GL_W)).sum(axis=1)


def _band_limit(profile):
    """callable t -> S(t)  ==>  callable t -> sum_{l even<=12} a_l P_l(t)."""
    a = _legendre_coeffs(profile(_GL_X))

    def f(t):
        out = np.zeros_like(t)
        for k, l in enumerate(range(0, LMAX + 1, 2)):
            out += a[k] * special.eval_legendre(l, t)
        return out

    f.coeffs = a
    return f


def _cyl_roots(n=60):
    return special.jnp_zeros(1, n)


def _sph_roots(n=60):
    f = lambda x: x * special.jvp(1.5, x) - 0.5 * special.jv(1.5, x)
    roots, x = [], 1.0
    while len(roots) < n:
        if f(x) * f(x + 0.25) < 0:
            roots.append(optimize.brentq(f, x, x + 0.25))
        x += 0.25
    return np.array(roots)


_CYL_AM = _cyl_roots()
_SPH_AM = _sph_roots()


def _gpd_sum(am, Delta, delta, D, R, n):
    a = am / R
    dam = D * a * a
    num = 2 * dam * delta - 2 + 2 * np.exp(-dam * delta) + 2 * np.exp(-dam * Delta) \
        - np.exp(-dam * (Delta - delta)) - np.exp(-dam * (Delta + delta))
    den = dam * dam * a * a * (R * R * a * a - n)
    return float((num / den).sum())


def _watson_tau1(kappa):
    if kappa < 1e-5:
        return 1.0 / 3.0
    sk = math.sqrt(kappa)
    return -1.0 / (2.0 * kappa) + 1.0 / (2.0 * special.dawsn(sk) * sk)


def _watson_coeffs(kappa):
    f = np.exp(kappa * (_GL_X ** 2 - 1.0))
    f /= 2.0 * math.pi * (f * _GL_W).sum()
    return _legendre_coeffs(f)


class _Atom:
    """One dictionary atom: per-shell zonal profile, or isotropic per-shell value."""

    def __init__(self, per_shell, isotropic):
        self.per_shell = per_shell
        self.isotropic = isotropic


def _shell_par(sh):
    return sh["b"], sh["G"], sh["Delta"], sh["delta"]


def atom_stick(scheme, d):
    return _Atom([_band_limit(lambda t, b=sh["b"]: np.exp(-b * d * t * t)) for sh in scheme.shells], False)


def atom_zeppelin(scheme, d_par, d_perp):
    return _Atom([_band_limit(lambda t, b=sh["b"]: np.exp(-b * (d_perp + (d_par - d_perp) * t * t)))
                  for sh in scheme.shells], False)


def atom_ball(scheme, d):
    return _Atom([math.exp(-sh["b"] * d) for sh in scheme.shells], True)


def atom_noddi(scheme, d_par, kappa, v_ic):
    """v_ic * Watson-dispersed sticks + (1-v_ic) * Watson-averaged tortuous zeppelin."""
    fw = _watson_coeffs(kappa)
    l = np.arange(0, LMAX + 1, 2)
    tau1 = _watson_tau1(kappa)
    d_perp = d_par * (1.0 - v_ic)
    dw_par = d_par * tau1 + d_perp * (1.0 - tau1)
    dw_perp = d_par * (1.0 - tau1) / 2.0 + d_perp * (1.0 + tau1) / 2.0
    out = []
    for sh in scheme.shells:
        b = sh["b"]
        ks = _legendre_coeffs(np.exp(-b * d_par * _GL_X ** 2))
        a_ic = fw * ks * 4.0 * math.pi / (2 * l + 1)
        a_ec = _legendre_coeffs(np.exp(-b * (dw_perp + (dw_par - dw_perp) * _GL_X ** 2)))
        a = v_ic * a_ic + (1.0 - v_ic) * a_ec

        def f(t, a=a):
            o = np.zeros_like(t)
            for k, ll in enumerate(range(0, LMAX + 1, 2)):
                o += a[k] * special.eval_legendre(ll, t)
            return o

        f.coeffs = a
        out.append(f)
    return _Atom(out, False)


def atom_cylinder(scheme, d_par, R):
    D = d_par * 1e-6  # mm^2/s -> m^2/s
    out = []
    for sh in scheme.shells:
        _, G, Delta, delta = _shell_par(sh)
        s = _gpd_sum(_CYL_AM, Delta, delta, D, R, 1)
        q2 = (GAMMA * delta * G) ** 2

        def prof(t, G=G, s=s, q2=q2, Delta=Delta, delta=delta):
            return np.exp(-2 * GAMMA ** 2 * G ** 2 * (1 - t * t) * s) * np.exp(-(Delta - delta / 3.0) * q2 * t * t * D)

        out.append(_band_limit(prof))
    return _Atom(out, False)


def atom_sphere(scheme, d_is, R):
    D = d_is * 1e-6
    out = []
    for sh in scheme.shells:
        _, G, Delta, delta = _shell_par(sh)
        out.append(math.exp(-2 * GAMMA ** 2 * G ** 2 * _gpd_sum(_SPH_AM, Delta, delta, D, R, 2)))
    return _Atom(out, True)


def atom_astrosticks(scheme, d):
    out = []
    for sh in scheme.shells:
        x = math.sqrt(sh["b"] * d)
        out.append(math.sqrt(math.pi) / (2 * x) * math.erf(x))
    return _Atom(out, True)


def _sample(atom, scheme, dirs):
    """float32 (ndirs, nS) for a rotated atom, (nS,) for an isotropic one; b0 rows are 1."""
    if atom.isotropic:
        k = np.ones(scheme.nS, dtype=np.float32)
        for sh, v in zip(scheme.shells, atom.per_shell):
            k[sh["idx"]] = np.float32(v)
        return k
    k = np.ones((len(dirs), scheme.nS), dtype=np.float32)
    for sh, f in zip(scheme.shells, atom.per_shell):
        t = np.clip(dirs @ sh["grad"].T, -1.0, 1.0)
        k[:, sh["idx"]] = f(t).astype(np.float32)
    return k
