"""Acquisition scheme: the attributes of ``amico/scheme.py:21-154`` that the fit path and its callers read."""
from __future__ import annotations

import numpy as np

GAMMA = 2.675987e8  # proton gyromagnetic ratio [rad/(s T)]


class Scheme:
    """Acquisition scheme: Nx4 (dir, b) or Nx7 (dir, G, Delta, delta, TE) table.

    Mirrors the attributes of ``amico/scheme.py:50-135`` consumed downstream.
    """

    def __init__(self, raw, b0_thr=0.0):
        raw = np.array(raw, dtype=np.float64)
        if raw.ndim != 2 or raw.shape[1] not in (4, 7):
            raise ValueError("Unrecognized scheme format")
        self.raw = raw
        if raw.shape[1] == 4:
            self.version = 0
            self.b = raw[:, 3].copy()
        else:
            self.version = 1
            self.b = (GAMMA * raw[:, 3] * raw[:, 5]) ** 2 * (raw[:, 4] - raw[:, 5] / 3.0) * 1e-6
        self.b0_thr = b0_thr
        self.b0_idx = np.where(self.b <= b0_thr)[0]
        self.b0_count = len(self.b0_idx)
        self.dwi_idx = np.where(self.b > b0_thr)[0]
        self.dwi_count = len(self.dwi_idx)
        flip = self.raw[:, 1] < 0
        self.raw[flip, 0:3] *= -1.0
        self.shells = []
        par = np.ascontiguousarray(self.raw[:, 3:])
        seen = []
        for i in range(par.shape[0]):
            if self.b[i] <= b0_thr:
                continue
            key = tuple(par[i])
            if key in seen:
                continue
            seen.append(key)
            idx = np.where((par == par[i]).all(axis=1))[0]
            sh = {"b": self.b[i], "idx": idx, "grad": self.raw[idx, 0:3]}
            if self.version == 1:
                sh.update(G=par[i, 0], Delta=par[i, 1], delta=par[i, 2], TE=par[i, 3])
            else:
                sh.update(G=None, Delta=None, delta=None, TE=None)
            self.shells.append(sh)

    @property
    def nS(self):
        return self.b0_count + self.dwi_count
