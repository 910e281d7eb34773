"""Synthetic CNT film generator and mesh-file I/O.

The reference ships no mesh data: ``input.json:41`` points at ``~/research/mesh/cnt_mesh_fiber``, six Armadillo
matrix files produced by a different project.  This module produces deterministic stand-in films
in exactly that on-disk format so that the reference program, the oracle and the CUDA engine all read the same
bytes (``monte_carlo.h:199-271`` is the reader on the reference side).

Generator: a 64-bit Mersenne Twister ``mt19937_64(seed)`` with ``u = (g() >> 11) * 2**-53``;
per tube: centre ``(u*LX, u*LY, u*LX)``, in-plane angle ``phi = 2*pi*u``, tilt ``t = (u - 0.5) * 0.1`` rad, unit
orientation ``o = (cos(phi)cos(t), sin(t), sin(phi)cos(t))``; site ``p`` sits at ``c + (p - (NP-1)/2) * a * o``.
Positions are in nm (the reference multiplies by 1e-9 on load), every site of a tube carries the tube orientation.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

MESH_FILES = tuple(f"single_cnt.{kind}.{ax}.dat" for kind in ("pos", "orient") for ax in "xyz")


class MT19937_64:
    """std::mt19937_64 (Matsumoto & Nishimura 2004), pure Python; only a few thousand draws are ever needed."""

    NN, MM = 312, 156
    MATRIX_A = 0xB5026F5AA96619E9
    UM, LM = 0xFFFFFFFF80000000, 0x7FFFFFFF
    MASK = 0xFFFFFFFFFFFFFFFF

    def __init__(self, seed: int = 5489):
        mt = [0] * self.NN
        mt[0] = seed & self.MASK
        for i in range(1, self.NN):
            mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & self.MASK
        self.mt, self.mti = mt, self.NN

    def _twist(self) -> None:
        mt, NN, MM = self.mt, self.NN, self.MM
        for i in range(NN):
            x = (mt[i] & self.UM) | (mt[(i + 1) % NN] & self.LM)
            mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (self.MATRIX_A if (x & 1) else 0)
        self.mti = 0

    def __call__(self) -> int:
        if self.mti >= self.NN:
            self._twist()
        x = self.mt[self.mti]
        self.mti += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & self.MASK

    def uniform(self) -> float:
        return (self() >> 11) * 2.0 ** -53


def film(NT: int, NP: int, a: float, LX: float, LY: float, seed: int = 1234) -> Tuple[np.ndarray, np.ndarray]:
    """Return ``(pos, orient)`` each of shape ``(3, NT, NP)``; positions in nm."""
    g = MT19937_64(seed)
    u = np.array([[g.uniform() for _ in range(5)] for _ in range(NT)], dtype=np.float64)
    c = np.stack([u[:, 0] * LX, u[:, 1] * LY, u[:, 2] * LX])  # (3, NT)
    phi = 2.0 * np.pi * u[:, 3]
    tilt = (u[:, 4] - 0.5) * 0.1
    o = np.stack([np.cos(phi) * np.cos(tilt), np.sin(tilt), np.sin(phi) * np.cos(tilt)])  # (3, NT)
    s = (np.arange(NP, dtype=np.float64) - (NP - 1) / 2.0) * a  # (NP,)
    pos = c[:, :, None] + s[None, None, :] * o[:, :, None]
    orient = np.broadcast_to(o[:, :, None], (3, NT, NP)).copy()
    return pos, orient


def lattice_film(n: int, a: float) -> Tuple[np.ndarray, np.ndarray]:
    """``n**3`` link-less sites on a cubic lattice of pitch ``a`` nm (one site per 'tube', so left=right=-1).

    With a constant-rate table this is a pure continuous-time random walk with the analytic diffusion coefficient
    ``D = r*a**2`` per axis for 6 nearest neighbours (SURVEY.md §8c pin 6).  Orientations alternate so that no two
    neighbouring sites are exactly parallel or antiparallel (the reference has no guard for those cases).
    """
    idx = np.arange(n, dtype=np.float64) * a
    x, y, z = np.meshgrid(idx, idx, idx, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()])[:, :, None]  # (3, n^3, 1)
    k = np.arange(n ** 3, dtype=np.float64)
    ang = 0.1 + 0.37 * k
    tilt = 0.05 * np.sin(1.3 * k)
    o = np.stack([np.cos(ang) * np.cos(tilt), np.sin(tilt), np.sin(ang) * np.cos(tilt)])
    o /= np.sqrt((o * o).sum(axis=0))
    return pos, o[:, :, None].copy()


def write_mesh(directory: str, pos: np.ndarray, orient: np.ndarray) -> None:
    """Write the six ``single_cnt.{pos,orient}.{x,y,z}.dat`` files as Armadillo ``arma_ascii`` matrices.

    Format read by ``arma::mat::load`` (auto-detect) and by ``visualization/monte_carlo_results.py:163-168`` (which
    skips two header lines): ``ARMA_MAT_TXT_FN008\\n<rows> <cols>\\n`` then one matrix row per line.  17 significant
    digits make the text round-trip exactly to the same doubles in every reader.
    """
    os.makedirs(directory, exist_ok=True)
    for kind, arr in (("pos", pos), ("orient", orient)):
        for ia, ax in enumerate("xyz"):
            m = np.asarray(arr[ia], dtype=np.float64)
            with open(os.path.join(directory, f"single_cnt.{kind}.{ax}.dat"), "w") as f:
                f.write("ARMA_MAT_TXT_FN008\n%d %d\n" % m.shape)
                for row in m:
                    f.write(" ".join("%.16e" % v for v in row))
                    f.write("\n")


def read_mesh(directory: str) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of :func:`write_mesh` (also accepts raw ASCII without the Armadillo header)."""
    out = []
    for kind in ("pos", "orient"):
        comps = []
        for ax in "xyz":
            path = os.path.join(directory, f"single_cnt.{kind}.{ax}.dat")
            with open(path) as f:
                first = f.readline()
                if first.startswith("ARMA_MAT"):
                    r, c = (int(t) for t in f.readline().split())
                    m = np.loadtxt(f, dtype=np.float64, ndmin=2).reshape(r, c)
                else:
                    f.seek(0)
                    m = np.loadtxt(f, dtype=np.float64, ndmin=2)
            comps.append(m)
        out.append(np.stack(comps))
    return out[0], out[1]


# Named films of BASELINE.json's configs.  LX also spans z.
CONFIG_FILMS = {
    "C1": dict(NT=200, NP=100, a=5.0, LX=400.0, LY=100.0, seed=1234),
    "C2": dict(NT=1000, NP=100, a=5.0, LX=1000.0, LY=100.0, seed=1234),
    "C4": dict(NT=20000, NP=250, a=2.0, LX=2000.0, LY=100.0, seed=1234),
}
