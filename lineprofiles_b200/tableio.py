"""Flat little-endian table container (".klpt") read by klp_table_load (include/klineprofiles_b200.h).

Layout: magic b"KLPT0001", int32 n_a, n_mu, n_r, n_g, f32 spins[n_a], f32 mus[n_mu], then per frame
(a-major, mu-minor; line-profile.zig:60-70) f32 radii[n_r], gmin[n_r], gmax[n_r], upper[n_r][n_g],
lower[n_r][n_g] -- the columns of one reference BinaryTable HDU (io.zig:22-49).
"""
from __future__ import annotations

import numpy as np

MAGIC = b"KLPT0001"


def save_table(path: str, t: dict) -> None:
    n_a, n_mu, n_r, n_g = (int(t[k]) for k in ("n_a", "n_mu", "n_r", "n_g"))
    with open(path, "wb") as f:
        f.write(MAGIC)
        np.array([n_a, n_mu, n_r, n_g], "<i4").tofile(f)
        t["spins"].astype("<f4").tofile(f)
        t["mus"].astype("<f4").tofile(f)
        for fr in range(n_a * n_mu):
            for k in ("radii", "gmin", "gmax", "upper", "lower"):
                np.ascontiguousarray(t[k][fr], "<f4").tofile(f)


def load_table(path: str) -> dict:
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a KLPT0001 table")
        n_a, n_mu, n_r, n_g = (int(v) for v in np.fromfile(f, "<i4", 4))
        spins = np.fromfile(f, "<f4", n_a)
        mus = np.fromfile(f, "<f4", n_mu)
        F = n_a * n_mu
        per = n_r * (2 * n_g + 3)
        body = np.fromfile(f, "<f4", F * per).reshape(F, per)
    o = 0
    out = dict(spins=spins, mus=mus, n_a=n_a, n_mu=n_mu, n_r=n_r, n_g=n_g)
    for k, n in (("radii", n_r), ("gmin", n_r), ("gmax", n_r), ("upper", n_r * n_g), ("lower", n_r * n_g)):
        out[k] = np.ascontiguousarray(body[:, o:o + n])
        o += n
    out["upper"] = out["upper"].reshape(F, n_r, n_g)
    out["lower"] = out["lower"].reshape(F, n_r, n_g)
    return out
