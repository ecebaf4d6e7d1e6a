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


# ---- FITS in the reference's schema (io.zig:8-87), written without astropy/cfitsio ----------------

def _card(key: str, value, comment: str = "") -> bytes:
    if isinstance(value, bool):
        v = ("T" if value else "F").rjust(20)
    elif isinstance(value, int):
        v = str(value).rjust(20)
    else:
        v = ("'" + str(value).ljust(8) + "'").ljust(20)
    s = f"{key:<8}= {v}"
    if comment:
        s += " / " + comment
    return s[:80].ljust(80).encode("ascii")


def _pad(b: bytes, fill: bytes) -> bytes:
    return b + fill * ((-len(b)) % 2880)


def _bintable(columns) -> bytes:
    """columns: list of (name, array[rows] or array[rows, repeat]); written as big-endian float32 (TFORM rE)."""
    rows = columns[0][1].shape[0]
    widths = [(1 if a.ndim == 1 else a.shape[1]) for _, a in columns]
    cards = [_card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", 4 * sum(widths)),
             _card("NAXIS2", rows), _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(columns))]
    for i, ((name, _), w) in enumerate(zip(columns, widths), 1):
        cards += [_card(f"TTYPE{i}", name), _card(f"TFORM{i}", f"{w}E")]
    hdr = _pad(b"".join(cards) + b"END".ljust(80), b" ")
    data = np.concatenate([np.asarray(a, np.float32).reshape(rows, -1) for _, a in columns], axis=1)
    return hdr + _pad(np.ascontiguousarray(data).astype(">f4").tobytes(), b"\0")   # FITS is big-endian


def save_fits(path: str, t: dict) -> None:
    """Write `t` as a FITS file laid out like the reference's kerr-transfer-functions.fits."""
    n_a, n_mu = int(t["n_a"]), int(t["n_mu"])
    with open(path, "wb") as f:
        f.write(_pad(b"".join([_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True)])
                     + b"END".ljust(80), b" "))
        f.write(_bintable([("a", np.asarray(t["spins"]))]))
        f.write(_bintable([("mu", np.asarray(t["mus"]))]))
        for fr in range(n_a * n_mu):
            f.write(_bintable([("r", t["radii"][fr]), ("gmin", t["gmin"][fr]), ("gmax", t["gmax"][fr]),
                               ("upper_f", t["upper"][fr]), ("lower_f", t["lower"][fr])]))
