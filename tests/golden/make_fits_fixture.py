#!/usr/bin/env python
"""A small FITS file in the schema of the reference's kerr-transfer-functions.fits (io.zig:8-87), written BYTE BY BYTE
from the FITS standard (80-character cards, 2880-byte blocks, big-endian binary tables) -- deliberately NOT through
lineprofiles_b200.tableio.save_fits, so that the C++ reader (klp_table_load) is checked against an independent writer
and against the features a cfitsio-written file shows that save_fits never emits:

  * TFORM4 = '30E' / TFORM5 = '30E' vector columns with TDIM, TTYPE, TUNIT cards (the real table's layout);
  * D-typed (f64) scalar columns (gmax of every frame; the spin axis of HDU 2), narrowed by the reader;
  * an extra, non-float column after the ones the reference reads (HDU 2: a J index column) that must be skipped;
  * EXTNAME / EXTVER / COMMENT / HISTORY cards, a string value with an embedded '' quote, so that a frame header spans
    more than 36 cards (two header blocks);
  * a primary HDU with extra cards.

Usage: python tests/golden/make_fits_fixture.py   -> tests/golden/fixture_table.fits (committed; the test checks that the
committed bytes equal a fresh run).  `fixture_arrays()` are the values the file holds."""
from __future__ import annotations

import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "fixture_table.fits")
N_A, N_MU, N_R, N_G = 3, 2, 5, 30


def fixture_arrays() -> dict:
    spins = np.array([0.0, 0.5, 0.998], np.float32)
    mus = np.array([0.1, 0.9], np.float32)
    F = N_A * N_MU
    f = np.arange(F, dtype=np.float64)[:, None]
    j = np.arange(N_R, dtype=np.float64)[None, :]
    k = np.arange(N_G, dtype=np.float64)[None, None, :]
    radii = (40.0 * 0.55 ** j + 1.5 + 0.01 * f).astype(np.float32)          # descending
    gmin = (0.35 + 0.02 * j + 0.003 * f).astype(np.float32)
    gmax = (1.05 + 0.03 * j + 0.002 * f)                                     # stored as f64 ('D'), narrowed by the reader
    upper = (1.0 + 0.1 * f[:, :, None] + 0.01 * j[:, :, None] + 0.002 * k).astype(np.float32)
    lower = (0.7 + 0.05 * f[:, :, None] + 0.02 * j[:, :, None] + 0.001 * k * k).astype(np.float32)
    return dict(spins=spins, mus=mus, n_a=N_A, n_mu=N_MU, n_r=N_R, n_g=N_G, radii=radii, gmin=gmin,
                gmax=gmax.astype(np.float32), gmax_f64=gmax, upper=upper, lower=lower)


def card(key: str, value=None, comment: str = "") -> bytes:
    """One 80-character header card (FITS standard 4.1-4.2: keyword in columns 1-8, '= ' in 9-10, fixed-format value)."""
    if value is None:                                   # commentary card: COMMENT / HISTORY / blank
        s = f"{key:<8}{comment}"
    else:
        if isinstance(value, bool):
            v = ("T" if value else "F").rjust(20)
        elif isinstance(value, int):
            v = str(value).rjust(20)
        elif isinstance(value, float):
            v = f"{value:.10E}".rjust(20)
        else:                                           # string: quote, at least 8 characters, '' escapes a quote
            v = "'" + str(value).replace("'", "''").ljust(8) + "'"
            v = v.ljust(20)
        s = f"{key:<8}= {v}"
        if comment:
            s += " / " + comment
    assert len(s) <= 80, s
    return s.ljust(80).encode("ascii")


def block(b: bytes, fill: bytes) -> bytes:
    return b + fill * ((-len(b)) % 2880)


def bintable(name: str, columns, extra_cards=()) -> bytes:
    """columns: (ttype, tform letter, unit, array [rows] or [rows, repeat]).  Row-major, big-endian (FITS standard 7.3)."""
    rows = columns[0][3].shape[0]
    fmt = {"E": ">f4", "D": ">f8", "J": ">i4"}
    width = {"E": 4, "D": 8, "J": 4}
    reps = [1 if a.ndim == 1 else a.shape[1] for _, _, _, a in columns]
    naxis1 = sum(width[t] * r for (_, t, _, _), r in zip(columns, reps))
    cards = [card("XTENSION", "BINTABLE", "binary table extension"), card("BITPIX", 8, "8-bit bytes"),
             card("NAXIS", 2, "2-dimensional binary table"), card("NAXIS1", naxis1, "width of table in bytes"),
             card("NAXIS2", rows, "number of rows in table"), card("PCOUNT", 0, "size of special data area"),
             card("GCOUNT", 1, "one data group (required keyword)"), card("TFIELDS", len(columns), "number of fields in each row")]
    for i, ((ttype, t, unit, a), r) in enumerate(zip(columns, reps), 1):
        cards.append(card(f"TTYPE{i}", ttype, "label for field"))
        cards.append(card(f"TFORM{i}", (str(r) if r > 1 else ("1" if t == "D" else "")) + t, "data format of field"))
        if unit:
            cards.append(card(f"TUNIT{i}", unit, "physical unit of field"))
        if r > 1:
            cards.append(card(f"TDIM{i}", f"({r})", "array dimensions"))
    cards.append(card("EXTNAME", name, "name of this binary table extension"))
    cards.append(card("EXTVER", 1))
    cards += list(extra_cards)
    hdr = block(b"".join(cards) + b"END".ljust(80), b" ")
    rows_b = []
    for r_ in range(rows):
        for (_, t, _, a) in columns:
            rows_b.append(np.asarray(a[r_]).astype(fmt[t]).tobytes())
    data = b"".join(rows_b)
    assert len(data) == naxis1 * rows
    return hdr + block(data, b"\0")


def build() -> bytes:
    t = fixture_arrays()
    out = [block(b"".join([card("SIMPLE", True, "file does conform to FITS standard"), card("BITPIX", 8), card("NAXIS", 0),
                           card("EXTEND", True), card("ORIGIN", "klp fixture"), card("COMMENT", None, "  hand-written fixture; schema of io.zig:8-87"),
                           card("HISTORY", None, " TFORM1  = 'not a real card' (commentary text must be ignored)")])
                 + b"END".ljust(80), b" ")]
    out.append(bintable("SPINS", [("a", "D", "", t["spins"].astype(np.float64)), ("index", "J", "", np.arange(N_A, dtype=np.int32))]))
    out.append(bintable("INCLINATIONS", [("mu", "E", "", t["mus"])]))
    filler = [card("COMMENT", None, f"  filler line {i:02d} so that the header needs a second 2880-byte block") for i in range(22)]
    filler.append(card("OBSERVER", "O'Neill", "a string value with an embedded quote"))
    for f in range(N_A * N_MU):
        out.append(bintable(f"TF{f:03d}", [("r", "E", "rg", t["radii"][f]), ("gmin", "E", "", t["gmin"][f]), ("gmax", "D", "", t["gmax_f64"][f]),
                                          ("upper_f", "E", "", t["upper"][f]), ("lower_f", "E", "", t["lower"][f])],
                            extra_cards=[card("SPIN", float(t["spins"][f // N_MU])), card("MU", float(t["mus"][f % N_MU]))] + filler))
    return b"".join(out)


if __name__ == "__main__":
    b = build()
    with open(PATH, "wb") as fh:
        fh.write(b)
    print(PATH, len(b), "bytes")
