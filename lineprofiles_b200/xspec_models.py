"""Model definitions of the XSPEC surface: names, parameter order, units, defaults and ranges.

Same content as the reference's xspec/lmodel.dat (lines 1-59); kept as data so that the generated
lmodel.dat, the ctypes layer and the tests share one source.  Each parameter:
(name, unit, default, hard_min, soft_min, soft_max, hard_max, delta)."""
from __future__ import annotations

_A = ("a", '" "', "0.998", "-0.998", "-0.998", "0.998", "0.998", "0.001")
_INCL = ("incl", "deg", "60", "0.", "0.", "90.", "90.", "1")
_ELINE = ("eline", "keV", "6.4", "0.", "0.", "10.", "10.", "0.1")
_RMIN = ("rmin", '" "', "1.0", "1.0", "1.0", "100.0", "100.0", "0.1")
_RMAX = ("rmax", '" "', "400.0", "1.0", "1.0", "400.0", "1000.0", "0.1")
_RCUT = ("rcut", '" "', "50.0", "1.0", "1.0", "400.0", "1000.0", "0.1")
_ALPHA = ("alpha", '" "', "3.0", "0.", "0.", "10", "10", "1")


def _e(i):
    return (f"e{i}", '" "', "3.0", "0.", "0.", "1e10", "1e10", "1")


MODELS = {
    # name: (kind, parameters)   kind: "add" additive, "con" convolution  (lmodel.dat:1,9,23,30,43)
    "kline": ("add", [_A, _INCL, _ELINE, _RMIN, _RMAX, _ALPHA]),
    "kline5": ("add", [_A, _INCL, _ELINE, _RMIN, _RMAX, _RCUT, _ALPHA] + [_e(i) for i in range(1, 6)]),
    "kconv": ("con", [_A, _INCL, _RMIN, _RMAX, _ALPHA]),
    "kconv5": ("con", [_A, _INCL, _RMIN, _RMAX, _RCUT, _ALPHA] + [_e(i) for i in range(1, 6)]),
    "kconv10": ("con", [_A, _INCL, _RMIN, _RMAX, _RCUT, _ALPHA] + [_e(i) for i in range(1, 11)]),
}


def lmodel_text() -> str:
    out = []
    for name, (kind, pars) in MODELS.items():
        out.append(f"{name}     {len(pars)}     \t0.  1.e20\tc_{name}\t{kind}    \t0 \t\t0")
        for p in pars:
            out.append("   ".join(p[:2]) + "     " + "  ".join(p[2:]))
        out.append("")
    return "\n".join(out)


def defaults(name: str):
    return [float(p[2]) for p in MODELS[name][1]]
