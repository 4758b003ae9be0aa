"""Vegetation-parameter generators with the call signatures of the microctools functions the reference's documented workflows use
(running-microclimc.md:229-382; R/code.R:1110-1113 `habitatvars(...)` in the runmodel example): PAIgeometry, thickgeometry, LAIfrac,
iwgeometry, PAIfromhabitat, habitatvars and the `habitats` table.

microctools' sources are not part of the reference (DESCRIPTION:23-24 pulls it from GitHub without a pin), so the ALGORITHMS here are
designed in this repository ([D] in the provenance tags of SURVEY.md Appendix A): they reproduce what the vignette documents — argument
meaning, return shapes (21-field vegp; PAI / pLAI as m x T matrices for a vector of times), the two printed values for "Deciduous
broadleaf forest" (height 15 m, x = 1.2), profiles with an understorey for forests, a MODIS-like seasonal cycle that follows latitude —
and are deterministic.  They are host-side input shaping, not part of the hot path; their output feeds runmodel / runmodelS (host mirror
or R wrappers) unchanged, including through the time-varying path (mc_set_vegt)."""
import datetime as _dt

import numpy as np

__all__ = ["habitats", "PAIgeometry", "thickgeometry", "LAIfrac", "iwgeometry", "PAIfromhabitat", "habitatvars"]

# descriptor -> (height m, x, PAI winter, PAI summer, evergreen, understorey fraction, leaf width m, gsmax, refls, clump)
_HAB = {
    "Evergreen needleleaf forest": (20.0, 0.4, 3.5, 4.5, True, 0.10, 0.01, 0.33, 0.30, 0.10),
    "Evergreen broadleaf forest": (25.0, 1.2, 5.0, 5.5, True, 0.15, 0.08, 0.33, 0.35, 0.00),
    "Deciduous needleleaf forest": (15.0, 0.4, 0.8, 3.5, False, 0.10, 0.01, 0.33, 0.30, 0.10),
    "Deciduous broadleaf forest": (15.0, 1.2, 0.9, 5.0, False, 0.15, 0.05, 0.33, 0.40, 0.00),
    "Mixed forest": (18.0, 0.8, 2.0, 4.8, False, 0.12, 0.03, 0.33, 0.35, 0.05),
    "Closed shrublands": (2.0, 1.0, 1.2, 2.5, False, 0.0, 0.02, 0.35, 0.35, 0.10),
    "Open shrublands": (1.5, 0.7, 0.5, 1.2, False, 0.0, 0.02, 0.35, 0.35, 0.30),
    "Woody savannas": (8.0, 0.7, 1.0, 2.2, False, 0.25, 0.03, 0.33, 0.35, 0.20),
    "Savannas": (3.0, 0.15, 0.6, 1.6, False, 0.35, 0.02, 0.33, 0.35, 0.30),
    "Short grasslands": (0.25, 0.15, 0.7, 1.8, False, 0.0, 0.01, 0.33, 0.35, 0.00),
    "Tall grasslands": (1.5, 0.15, 1.0, 3.0, False, 0.0, 0.01, 0.33, 0.35, 0.00),
    "Permanent wetlands": (0.5, 1.4, 0.8, 2.5, False, 0.0, 0.05, 0.55, 0.40, 0.00),
    "Croplands": (0.8, 0.2, 0.2, 3.5, False, 0.0, 0.03, 0.33, 0.40, 0.00),
    "Urban and built-up": (1.5, 1.0, 0.4, 0.9, False, 0.0, 0.04, 0.33, 0.35, 0.40),
    "Cropland/Natural vegetation mosaic": (1.2, 0.5, 0.5, 2.8, False, 0.0, 0.03, 0.33, 0.40, 0.10),
    "Barren or sparsely vegetated": (0.15, 0.6, 0.05, 0.15, False, 0.0, 0.015, 0.33, 0.35, 0.50),
}
habitats = [dict(number=i + 1, descriptor=k) for i, k in enumerate(_HAB)]


def _habitat(h):
    if isinstance(h, (int, np.integer)):
        h = habitats[int(h) - 1]["descriptor"]
    if h not in _HAB:
        raise ValueError("unknown habitat %r (see vegetation.habitats)" % (h,))
    return h, _HAB[h]


def PAIgeometry(m, PAI, skew, spread):
    """Plant area index of each of `m` canopy nodes (bottom to top), summing to `PAI`: a beta-shaped profile whose mode moves up with
    `skew` (1..10; 7 puts the bulk of the foliage in the upper canopy, as in the vignette's second plot) and which narrows as `spread`
    falls (1..100)."""
    m = int(m)
    z = (np.arange(1, m + 1) - 0.5) / m
    mode = np.clip(0.08 * float(skew) + 0.12, 0.05, 0.95)
    conc = 2.0 + 400.0 / max(float(spread), 1.0)               # concentration of the beta density
    a, b = 1 + mode * conc, 1 + (1 - mode) * conc
    w = z ** (a - 1) * (1 - z) ** (b - 1)
    return float(PAI) * w / w.sum()


def thickgeometry(m, thickw, hprop=0.7, thinprop=0.1):
    """Mean thickness (m) of the vegetation elements of each node: `thickw` (trunks, lower branches) below the relative height `hprop`,
    tapering to `thinprop * thickw` (twigs, leaves) at the canopy top."""
    z = (np.arange(1, int(m) + 1) - 0.5) / int(m)
    t = np.where(z <= hprop, 1.0, 1.0 - (1.0 - thinprop) * (z - hprop) / max(1.0 - hprop, 1e-9))
    return float(thickw) * np.maximum(t, thinprop)


def LAIfrac(m, pLAI, skew):
    """Fraction of the plant area that is green leaf in each node, with canopy mean `pLAI`: low near the ground (stems, litter) and
    saturating towards the top, the faster the larger `skew`."""
    z = (np.arange(1, int(m) + 1) - 0.5) / int(m)
    f = 1.0 - np.exp(-float(skew) * z)
    f = f * (float(pLAI) / f.mean())
    return np.clip(f, 0.0, 1.0)


def iwgeometry(m, iwmin=0.36, iwmax=0.9):
    """Relative turbulence intensity of each node (bottom to top); defaults after the maize canopy of Shaw et al. (1974): strongest in the
    lower canopy, falling towards the canopy top."""
    z = (np.arange(1, int(m) + 1) - 0.5) / int(m)
    return iwmin + (iwmax - iwmin) * (1 - z) ** 1.5


def _hours_of_year(year):
    t0 = _dt.datetime(int(year), 1, 1)
    n = (_dt.datetime(int(year) + 1, 1, 1) - t0).days * 24
    return [t0 + _dt.timedelta(hours=h) for h in range(n)]


def _season(doy, lat, evergreen):
    """0..1 seasonal factor: leaf-on fraction from a smooth growing season whose centre and length follow the latitude (day 196 in the
    north, 14 in the south; ~200 days at 50 degrees, all year in the tropics)."""
    lat = float(lat)
    centre = 196.0 if lat >= 0 else 14.0
    length = np.clip(365.0 - 3.3 * max(abs(lat) - 10.0, 0.0), 60.0, 365.0)
    d = np.abs(((np.asarray(doy, float) - centre + 182.5) % 365.0) - 182.5)      # days from the centre of the season
    s = 1.0 / (1.0 + np.exp((d - length / 2.0) / 12.0))
    if evergreen or abs(lat) < 10:
        s = 0.8 + 0.2 * s
    return s


def PAIfromhabitat(habitat, lat, long, year):
    """Hourly total plant area index over `year` for a habitat type, plus canopy height and the leaf-angle coefficient x:
    dict(lai [hours], obs_time [hours], height, x)."""
    name, (hgt, x, p0, p1, ever, *_rest) = _habitat(habitat)
    tme = _hours_of_year(year)
    doy = np.array([t.timetuple().tm_yday + t.hour / 24.0 for t in tme])
    lai = p0 + (p1 - p0) * _season(doy, lat, ever)
    return dict(lai=lai, obs_time=[t.strftime("%Y-%m-%d %H:%M:%S") for t in tme], height=hgt, x=x)


def _as_hours(tme):
    """tme (datetime, time string, or a sequence of them) -> list of hourly datetimes from 00:00 of the first day to 23:00 of the last
    (a single time stays a single time), as the vignette describes."""
    def one(t):
        if isinstance(t, _dt.datetime):
            return t
        for f in ("%Y-%m-%d %H:%M:%S", "%Y-%m-%d %H:%M", "%Y-%m-%d"):
            try:
                return _dt.datetime.strptime(str(t), f)
            except ValueError:
                pass
        raise ValueError("cannot read time %r" % (t,))
    if isinstance(tme, (str, _dt.datetime)):
        return [one(tme)], True
    ts = [one(t) for t in tme]
    if len(ts) == 1:
        return ts, True
    hourly = all((b - a) == _dt.timedelta(hours=1) for a, b in zip(ts[:-1], ts[1:]))
    if hourly:
        return ts, False
    t0 = ts[0].replace(hour=0, minute=0, second=0); t1 = ts[-1].replace(hour=23, minute=0, second=0)
    n = int((t1 - t0).total_seconds() // 3600) + 1
    return [t0 + _dt.timedelta(hours=h) for h in range(n)], False


def habitatvars(habitat, lat, long, tme, m, PAIt=None):
    """All 21 vegetation parameters of the model from a habitat type (running-microclimc.md:367-370): hgt, PAI, x, lw, cd, iw, hgtg, zm0,
    pLAI, refls, refg, refw, reflp, vegem, gsmax, q50, thickw, cpw, phw, kwood, clump.  A single time gives PAI / pLAI as vectors of
    length m; a vector of times gives m x hours matrices (what .vegpsort indexes, R/code.R:904-916).  `PAIt`: None (from PAIfromhabitat), one
    value (vectors returned whatever `tme`), m values (a per-node profile, seasonally adjusted) or any other length (total PAI over the
    period, interpolated to the hours)."""
    name, (hgt, x, p0, p1, ever, under, lw, gsmax, refls, clump) = _habitat(habitat)
    m = int(m)
    ts, single = _as_hours(tme)
    doy = np.array([t.timetuple().tm_yday + t.hour / 24.0 for t in ts])
    seas = _season(doy, lat, ever)
    profile = None
    if PAIt is None:
        tot = p0 + (p1 - p0) * seas
    else:
        P = np.atleast_1d(np.asarray(PAIt, float))
        if P.size == 1:
            tot = np.array([P[0]]); single = True
        elif P.size == m:
            profile = P / P.sum(); tot = P.sum() * (seas / seas.max())
        else:
            tot = np.interp(np.linspace(0, 1, len(ts)), np.linspace(0, 1, P.size), P)
    forest = hgt >= 5.0
    crown = PAIgeometry(m, 1.0, 7 if forest else 4, 70 if forest else 90)
    if profile is None:
        ush = PAIgeometry(m, 1.0, 0.5, 25) if under > 0 else np.zeros(m)       # understorey near the ground
        profile = (1 - under) * crown + under * ush
    green_crown = LAIfrac(m, 0.8 if forest else 0.9, 5)
    if single:
        PAI = profile * tot[0]
        pLAI = green_crown * (0.35 + 0.65 * seas[0]) if not ever else green_crown
    else:
        PAI = profile[:, None] * tot[None, :]
        leaf_on = seas if not ever else np.ones_like(seas)
        pLAI = green_crown[:, None] * (0.35 + 0.65 * leaf_on[None, :])
    return dict(hgt=hgt, PAI=PAI, x=x, lw=lw, cd=0.2, iw=iwgeometry(m), hgtg=min(0.05 * hgt, 0.5) if forest else hgt, zm0=0.004,
                pLAI=np.clip(pLAI, 0.0, 1.0), refls=refls, refg=0.15, refw=0.15, reflp=refls, vegem=0.97, gsmax=gsmax, q50=100.0,
                thickw=thickgeometry(m, 0.4 if forest else 0.01, 0.7, 0.1) * (0.1 if forest else 1.0) + 0.002, cpw=2500.0, phw=700.0,
                kwood=0.3, clump=clump)
