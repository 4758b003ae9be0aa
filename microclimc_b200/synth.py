"""Synthetic inputs of the BASELINE configurations and host-side input shaping.

microctools' vegetation generators (habitatvars, PAIgeometry, ...) are not available (source absent from the
reference), so `vegp_deciduous` is a documented stand-in: hgt = 15 m and x = 1.2 are the only values the rendered
vignette prints (running-microclimc.md:346-349); the rest follow the ranges quoted in the vignette prose.
All generators are deterministic (PCG64 with an explicit seed).
"""
import json
import os
import numpy as np
from . import layout as L

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_weather():
    """The reference's `weather` dataset (R/data.R:49-65): dict of arrays, 8760 hourly rows."""
    z = np.load(os.path.join(_DATA, "weather.npz"))
    return {k: z[k] for k in z.files}


def load_dailyprecip():
    return np.load(os.path.join(_DATA, "dailyprecip.npy"))


def load_soilparams():
    return json.load(open(os.path.join(_DATA, "soilparams.json")))


def load_climvars():
    return json.load(open(os.path.join(_DATA, "climvars.json")))


def jday(year, month, day):
    """Astronomical Julian day number (microctools::jday semantics, integer day)."""
    a = (14 - month) // 12
    y = year + 4800 - a
    mm = month + 12 * a - 3
    return day + (153 * mm + 2) // 5 + 365 * y + y // 4 - y // 100 + y // 400 - 32045


def parse_obs_time(obs_time):
    """obs_time strings "YYYY-mm-dd HH:MM[:SS]" -> (jd[T], lt[T], epoch_seconds[T]); as.POSIXlt(..., tz="UTC")."""
    jd = np.empty(len(obs_time)); lt = np.empty(len(obs_time)); ep = np.empty(len(obs_time))
    for i, s in enumerate(obs_time):
        s = str(s)
        date, _, tm = s.partition(" ")
        y, mo, d = (int(v) for v in date.split("-"))
        parts = [float(v) for v in tm.split(":")] if tm else [0.0]
        parts += [0.0] * (3 - len(parts))
        jd[i] = jday(y, mo, d)
        lt[i] = parts[0] + parts[1] / 60 + parts[2] / 3600
        ep[i] = (jd[i] - 2440588) * 86400.0 + parts[0] * 3600 + parts[1] * 60 + parts[2]
    return jd, lt, ep


def forcing_from_climdata(climdata, theta=0.3):
    """climdata (dict like `weather`) -> forcing [T][NFORC] (layout.FORCING), timestep in seconds.

    dp = difrad/swrad with non-finite -> 0.5 and NO clamp to [0,1] (code.R:1166-1167, quirk Q19)."""
    jd, lt, ep = parse_obs_time(climdata["obs_time"])
    T = jd.size
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = np.asarray(climdata["difrad"], float) / np.asarray(climdata["swrad"], float)
    dp[~np.isfinite(dp)] = 0.5
    th = np.broadcast_to(np.asarray(theta, float), (T,)) if np.ndim(theta) <= 1 else np.asarray(theta, float)[0]
    f = np.empty((T, L.NFORC))
    f[:, 0] = climdata["temp"]; f[:, 1] = climdata["relhum"]; f[:, 2] = climdata["pres"]; f[:, 3] = climdata["windspeed"]
    f[:, 4] = climdata["skyem"]; f[:, 5] = climdata["swrad"]; f[:, 6] = dp; f[:, 7] = th; f[:, 8] = jd; f[:, 9] = lt
    timestep = float(round(ep[1] - ep[0])) if T > 1 else 3600.0
    return f, timestep


def steady_clim_from_climdata(climdata):
    jd, lt, _ = parse_obs_time(climdata["obs_time"])
    c = np.empty((jd.size, len(L.STEADY_CLIM)))
    for i, k in enumerate(["temp", "relhum", "pres", "windspeed", "swrad", "difrad", "skyem"]):
        c[:, i] = climdata[k]
    c[:, 7] = jd; c[:, 8] = lt
    return c


def synthetic_forcing(T, seed=7, lat=50.0):
    """Hourly synthetic forcing of the `weather` shape for arbitrary T (sinusoidal seasons/diurnal + noise)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h = np.arange(T)
    doy = (h // 24) % 365
    hr = h % 24
    season = -np.cos(2 * np.pi * (doy - 15) / 365.0)
    sun = np.maximum(0.0, np.sin(np.pi * (hr - 6 - 2 * (1 - season) / 2 + 0.0) / (12 + 4 * season * 0 + 0.0)))
    sun = np.where((hr > 6) & (hr < 18), np.sin(np.pi * (hr - 6) / 12.0), 0.0)
    cloud = np.clip(0.5 + 0.35 * np.sin(2 * np.pi * h / 97.0) + 0.15 * rng.standard_normal(T), 0.05, 1.0)
    swrad = (450 + 350 * season) * sun * (1 - 0.7 * cloud)
    temp = 11 + 7 * season + 4 * np.sin(2 * np.pi * (hr - 9) / 24.0) * (1 - 0.5 * cloud) + rng.standard_normal(T)
    relhum = np.clip(80 - 15 * sun * (1 - cloud) + 8 * rng.standard_normal(T), 30, 100)
    pres = 101.0 + 0.8 * np.sin(2 * np.pi * h / 211.0)
    wind = np.clip(4.5 + 2.5 * np.sin(2 * np.pi * h / 131.0) + 1.5 * rng.standard_normal(T), 0.1, 20)
    difrad = swrad * np.clip(0.25 + 0.7 * cloud, 0, 1)
    skyem = np.clip(0.7 + 0.28 * cloud, 0.6, 1.0)
    obs = []
    mdays = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    y, mo, d = 1995, 1, 1
    for i in range(T):
        obs.append("%04d-%02d-%02d %02d:00:00" % (y, mo, d, i % 24))
        if i % 24 == 23:
            d += 1
            if d > mdays[mo - 1]:
                d = 1; mo += 1
                if mo > 12:
                    mo = 1; y += 1
    return dict(obs_time=np.array(obs), temp=temp, relhum=relhum, pres=pres, swrad=swrad, difrad=difrad, skyem=skyem,
                windspeed=wind, winddir=np.zeros(T))


def theta_from_dailyprecip(dailyprecip, soil, temp, swrad, T=None, depth=0.1, theta0=None):
    """Hourly volumetric water content of the top soil layer from daily precipitation — a host-side single-bucket recipe for
    BASELINE config C4.  Soil moisture is an INPUT of the hot path: the reference has no soil-water model of its
    own (vignette: `theta` is supplied; NicheMapR provides it in runwithNMR, NMR.R:796), so this stands in for that provider.

    Bucket of `depth` m between Smin and Smax of `soil` (soilparams row): every hour
      theta += (P/24 - ET - D) / (1000 * depth)   [P in mm/day, ET and D in mm/h]
      ET = 0.0008 * max(swrad, 0) * max(temp, 0) / 15 * rel   (radiation- and temperature-scaled, rel = relative saturation)
      D  = Ksat/24 * rel ** (2 * b + 3)                       (Campbell unit-gradient drainage, Ksat in mm/day)
    clipped to [Smin, Smax] (excess runs off)."""
    P = np.asarray(dailyprecip, float)
    temp = np.asarray(temp, float); swrad = np.asarray(swrad, float)
    T = temp.size if T is None else T
    smin, smax, b, ksat = float(soil["Smin"]), float(soil["Smax"]), float(soil["b"]), float(soil["Ksat"])
    th = np.empty(T)
    cur = 0.5 * (smin + smax) if theta0 is None else float(theta0)
    for t in range(T):
        rel = (cur - smin) / (smax - smin)
        et = 0.0008 * max(swrad[t], 0.0) * max(temp[t], 0.0) / 15.0 * rel
        dr = ksat / 24.0 * rel ** (2 * b + 3)
        cur = min(smax, max(smin, cur + (P[(t // 24) % P.size] / 24.0 - et - dr) / (1000.0 * depth)))
        th[t] = cur
    return th


def snow_from_forcing(dailyprecip, temp, density=100.0, melt_per_degree_hour=0.2, seed=0):
    """Synthetic NicheMapR-style snow outputs for config C4: snow depth (cm) accumulates from precipitation falling at temp < 0.5 C
    (fresh-snow density `density` kg m-3) and melts by `melt_per_degree_hour` cm per degree-hour above 0 C; the eight snow-layer
    temperatures SN1..SN8 grade from the (sub-zero) air temperature at the surface to -0.5 C at the base.  Returns (SNOWDEP[T],
    SN[T][8])."""
    P = np.asarray(dailyprecip, float); temp = np.asarray(temp, float)
    T = temp.size
    dep = np.zeros(T); cur = 0.0
    for t in range(T):
        if temp[t] < 0.5:
            cur += P[(t // 24) % P.size] / 24.0 * (1000.0 / density) / 10.0       # mm water -> cm snow
        else:
            cur = max(0.0, cur - melt_per_degree_hour * temp[t])
        dep[t] = cur
    top = np.minimum(temp, -0.1)
    w = np.linspace(1.0, 0.0, 8)[None, :]
    return dep, top[:, None] * w + (-0.5) * (1 - w)


def vegp_deciduous(m=20, hgt=15.0, PAIt=4.0):
    """Synthetic 'deciduous broadleaf forest' vegp (dict with the 21 reference field names)."""
    zz = (np.arange(1, m + 1) - 0.5) / m
    bell = np.exp(-0.5 * ((zz - 0.7) / 0.18) ** 2) + 0.15 * np.exp(-0.5 * ((zz - 0.1) / 0.08) ** 2) + 0.05
    PAI = PAIt * bell / bell.sum()
    return dict(hgt=float(hgt), PAI=PAI, x=1.2, lw=0.05, cd=0.2, iw=np.linspace(0.8, 0.3, m), hgtg=0.05 * hgt,
                zm0=0.004, pLAI=np.full(m, 0.8), refls=0.4, refg=0.15, refw=0.15, reflp=0.4, vegem=0.97, gsmax=0.33,
                q50=100.0, thickw=np.linspace(0.4, 0.01, m) * 0.1 + 0.002, cpw=2500.0, phw=700.0, kwood=0.3, clump=0.0)


def pack_vegp(vegps, m):
    """list of vegp dicts (static PAI vectors) -> array [NVEG][ncol]."""
    rows = L.veg_rows(m)
    out = np.empty((L.nveg(m), len(vegps)))
    for c, v in enumerate(vegps):
        for k in L.VEG_SCALARS:
            out[rows[k], c] = float(np.asarray(v[k]).reshape(-1)[0])
        for k in L.VEG_ARRAYS:
            a = np.asarray(v[k], float)
            out[rows[k], c] = a if a.ndim == 1 else a[:, 0]
    return out


def pack_soilp(soilps, sm):
    rows = L.soil_rows(sm)
    out = np.empty((L.nsoil(sm), len(soilps)))
    for c, s in enumerate(soilps):
        for k in L.SOIL_SCALARS:
            out[rows[k], c] = float(np.asarray(s[k]).reshape(-1)[0])
        out[rows["z"], c] = np.asarray(s["z"], float)
    return out


def ensemble(base_vegp, base_soilp, ncol, m=20, sm=10, seed=1234, sigma=0.1):
    """C2/C3-style perturbed-parameter ensemble: log-normal(sigma) multipliers on PAI, gsmax, refls, lw, hgt."""
    rng = np.random.Generator(np.random.PCG64(seed))
    veg = np.repeat(pack_vegp([base_vegp], m), ncol, axis=1)
    soil = np.repeat(pack_soilp([base_soilp], sm), ncol, axis=1)
    r = L.veg_rows(m)
    f = np.exp(sigma * rng.standard_normal((5, ncol)))
    veg[r["PAI"]] *= f[0]
    veg[r["gsmax"]] *= f[1]
    veg[r["refls"]] = np.minimum(veg[r["refls"]] * f[2], 0.9)
    veg[r["lw"]] *= f[3]
    veg[r["hgt"]] *= f[4]
    veg[r["hgtg"]] = 0.05 * veg[r["hgt"]]
    return np.ascontiguousarray(veg), np.ascontiguousarray(soil)


def forcing_perturbation(ncol, seed=4321):
    """Per-column affine perturbation (scale, offset) of forcing rows 0..7 (C3: e.g. lapse-rate offsets from a DEM)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    scale = np.ones((L.NPERT, ncol)); off = np.zeros((L.NPERT, ncol))
    elev = rng.uniform(0, 600, ncol)                     # synthetic DEM (m)
    off[0] = -0.0065 * elev + 0.3 * rng.standard_normal(ncol)      # tair lapse rate
    off[2] = -0.012 * elev                                          # pressure (kPa)
    scale[1] = np.exp(0.03 * rng.standard_normal(ncol))             # relhum
    scale[3] = np.exp(0.15 * rng.standard_normal(ncol))             # wind
    scale[5] = np.exp(0.05 * rng.standard_normal(ncol))             # Rsw
    scale[7] = np.exp(0.1 * rng.standard_normal(ncol))              # theta
    return scale, off
