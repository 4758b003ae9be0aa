"""A NicheMapR-free provider of the soil / snow inputs of the steady path.

The reference obtains `nmrout` — hourly soil-surface temperature D0cm, near-surface soil moisture WC0cm, snow depth and eight snow-node
temperatures — from the NicheMapR Fortran microclimate model (runNMR, R/NicheMapRcode.R:279-300; consumed by runmodelS at NMR.R:795-796,
872, 518), which is neither part of the reference nor installable here.  `soil_snow` below produces the same tables from the same kind
of inputs runwithNMR takes (hourly climdata, daily precipitation, vegp, soilp; NMR.R:1007-1035), so that `runwithprovider` offers the
runwithNMR workflow end to end without NicheMapR.

This is a model DESIGNED HERE ([D]), not a port (there is nothing in the reference to port, hence no parity claim): an implicit
heat-conduction column on NicheMapR's ten depths under a linearised ground-surface energy balance below the canopy, a cascading bucket
for soil water, and a degree-hour snow pack with linear node temperatures.  It is host-side input preparation for the GPU hot path
(vectorised over habitat classes: arrays [T][K], K << ncol — per-column tables of a 4 x 10^6-column grid would be 3 TB)."""
import numpy as np

from . import synth as S

__all__ = ["DEP_CM", "soil_snow", "runwithprovider"]

DEP_CM = np.array([0.0, 2.5, 5.0, 10.0, 15.0, 20.0, 30.0, 50.0, 100.0, 200.0])     # NicheMapR's DEP (NMR.R:1026)
SB = 5.67e-8


def _satvap(t):
    return 0.6108 * np.exp(17.27 * t / (t + 237.3))


def _soil_thermal(theta, soilp):
    """Campbell-type conductivity (W m-1 K-1) and volumetric heat capacity (J m-3 K-1), the forms the transient path's soilk uses."""
    rho, mc = float(soilp["rho"]), float(soilp["Mc"])
    A = 0.65 - 0.78 * rho + 0.6 * rho * rho; B = 1.06 * rho; C = 1 + 2.6 / np.sqrt(mc); D = 0.03 + 0.1 * rho * rho
    lam = A + B * theta - (A - D) * np.exp(-(C * theta) ** 4)
    cv = (2.128 * float(soilp["Vq"]) + 2.385 * float(soilp["Vm"]) + 2.496 * float(soilp["Vo"]) + 4.18 * theta) * 1e6
    return lam, cv


def soil_snow(climdata, prec, vegp, soilp, lat=50.0, long=-5.0, groundem=0.95, tdeep=None, theta0=None, rainmult=1.0):
    """Hourly soil temperature and moisture at NicheMapR's ten depths, snow depth (cm) and eight snow-node temperatures.

    climdata: dict like `weather` (temp, relhum, pres, swrad, skyem, windspeed, obs_time), T hourly rows; prec: daily precipitation (mm);
    vegp: the reference's list (PAI a vector or m x T matrix); soilp: soilinit()-style dict.  Returns the reference's nmrout layout:
    dict(soiltemps={D0cm..D200cm}, soilmoist={WC0cm..WC200cm}, metout={SNOWDEP}, snowtemp={DOY, TIME, SN1..SN8})."""
    temp = np.asarray(climdata["temp"], float); T = temp.size
    rh = np.asarray(climdata["relhum"], float); pk = np.asarray(climdata["pres"], float)
    sw = np.maximum(np.asarray(climdata["swrad"], float), 0.0); skyem = np.asarray(climdata["skyem"], float)
    u = np.maximum(np.asarray(climdata["windspeed"], float), 0.1)
    P = np.asarray(prec, float) * rainmult
    PAI = np.asarray(vegp["PAI"], float)
    PAIt = PAI.sum(axis=0) if PAI.ndim == 2 else np.full(T, PAI.sum())
    refg, refls, vegem = float(vegp["refg"]), float(vegp["refls"]), float(vegp["vegem"])
    smin, smax, bb, ksat = float(soilp["Smin"]), float(soilp["Smax"]), float(soilp["b"]), float(soilp["Ksat"])
    z = DEP_CM / 100.0
    n = z.size
    dz = np.diff(z)                                               # between nodes
    thick = np.empty(n); thick[0] = dz[0] / 2; thick[1:-1] = (dz[:-1] + dz[1:]) / 2; thick[-1] = dz[-1] / 2
    tdeep = float(np.mean(temp)) if tdeep is None else float(tdeep)
    Ts = np.linspace(temp[0], tdeep, n)
    th = np.full(n, 0.5 * (smin + smax) if theta0 is None else float(theta0))
    out_t = np.empty((T, n)); out_w = np.empty((T, n)); dep = np.zeros(T); sn = np.zeros((T, 8))
    snow_we = 0.0; snow_rho = 100.0                                # water equivalent (mm), bulk density (kg m-3)
    dt = 3600.0
    for t in range(T):
        ta, ea = temp[t], _satvap(temp[t]) * rh[t] / 100.0
        tr_sw = np.exp(-np.sqrt(1 - refls) * PAIt[t]); tr_lw = np.exp(-np.sqrt(vegem) * PAIt[t])
        swg = sw[t] * tr_sw
        lwd = SB * (ta + 273.15) ** 4 * (tr_lw * skyem[t] + (1 - tr_lw) * vegem)
        ug = u[t] * np.exp(-0.6 * np.sqrt(PAIt[t]))                # wind near the ground, attenuated by the canopy
        gh = 44.6 * (pk[t] / 101.3) * (273.15 / (ta + 273.15)) * (0.004 + 0.012 * ug) * 29.3      # W m-2 K-1 (mol m-2 s-1 x cp)
        # ---- snow pack: accumulation / ageing / melt (degree-hours + absorbed short-wave) ----
        p_h = P[(t // 24) % P.size] / 24.0
        if ta < 0.5:
            snow_we += p_h; snow_rho = min(350.0, snow_rho + 0.05)
            if p_h > 0:
                snow_rho = max(100.0, snow_rho - 2.0 * p_h)
            rain = 0.0
        else:
            melt = min(snow_we, 0.15 * ta + 0.0012 * swg * (1 - 0.8))        # mm per hour
            snow_we -= melt; rain = p_h + melt
            snow_rho = min(400.0, snow_rho + 0.5) if snow_we > 0 else 100.0
        depth_m = snow_we / snow_rho                                # mm w.e. / (kg m-3) = m of snow
        dep[t] = depth_m * 100.0
        # ---- ground-surface energy balance, linearised about the previous surface temperature ----
        lam, cv = _soil_thermal(th, soilp)
        kk = 0.5 * (lam[:-1] + lam[1:]) / dz                        # W m-2 K-1 between nodes
        cap = cv * thick / dt
        rel = np.clip((th[0] - smin) / (smax - smin), 0.0, 1.0)
        t0 = Ts[0]
        es0 = _satvap(t0); des = 4098 * es0 / (t0 + 237.3) ** 2
        lam_v = (-42.575 * t0 + 44994.0)                            # J mol-1, as the reference's latent heat
        gv = rel * gh / 29.3                                        # mol m-2 s-1
        A = np.zeros((n, n)); rhs = cap * Ts
        if depth_m > 0.02:                                          # snow-covered: the surface sees the pack (conductive coupling to 0 C base)
            ksn = 2.2 * (snow_rho / 1000.0) ** 1.88 / max(depth_m, 0.02)
            tsurf_snow = min(ta, 0.0)
            A[0, 0] += ksn; rhs[0] += ksn * min(0.5 * tsurf_snow, 0.0)
        else:
            emis = groundem * SB
            b0 = 4 * emis * (t0 + 273.15) ** 3 + gh + lam_v * gv * des / pk[t]
            a0 = (1 - refg) * swg + groundem * lwd - emis * (t0 + 273.15) ** 4 - lam_v * gv * (es0 - ea) / pk[t] + gh * ta
            A[0, 0] += b0; rhs[0] += a0 + (4 * emis * (t0 + 273.15) ** 3 + lam_v * gv * des / pk[t]) * t0
        for j in range(n):
            A[j, j] += cap[j]
            if j > 0:
                A[j, j] += kk[j - 1]; A[j, j - 1] -= kk[j - 1]
            if j < n - 1:
                A[j, j] += kk[j]; A[j, j + 1] -= kk[j]
        A[n - 1, n - 1] += lam[-1] / 1.0; rhs[n - 1] += lam[-1] / 1.0 * tdeep       # deep boundary 1 m below the last node
        Ts = np.linalg.solve(A, rhs)
        # ---- soil water: infiltration at the top, evaporation from the top two nodes, Campbell drainage downwards ----
        store = th * thick * 1000.0                                 # mm
        store[0] += rain
        et = min(max(lam_v * gv * max(_satvap(Ts[0]) - ea, 0.0) / pk[t] / 2.45e6 * 3600.0, 0.0), max(store[0] - smin * thick[0] * 1000, 0.0)) \
            if depth_m <= 0.02 else 0.0
        store[0] -= et
        for j in range(n):
            cap_mm = smax * thick[j] * 1000.0
            relj = np.clip((store[j] / (thick[j] * 1000.0) - smin) / (smax - smin), 0.0, 1.0)
            dr = min(ksat / 24.0 * relj ** (2 * bb + 3), max(store[j] - smin * thick[j] * 1000.0, 0.0))
            over = max(store[j] - dr - cap_mm, 0.0)
            store[j] -= dr + over
            if j < n - 1:
                store[j + 1] += dr + over
        th = np.clip(store / (thick * 1000.0), smin, smax)
        out_t[t] = Ts; out_w[t] = th
        if depth_m > 0:
            top = min(ta, 0.0); base = min(Ts[0], 0.0)
            sn[t] = top + (base - top) * np.linspace(0.0, 1.0, 8)
    jd, lt, _ = S.parse_obs_time(climdata["obs_time"])
    names = ["D%gcm" % d for d in DEP_CM]; wn = ["WC%gcm" % d for d in DEP_CM]
    snowtemp = {"DOY": jd - jd[0] + 1, "TIME": lt * 60.0}
    snowtemp.update({"SN%d" % (q + 1): sn[:, q] for q in range(8)})
    return dict(soiltemps={k: out_t[:, j] for j, k in enumerate(names)}, soilmoist={k: out_w[:, j] for j, k in enumerate(wn)},
                metout={"SNOWDEP": dep}, snowtemp=snowtemp)


def runwithprovider(climdata, prec, vegp, soilp, reqhgt, lat, long, metopen=True, windhgt=2, surfwet=1, groundem=0.95, rainmult=1):
    """The runwithNMR workflow (R/NicheMapRcode.R:1007-1070) with `soil_snow` in NicheMapR's place: nmrout from the provider, then the
    steady-state series on the GPU (runmodelS), or the interpolated soil temperature for a negative reqhgt (NMR.R:1036-1047).
    Returns dict(metout=DataFrame, nmrout=dict)."""
    import pandas as pd
    from . import api
    hgt = float(vegp["hgt"])
    if (not metopen) and windhgt < hgt:
        raise ValueError("When metopen is FALSE, windhgt must be above canopy\\n")
    nmr = soil_snow(climdata, prec, vegp, soilp, lat, long, groundem=groundem, rainmult=rainmult)
    obs = np.asarray(climdata["obs_time"]); tair = np.asarray(climdata["temp"], float); rh = np.asarray(climdata["relhum"], float)
    if reqhgt < 0:
        soilz = -reqhgt * 100.0
        dif = np.abs(soilz - DEP_CM); o = np.argsort(dif, kind="stable")
        d1, d2 = dif[o[0]], dif[o[1]]
        names = ["D%gcm" % d for d in DEP_CM]
        tz = nmr["soiltemps"][names[o[0]]] * (d2 / (d1 + d2)) + nmr["soiltemps"][names[o[1]]] * (d1 / (d1 + d2))
        metout = pd.DataFrame(dict(obs_time=obs, Tref=tair, Tloc=tz, tleaf=-999.0, RHref=rh, RHloc=-999.0))
    elif reqhgt < hgt + 2:
        v = dict(vegp)
        T = tair.size
        for k in ("PAI", "pLAI"):                                       # runmodelS wants m x T matrices (NMR.R:730-735)
            a = np.asarray(v[k], float)
            v[k] = a if a.ndim == 2 else np.repeat(a[:, None], T, axis=1)
        metout = api.runmodelS(climdata, v, soilp, nmr, reqhgt, lat, long, metopen, windhgt, surfwet, groundem)
    else:
        import warnings
        warnings.warn("Height out of range. Output climate identical to input")
        metout = pd.DataFrame(dict(obs_time=obs, Tref=tair, Tloc=tair, tleaf=-999.0, RHref=rh, RHloc=rh))
    return dict(metout=metout, nmrout=nmr)
