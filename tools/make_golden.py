"""Generate tests/golden/intree_known_answers.json.

The reference ships no tests and no golden vectors and R is not installed here, so the only
hard pins are the reference functions that are SELF-CONTAINED in the reference tree.  This script is a second,
independent transliteration of those functions — vectorised numpy written straight from the R text, sharing no
code with oracle/*.hpp — and evaluates them on fixed inputs.  tests/test_oracle_golden.py requires the C++ oracle
(and, on the GPU, the CUDA path) to reproduce these numbers, and also carries the values of SURVEY.md Appendix F.

Functions restated (file:line under /root/reference):
  Thomas            R/code.R:94-125          leafem      R/code.R:59-62
  paraminit         R/code.R:566-585         soilinit$z  R/code.R:619-627
  .snowtempf        R/NicheMapRcode.R:495-512   .snowsort R/code.R:918-933
  windprofile       R/code.R:191-221   (zeroplanedis / roughlength from the in-tree copies NicheMapRcode.R:103-113)
  .windprofile      R/NicheMapRcode.R:378-404
  tleafS            R/NicheMapRcode.R:432-493 (needs cpair, satvap, soilrh, dewpoint of the absent microctools: the
                    candidate definitions of DESIGN.md "compat layer" are restated here in numpy)
Run:  python tools/make_golden.py      (no GPU, no reference checkout needed; output is committed)
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "intree_known_answers.json")


# ---- R helpers ------------------------------------------------------------------------------------------------------
def r_seq_len_out(a, b, n):
    """seq(a, b, length.out = n)"""
    return a + (b - a) * (np.arange(n) / (n - 1))


def r_spline2(y1, y2, n):
    """stats::spline(c(1,2), c(y1,y2), n = n)$y — with two knots the fmm spline degenerates to the chord."""
    x = r_seq_len_out(1.0, 2.0, n)
    return y1 + (y2 - y1) * (x - 1.0)


# ---- R/code.R:94-125 ------------------------------------------------------------------------------------------------
def Thomas(tc, tmsoil, tair, k, cd, f=0.6, X=0.0):
    tc = np.array(tc, float)
    k = np.asarray(k, float)
    m = tc.size - 2
    cd = np.broadcast_to(np.asarray(cd, float), (m,))
    X = np.broadcast_to(np.asarray(X, float), (m,))
    # R indices 1..m+2 are kept by padding a dummy element 0
    TC = np.concatenate([[np.nan], tc]); K = np.concatenate([[np.nan], k])
    tn = np.zeros(m + 3); tn[m + 2] = tmsoil; tn[1] = tair
    g = 1 - f
    a = np.zeros(m + 4); b = np.zeros(m + 4); cc = np.zeros(m + 4); d = np.zeros(m + 4)
    xx = np.arange(2, m + 2)
    TC[xx] = TC[xx] + g * X
    cc[xx] = -K[xx] * f
    a[xx + 1] = cc[xx]
    b[xx] = f * (K[xx] + K[xx - 1]) + cd
    d[xx] = g * K[xx - 1] * TC[xx - 1] + (cd - g * (K[xx] + K[xx - 1])) * TC[xx] + g * K[xx] * TC[xx + 1]
    d[2] = d[2] + K[1] * tn[1] * f
    d[m + 1] = d[m + 1] + K[m + 1] * f * tn[m + 2]
    for i in range(2, m + 1):
        cc[i] = cc[i] / b[i]
        d[i] = d[i] / b[i]
        b[i + 1] = b[i + 1] - a[i + 1] * cc[i]
        d[i + 1] = d[i + 1] - a[i + 1] * d[i]
    tn[m + 1] = d[m + 1] / b[m + 1]
    for i in range(m, 1, -1):
        tn[i] = d[i] - cc[i] * tn[i + 1]
    xmn = np.minimum(np.minimum(TC[xx], TC[xx - 1]), TC[xx + 1])
    xmx = np.maximum(np.maximum(TC[xx], TC[xx - 1]), TC[xx + 1])
    tn[xx] = np.where(tn[xx] < xmn, xmn, tn[xx])
    tn[xx] = np.where(tn[xx] > xmx, xmx, tn[xx])
    tn[xx] = tn[xx] + f * X
    return tn[1:]


# ---- R/code.R:59-62 -------------------------------------------------------------------------------------------------
def leafem(tc, vegem):
    return vegem * 5.67 * 10 ** -8 * (np.asarray(tc, float) + 273.15) ** 4


# ---- R/code.R:566-585 -----------------------------------------------------------------------------------------------
def paraminit(m, sm, hgt, tair, u, relhum, tsoil, Rsw):
    tcb = r_spline2(tsoil, tair, m + sm)
    tc = tcb[sm:sm + m]
    soiltc = tcb[:sm][::-1]
    rh = np.repeat(float(relhum), m)
    Rabs = r_spline2(0.2 * Rsw + 20, Rsw + 20, m)
    tleaf = tc + 0.01 * Rabs
    gt = np.repeat(50.0, m + 1) * (m / hgt) * (2 / m)
    gt[0] = 2 * gt[0]
    gt[m] = gt[m] * (hgt / m) * 0.5
    gv = r_spline2(0.25, 0.32, m)
    gha = r_spline2(0.13, 0.19, m)
    z = (np.arange(1, m + 1) - 0.5) / m * hgt
    sz = 2 / sm ** 2.42 * np.arange(1, sm + 1) ** 2.42
    uz = (np.arange(1, m + 1) / m) * u
    return dict(tc=tc, soiltc=soiltc, tleaf=tleaf, tabove=tair, uz=uz, z=z, sz=sz, zabove=hgt, rh=rh, relhum=relhum, tair=tair,
                tsoil=tsoil, pk=101.3, Rabs=Rabs, gt=gt, gv=gv, gha=gha, H=0.0, L=np.zeros(m), G=0.0, Rswin=Rabs.copy(),
                Rlwin=0.2 * Rabs, psi_m=0.0)


# ---- R/code.R:619-627 -----------------------------------------------------------------------------------------------
def soilinit_z(m=10, sdepth=2.0, reqdepth=None):
    z = (sdepth / m ** 1.2) * np.arange(1, m + 1) ** 1.2
    if reqdepth is not None:
        dif = np.abs(z - reqdepth)
        z[np.nonzero(dif == dif.min())[0][0]] = reqdepth
    return z


# ---- R/NicheMapRcode.R:495-512 (vectorised over hours like the R) ---------------------------------------------------
def snowtempf(snowdepth, snow, reqhgt):
    """snow: array [T][10] standing for the data.frame whose columns 3..10 are SN1..SN8."""
    ndepths = np.array([0, 2.5, 5, 10, 20, 50, 100, 200, 300], float)[::-1] / 100
    ha = ndepths - reqhgt
    selb = int(np.nonzero(ha <= 0)[0][0]) + 1       # 1-based
    sela = selb - 1
    sm = abs(reqhgt - ndepths[sela - 1]) + abs(reqhgt - ndepths[selb - 1])
    wgt1 = 1 - abs(reqhgt - ndepths[sela - 1]) / sm
    wgt2 = 1 - abs(reqhgt - ndepths[selb - 1]) / sm
    snowdepth = np.asarray(snowdepth, float)
    w1 = np.repeat(wgt1, snowdepth.size); w2 = np.repeat(wgt2, snowdepth.size)
    hs = ndepths[sela - 1] > (snowdepth / 100)
    w1[hs] = 0; w2[hs] = 1
    stemp = w1 * snow[:, sela + 2 - 1] + w2 * snow[:, selb + 2 - 1]
    stemp[snowdepth < reqhgt] = 0
    return stemp


# ---- in-tree copies of two microctools functions: R/NicheMapRcode.R:103-113 ----------------------------------------
def zeroplanedis(hgt, PAI):
    PAI = np.where(np.asarray(PAI, float) < 0.001, 0.001, PAI)
    return (1 - (1 - np.exp(-np.sqrt(7.5 * PAI))) / np.sqrt(7.5 * PAI)) * hgt


def roughlength(hgt, PAI, zm0=0.003):
    Be = np.sqrt(0.003 + (0.2 * PAI) / 2)
    zm = (hgt - zeroplanedis(hgt, PAI)) * np.exp(-0.4 / Be)
    return np.where(zm < zm0, zm0, zm)          # floor at zm0: compat-layer assumption


# ---- R/code.R:191-221 -----------------------------------------------------------------------------------------------
def windprofile(ui, zi, zo, a, PAI, hgt, psi_m=0.0, hgtg=None, zm0=0.004):
    zo = np.atleast_1d(np.asarray(zo, float))
    if hgtg is None:
        hgtg = 0.05 * hgt
    hgt = hgt + zo * 0; psi_m = psi_m + zo * 0; a = a + zo * 0; hgtg = hgtg + zo * 0
    d = zeroplanedis(hgt, PAI)
    zm = roughlength(hgt, PAI, zm0)
    zmg = 0.1 * hgtg
    ln1 = np.log((zi - d) / zm) + psi_m
    uf = 0.4 * (ui / ln1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ln2 = np.log((zo - d) / zm) + psi_m
    ln2 = np.where(ln2 < 0.55, 0.55, ln2)
    uo = (uf / 0.4) * ln2
    sel = zo < hgt
    uo[sel] = (uf[sel] / 0.4) * (np.log((hgt[sel] - d[sel]) / zm[sel]) + psi_m[sel])
    sel = (zo > 0.1 * hgt) & (zo < hgt)
    uo[sel] = uo[sel] * np.exp(a[sel] * ((zo[sel] / hgt[sel]) - 1))
    sel = zo <= 0.1 * hgt
    uo[sel] = uo[sel] * np.exp(a[sel] * (((0.1 * hgt[sel]) / hgt[sel]) - 1))
    ln1 = np.log((0.1 * hgt[sel]) / zmg[sel]) + psi_m[sel]
    uf2 = 0.4 * (uo[sel] / ln1)
    ln2 = np.log(zo[sel] / zmg[sel]) + psi_m[sel]
    uo[sel] = (uf2 / 0.4) * ln2
    return uo


# ---- R/NicheMapRcode.R:378-404 --------------------------------------------------------------------------------------
def dot_windprofile(ui, zi, zo, a, hgt, PAI, psi_m=0.0, hgtg=None, zm0=0.004):
    if hgtg is None:
        hgtg = 0.05 * hgt
    d = zeroplanedis(hgt, PAI); zm = roughlength(hgt, PAI, zm0)
    ln2 = np.log((zi - d) / zm) + psi_m
    ln2 = np.where(ln2 < 0.55, 0.55, ln2)
    uf = (0.4 * ui) / ln2
    if zo >= hgt:
        ln2 = np.log((zo - d) / zm) + psi_m
        ln2 = np.where(ln2 < 0.55, 0.55, ln2)
        return (uf / 0.4) * ln2
    ln2 = np.log((hgt - d) / zm) + psi_m
    ln2 = np.where(ln2 < 0.55, 0.55, ln2)
    uh = (uf / 0.4) * ln2
    if zi >= 0.1 * hgt:
        return uh * np.exp(a * (zo / hgt) - 1)
    uo = uh * np.exp(a * ((0.1 * hgt) / hgt) - 1)
    zmg = 0.1 * hgtg
    ln1 = np.log((0.1 * hgt) / zmg) + psi_m
    uf = 0.4 * (uo / ln1)
    return (uf / 0.4) * (np.log(zo / zmg) + psi_m)


# ---- compat-layer candidates needed by tleafS (DESIGN.md "compat layer"; source absent from the reference) ---------
def cpair(tc):
    return 2e-05 * tc ** 2 + 0.0002 * tc + 29.119


def _sv_consts(tc):
    ice = tc < 0
    e0 = np.where(ice, 610.78 / 1000, 611.2 / 1000)
    L = np.where(ice, 2.834e6, 2.501e6 - 2340 * tc)
    return e0, L


def satvap(tc):
    tc = np.asarray(tc, float)
    e0, L = _sv_consts(tc)
    return e0 * np.exp((L / 461.5) * (1 / 273.15 - 1 / (tc + 273.15)))


def dewpoint(ea, tc):
    tc = np.asarray(tc, float)
    e0, L = _sv_consts(tc)
    it = 1 / 273.15 - (461.5 / L) * np.log(ea / e0)
    return 1 / it - 273.15


def soilrh(theta, b, Psie, Smax, tc):
    matric = -Psie * (theta / Smax) ** (-b)
    return np.exp((0.018 * matric) / (8.31 * (tc + 273.15)))


# ---- R/NicheMapRcode.R:432-493 --------------------------------------------------------------------------------------
def tleafS(tair, tground, relhum, pk, theta, gtt, gt0, gha, gv, gL, Rabs, vegem, soilb, Psie, Smax, surfwet, leafdens):
    tair = np.asarray(tair, float)
    cp = cpair(tair)
    aL = (gtt * (tair + 273.15) + gt0 * (tground + 273.15)) / (gtt + gt0)
    bL = (leafdens * gL) / (gtt + gt0)
    es = satvap(tair)
    eref = (relhum / 100) * es
    rhsoil = soilrh(theta, soilb, Psie, Smax, tground)
    esoil = rhsoil * satvap(tground)
    sb = 5.67 * 10 ** -8
    Rnet = Rabs - 0.97 * sb * (tair + 273.15) ** 4
    tle = tair + (0.5 * Rnet) / (cp * gha)
    wgt1 = np.where(leafdens < 1, leafdens / 2, 1 - 0.5 / leafdens)
    wgt2 = 1 - wgt1
    tapprox = wgt1 * tle + wgt2 * tair
    delta = 4098 * (0.6108 * np.exp(17.27 * tapprox / (tapprox + 237.3))) / (tapprox + 237.3) ** 2
    ae = (gtt * eref + gt0 * esoil + gv * es) / (gtt + gt0 + gv)
    be = (gv * delta) / (gtt + gt0 + gv)
    bH = gha * cp
    lam = (-42.575 * tair + 44994)
    aX = ((lam * gv) / pk) * (surfwet * es - ae)
    bX = ((lam * gv) / pk) * (surfwet * delta - be)
    aX = np.where(aX < 0, 0, aX); bX = np.where(bX < 0, 0, bX)
    aR = sb * vegem * aL ** 4
    bR = 4 * vegem * sb * (aL ** 3 * bL + (tair + 273.15) ** 3)
    dTL = (Rabs - aR - aX) / (1 + bR + bX + bH)
    tn = aL - 273.15 + bL * dTL
    tleaf = tn + dTL
    eanew = ae + be * dTL
    eanew = np.where(eanew < 0.01, 0.01, eanew)
    tmin = dewpoint(eanew, tn)
    esnew = satvap(tn)
    eanew = np.where(eanew > esnew, esnew, eanew)
    rh = (eanew / esnew) * 100
    tleaf = np.where(tleaf < tmin, tmin, tleaf)
    tn = np.where(tn < tmin, tmin, tn)
    tmax = np.where(tn + 20 < 80, tn + 20, 80)
    tleaf = np.where(tleaf > tmax, tmax, tleaf)
    tmx = np.maximum(tground + 5, tair + 5)
    tn = np.where(tn > tmx, tmx, tn)
    tmx = np.maximum(tground + 20, tair + 20)
    tleaf = np.where(tleaf > tmx, tmx, tleaf)
    tn = np.where(tn < tair - 7, tair - 7, tn)
    tleaf = np.where(tleaf < tair - 20, tair - 20, tleaf)
    return tleaf, tn, rh


def tolist(v):
    if isinstance(v, dict):
        return {k: tolist(x) for k, x in v.items()}
    if isinstance(v, np.ndarray):
        return [float(x) for x in v.reshape(-1)]
    if isinstance(v, (list, tuple)):
        return [tolist(x) for x in v]
    return None if v is None else float(v)


def main():
    G = {"_about": "independent numpy transliteration of the self-contained reference functions; see tools/make_golden.py"}
    rng = np.random.Generator(np.random.PCG64(20261019))
    # Thomas: Appendix-F cases + random diagonally dominant systems of the sizes the hot path uses (N = 10 .. 30)
    cases = [dict(tc=[10, 12, 14, 13, 11, 9], tmsoil=9.5, tair=10.5, k=[5, 4, 3, 2, 1.5], cd=[20, 25, 30, 35], f=0.6, X=[0.0]),
             dict(tc=[10, 12, 14, 13, 11, 9], tmsoil=9.5, tair=10.5, k=[5, 4, 3, 2, 1.5], cd=[20, 25, 30, 35], f=0.6,
                  X=[0.5, 0, -0.25, 0]),
             dict(tc=[10, 12, 14, 13, 11, 9], tmsoil=30, tair=-5, k=[50, 4, 3, 2, 50], cd=[0.1], f=1.0, X=[0.0])]
    for N in (3, 10, 17, 30):
        for f in (0.4, 0.6, 1.0):
            cases.append(dict(tc=list(rng.uniform(-5, 30, N + 2)), tmsoil=float(rng.uniform(5, 15)), tair=float(rng.uniform(-5, 30)),
                              k=list(rng.uniform(0.2, 40, N + 1)), cd=list(rng.uniform(1, 400, N)), f=f,
                              X=list(rng.normal(0, 0.5, N))))
    for c in cases:
        c["tn"] = Thomas(c["tc"], c["tmsoil"], c["tair"], c["k"], c["cd"] if len(c["cd"]) > 1 else c["cd"][0], c["f"],
                         c["X"] if len(c["X"]) > 1 else c["X"][0])
    G["Thomas"] = tolist(cases)
    G["leafem"] = tolist([dict(tc=t, vegem=e, out=leafem(t, e)) for t, e in [(11.0, 0.97), (-3.5, 0.95), (35.25, 1.0)]])
    pi = [(20, 10, 10, 15, 2, 80, 11, 500), (20, 10, 15.0, 2.8, 3.601108, 82.3555, 11.3383, 0.0), (3, 2, 0.5, -4.0, 0.2, 99.0, 8.0, 120.0),
          (64, 32, 32.0, 25.0, 9.0, 40.0, 12.0, 900.0)]
    G["paraminit"] = tolist([dict(args=list(a), out=paraminit(*a)) for a in pi])
    G["soilinit_z"] = tolist([dict(m=10, sdepth=2.0, reqdepth=None, z=soilinit_z()),
                              dict(m=10, sdepth=2.0, reqdepth=0.5, z=soilinit_z(10, 2.0, 0.5)),
                              dict(m=32, sdepth=3.5, reqdepth=None, z=soilinit_z(32, 3.5))])
    # .snowtempf: snowdepth in m (as passed by .runmodelsnow), SN1..8, several request heights
    T = 12
    snow = np.zeros((T, 10)); snow[:, 2:] = rng.uniform(-12, 0, (T, 8))
    sd = np.array([0, 0.01, 0.04, 0.06, 0.12, 0.3, 0.7, 1.2, 2.5, 3.5, 0.05, 0.5])
    G["snowtempf"] = tolist([dict(snowdepth=sd, SN=snow[:, 2:], reqhgt=rq, out=snowtempf(sd, snow, rq))
                             for rq in (0.03, 0.05, 0.08, 0.3, 1.0, 2.9)])
    wp = []
    for (ui, zi, a, PAI, hgt, psi) in [(3.0, 17.0, 1.9, 4.0, 15.0, 0.0), (5.5, 4.0, 2.0, 1.146, 0.12, 0.0), (2.0, 12.0, 1.2, 2.5, 10.0, -0.6),
                                      (6.0, 27.0, 2.8, 6.0, 25.0, 1.1)]:
        zo = np.array([0.02, 0.05, 0.1, 0.3, 0.9, 1.0, 1.5, 2.0]) * hgt
        wp.append(dict(ui=ui, zi=zi, zo=zo, a=a, PAI=PAI, hgt=hgt, psi_m=psi, out=windprofile(ui, zi, zo, a, PAI, hgt, psi)))
    G["windprofile"] = tolist(wp)
    wp2 = []
    for (ui, zi, zo, a, hgt, PAI, psi) in [(3.0, 17.0, 1.0, 1.9, 15.0, 4.0, 0.0), (3.0, 17.0, 16.0, 1.9, 15.0, 4.0, 0.3),
                                           (3.0, 1.0, 0.5, 1.9, 15.0, 4.0, 0.0), (4.0, 12.0, 10.0, 2.2, 10.0, 3.0, -1.0)]:
        wp2.append(dict(ui=ui, zi=zi, zo=zo, a=a, hgt=hgt, PAI=PAI, psi_m=psi, out=float(dot_windprofile(ui, zi, zo, a, hgt, PAI, psi))))
    G["dot_windprofile"] = tolist(wp2)
    n = 24
    args = dict(tair=rng.uniform(-8, 30, n), tground=rng.uniform(-5, 35, n), relhum=rng.uniform(30, 100, n), pk=rng.uniform(96, 103, n),
                theta=rng.uniform(0.1, 0.42, n), gtt=rng.uniform(0.05, 3, n), gt0=rng.uniform(0.02, 2, n), gha=rng.uniform(0.05, 0.6, n),
                gv=rng.uniform(0.0, 0.3, n), gL=rng.uniform(0.05, 0.5, n), Rabs=rng.uniform(250, 900, n), vegem=np.repeat(0.97, n),
                soilb=np.repeat(4.5, n), Psie=np.repeat(1.1, n), Smax=np.repeat(0.422, n), surfwet=rng.uniform(0.2, 1, n),
                leafdens=rng.uniform(0.05, 2.5, n))
    tl, tn, rh = tleafS(**args)
    G["tleafS"] = tolist(dict(args=args, tleaf=tl, tn=tn, rh=rh))
    G["tleafS"]["order"] = ["tair", "tground", "relhum", "pk", "theta", "gtt", "gt0", "gha", "gv", "gL", "Rabs", "vegem", "soilb", "Psie",
                            "Smax", "surfwet", "leafdens"]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(G, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
