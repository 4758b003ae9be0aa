"""Flat, column-interleaved array layouts shared by the C-ABI (include/microclimc_b200.h) and the host API.

Every per-column structure of the reference is a 2-D float64 array
``[row][column]`` (C-contiguous, column index fastest) whose rows are listed here.
 * vegp  : 17 scalars then PAI[m], pLAI[m], iw[m], thickw[m]          (21 fields, vig.md:367-370)
 * soilp : 12 scalars then z[sm]                                        (code.R:619-627)
 * state : 9 scalars then tc, tleaf, uz, z, rh, Rabs, gv, gha, L, Rswin, Rlwin [m each], gt[m+1],
           soiltc[sm], sz[sm]                                           (previn, code.R:581-584, 898-901)
"""
VEG_SCALARS = ["hgt", "x", "lw", "cd", "hgtg", "zm0", "refls", "refg", "refw", "reflp", "vegem", "gsmax", "q50",
               "cpw", "phw", "kwood", "clump"]
VEG_ARRAYS = ["PAI", "pLAI", "iw", "thickw"]
SOIL_SCALARS = ["Smax", "Smin", "alpha", "n", "Ksat", "Vq", "Vm", "Vo", "Mc", "rho", "b", "psi_e"]
STATE_SCALARS = ["tabove", "zabove", "relhum", "tair", "tsoil", "pk", "H", "G", "psi_m"]
STATE_ARRAYS_M = ["tc", "tleaf", "uz", "z", "rh", "Rabs", "gv", "gha", "L", "Rswin", "Rlwin"]
# forcing row per time step (shared base series); rows 0..7 accept a per-column affine perturbation
FORCING = ["tair", "relhum", "pk", "u", "skyem", "Rsw", "dp", "theta", "jd", "lt"]
NFORC = len(FORCING)
NPERT = 8
# scalar run configuration (runonestep / runmodel arguments, code.R:687-689, 1125-1129)
CFG = ["timestep", "edgedist", "reqhgt", "zu", "merid", "dst", "n", "metopen", "windhgt", "zlafact", "surfwet",
       "spin_steps", "harvest_char"]
NCFG = len(CFG)
# runmodel default outputs per step (code.R:1324-1325 minus obs_time/reftemp)
SERIES = ["tout", "tleaf", "relhum", "SWin", "LWin", "L", "H", "G"]
# steady path
STEADY_CLIM = ["temp", "relhum", "pres", "windspeed", "swrad", "difrad", "skyem", "jd", "lt"]
STEADY_NMR = ["D0cm", "WC0cm", "SNOWDEP", "SN1", "SN2", "SN3", "SN4", "SN5", "SN6", "SN7", "SN8"]
STEADY_CFG = ["reqhgt", "metopen", "windhgt", "surfwet", "groundem"]
STEADY_OUT = ["Tloc", "tleaf", "RHloc", "RSWloc", "RLWloc", "windspeed", "dp"]


def nveg(m):
    return len(VEG_SCALARS) + 4 * m


def nsoil(sm):
    return len(SOIL_SCALARS) + sm


def nstate(m, sm):
    return len(STATE_SCALARS) + 11 * m + (m + 1) + 2 * sm


def veg_rows(m):
    """name -> slice of rows in the vegp array."""
    r = {k: slice(i, i + 1) for i, k in enumerate(VEG_SCALARS)}
    o = len(VEG_SCALARS)
    for k in VEG_ARRAYS:
        r[k] = slice(o, o + m); o += m
    return r


def soil_rows(sm):
    r = {k: slice(i, i + 1) for i, k in enumerate(SOIL_SCALARS)}
    r["z"] = slice(len(SOIL_SCALARS), len(SOIL_SCALARS) + sm)
    return r


def state_rows(m, sm):
    r = {k: slice(i, i + 1) for i, k in enumerate(STATE_SCALARS)}
    o = len(STATE_SCALARS)
    for k in STATE_ARRAYS_M:
        r[k] = slice(o, o + m); o += m
    r["gt"] = slice(o, o + m + 1); o += m + 1
    r["soiltc"] = slice(o, o + sm); o += sm
    r["sz"] = slice(o, o + sm); o += sm
    return r
