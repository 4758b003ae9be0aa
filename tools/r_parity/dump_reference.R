# Real-R parity harness, producer side (SURVEY.md §8f row 1).
#
# Run this where R, microctools and microclimc are installed (they are NOT in the build image, so this script has never been
# executed there — it is written against the reference sources, code.R:566, 619, 687-689, 982-1013, 1125-1328 and
# NMR.R:717-885):
#     Rscript tools/r_parity/dump_reference.R out_dir
# It runs the reference on BASELINE config 1 (single column, bundled `weather`, Loam, the synthetic deciduous-broadleaf vegp
# of microclimc_b200/synth.py::vegp_deciduous) and writes plain CSV fixtures that tests/test_r_parity.py compares with the
# in-repo oracle (and through it with the CUDA path):
#     vegp.csv, soilp.csv             the parameter lists (one row per field; vectors comma-separated, 17 significant digits)
#     step_in.csv / step_out.csv      previn before / after ONE runonestep from paraminit on the first weather row
#     step60_in.csv / step60_out.csv  the same with timestep = 60 s from a spun-up state: the merged-layer branch (code.R:813-852)
#     spin_out.csv                    previn after spinup(steps = 50)
#     runmodel_out.csv                data.frame returned by runmodel on the first 72 hours (reqhgt = NA, steps = 50)
#     runmodel_h1_out.csv             the same with reqhgt = 1 (node snapped to the output height, code.R:719)
#     nmrout.csv, runmodelS_out.csv   a synthetic nmrout (soil surface temperature / moisture / snow pack) and runmodelS on 240 hours
#     mt_<function>.csv               inputs and results of the first calls of every microctools function the path uses (traced inside the
#                                     runs above): "<call>,<arg or =field>,<values...>" — what re-bases SURVEY.md Appendix A one by one
# Copy out_dir to tests/golden/r_reference/ and the parity statements of this repository stop being "vs in-repo oracle" only.
suppressMessages({library(microctools); library(microclimc)})
args <- commandArgs(trailingOnly = TRUE)
out <- if (length(args) >= 1) args[1] else "r_reference"
dir.create(out, showWarnings = FALSE, recursive = TRUE)
fmt <- function(x) { x <- suppressWarnings(as.numeric(unlist(x))); ifelse(is.na(x), "NA", format(x, digits = 17)) }
wlist <- function(l, file) {
  con <- file(file, "w")
  for (k in names(l)) writeLines(paste0(k, ",", paste(fmt(l[[k]]), collapse = ",")), con)
  close(con)
}

# ---- trace every microctools function the hot path calls (SURVEY.md Appendix A) --------------------------------------------------------
MT <- c("phair", "cpair", "satvap", "dewpoint", "zeroplanedis", "roughlength", "mixinglength", "attencoef", "diabatic_cor",
        "diabatic_cor_can", "gturb", "gcanopy", "gforcedfree", "layercond", "soilrh", "soilk", "layermerge", "layeraverage", "layerinterp",
        "jday", "solalt", "difprop", "psunlit", "radmult", "cansw", "cantransdif", "canlw")
KEEP <- 40                                             # calls recorded per function
.log <- new.env()
tracer <- function(fname, orig) {
  function(...) {
    res <- orig(...)
    n <- if (is.null(.log[[fname]])) 0 else length(.log[[fname]])
    if (n < KEEP) {
      a <- as.list(match.call(orig, sys.call()))[-1]
      a <- lapply(a, function(e) tryCatch(eval(e, parent.frame(2)), error = function(err) NA))
      .log[[fname]] <- c(.log[[fname]], list(list(args = a, res = res)))
    }
    res
  }
}
patch <- function(env, fname, f) if (exists(fname, envir = env, inherits = FALSE)) {
  locked <- bindingIsLocked(fname, env); if (locked) unlockBinding(fname, env)
  assign(fname, f, envir = env); if (locked) lockBinding(fname, env)
}
for (fn in MT) if (exists(fn, envir = asNamespace("microctools"), inherits = FALSE)) {
  w <- tracer(fn, get(fn, envir = asNamespace("microctools")))
  patch(asNamespace("microctools"), fn, w)
  patch(parent.env(asNamespace("microclimc")), fn, w)                     # microclimc's imports environment
  if (exists(fn, envir = as.environment("package:microctools"), inherits = FALSE)) patch(as.environment("package:microctools"), fn, w)
}
dump_traces <- function() for (fn in ls(.log)) {
  con <- file(file.path(out, paste0("mt_", fn, ".csv")), "w")
  k <- 0
  for (rec in .log[[fn]]) {
    k <- k + 1
    for (an in names(rec$args)) {
      v <- rec$args[[an]]
      if (inherits(v, "POSIXt")) { lt <- as.POSIXlt(v); v <- c(lt$year + 1900, lt$mon + 1, lt$mday, lt$hour + lt$min / 60 + lt$sec / 3600) }
      if (is.list(v)) for (f2 in names(v)) writeLines(paste0(k, ",", an, "$", f2, ",", paste(fmt(v[[f2]]), collapse = ",")), con)
      else writeLines(paste0(k, ",", an, ",", paste(fmt(v), collapse = ",")), con)
    }
    r <- rec$res
    if (is.list(r)) for (f2 in names(r)) writeLines(paste0(k, ",=", f2, ",", paste(fmt(r[[f2]]), collapse = ",")), con)
    else writeLines(paste0(k, ",=,", paste(fmt(r), collapse = ",")), con)
  }
  close(con)
}

# ---- BASELINE config 1 inputs ------------------------------------------------------------------------------------------------------
m <- 20; sm <- 10; hgt <- 15; PAIt <- 4
zz <- ((1:m) - 0.5) / m
bell <- exp(-0.5 * ((zz - 0.7) / 0.18)^2) + 0.15 * exp(-0.5 * ((zz - 0.1) / 0.08)^2) + 0.05
vegp <- list(hgt = hgt, PAI = PAIt * bell / sum(bell), x = 1.2, lw = 0.05, cd = 0.2, iw = seq(0.8, 0.3, length.out = m),
             hgtg = 0.05 * hgt, zm0 = 0.004, pLAI = rep(0.8, m), refls = 0.4, refg = 0.15, refw = 0.15, reflp = 0.4,
             vegem = 0.97, gsmax = 0.33, q50 = 100, thickw = seq(0.4, 0.01, length.out = m) * 0.1 + 0.002, cpw = 2500,
             phw = 700, kwood = 0.3, clump = 0)
soilp <- soilinit("Loam", m = sm)
wlist(vegp, file.path(out, "vegp.csv")); wlist(soilp, file.path(out, "soilp.csv"))
data(weather, package = "microclimc")
w <- weather
tme <- as.POSIXlt(w$obs_time, format = "%Y-%m-%d %H:%M:%S", tz = "UTC")
lat <- 50; long <- -5
tsoil <- mean(w$temp, na.rm = TRUE)
# the climvars list exactly as spinup / runmodel build it (code.R:999-1001, 1194-1196); runonestep reads $u through partial matching of u2
climvars_at <- function(i) {
  dp <- w$difrad[i] / w$swrad[i]; dp[!is.finite(dp)] <- 0.5
  list(tair = w$temp[i], relhum = w$relhum[i], pk = w$pres[i], u2 = w$windspeed[i], tsoil = tsoil, skyem = w$skyem[i], Rsw = w$swrad[i], dp = dp)
}

# ---- one step from paraminit ---------------------------------------------------------------------------------------------------------
previn <- paraminit(m, sm, hgt, w$temp[1], w$windspeed[1], w$relhum[1], tsoil, w$swrad[1])
wlist(previn, file.path(out, "step_in.csv"))
wlist(climvars_at(1), file.path(out, "step_climvars.csv"))
one <- runonestep(climvars_at(1), previn, vegp, soilp, timestep = 3600, tme = tme[1], lat = lat, long = long, edgedist = 100, reqhgt = NA,
                  theta = 0.3, thetap = 0.3)
wlist(one, file.path(out, "step_out.csv"))

# ---- spin-up, then one 60 s step at noon of day 180 from the spun-up state (merged air layers) --------------------------------------------
sp <- spinup(w, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, theta = 0.3, thetap = 0.3, plotout = FALSE, steps = 50)
wlist(sp, file.path(out, "spin_out.csv"))
i60 <- 180 * 24 + 13
wlist(sp, file.path(out, "step60_in.csv"))
wlist(c(climvars_at(i60), list(jd = microctools::jday(tme = tme[i60]), lt = tme[i60]$hour)), file.path(out, "step60_climvars.csv"))
s60 <- runonestep(climvars_at(i60), sp, vegp, soilp, timestep = 60, tme = tme[i60], lat = lat, long = long, edgedist = 100, reqhgt = NA,
                  theta = 0.3, thetap = 0.3)
wlist(s60, file.path(out, "step60_out.csv"))

# ---- runmodel --------------------------------------------------------------------------------------------------------------------------
rm1 <- runmodel(w[1:72, ], vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, theta = 0.3, steps = 50, plotout = FALSE)
write.csv(rm1, file.path(out, "runmodel_out.csv"), row.names = FALSE)
rm2 <- runmodel(w[1:72, ], vegp, soilp, lat, long, edgedist = 100, reqhgt = 1, theta = 0.3, steps = 50, plotout = FALSE)
write.csv(rm2, file.path(out, "runmodel_h1_out.csv"), row.names = FALSE)

# ---- runmodelS on a synthetic nmrout (the NicheMapR outputs it reads: NMR.R:795-796, 872, 518) ------------------------------------------
nS <- 240; sel <- 300 + 1:nS                                             # midwinter: some hours below zero
ws <- w[sel, ]
snowdep <- ifelse(ws$temp < 3, 12 + 8 * sin((1:nS) / 17), 0)             # cm
sn <- sapply(1:8, function(q) pmin(0, ws$temp - 1 - 0.3 * q))
nmrout <- list(soiltemps = data.frame(D0cm = ws$temp + 0.02 * ws$swrad * exp(-PAIt)),
               soilmoist = data.frame(WC0cm = rep(0.3, nS)),
               metout = data.frame(SNOWDEP = snowdep),
               snowtemp = data.frame(DOY = 1, TIME = 1, sn))
names(nmrout$snowtemp) <- c("DOY", "TIME", paste0("SN", 1:8))
write.csv(data.frame(D0cm = nmrout$soiltemps$D0cm, WC0cm = nmrout$soilmoist$WC0cm, SNOWDEP = snowdep, sn), file.path(out, "nmrout.csv"),
          row.names = FALSE)
vegpS <- vegp; vegpS$PAI <- matrix(rep(vegp$PAI, nS), ncol = nS); vegpS$pLAI <- matrix(rep(vegp$pLAI, nS), ncol = nS)   # runmodelS needs matrices (NMR.R:730-735)
rs <- runmodelS(ws, vegpS, soilp, nmrout, reqhgt = 1, lat, long)
write.csv(cbind(rs, row = sel), file.path(out, "runmodelS_out.csv"), row.names = FALSE)

dump_traces()
cat("wrote", out, "\n")
