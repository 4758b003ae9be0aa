# Real-R parity harness, producer side (SURVEY.md §8f row 1).
#
# Run this where R, microctools and microclimc are installed (they are NOT in the build image, so this script has never been
# executed there — it is written against the reference signatures, code.R:566, 619, 687-689, 982-985, 1125-1129 and
# NMR.R:717-718):
#     Rscript tools/r_parity/dump_reference.R out_dir
# It runs the reference on BASELINE config 1 (single column, bundled `weather`, Loam, the synthetic deciduous-broadleaf vegp
# of microclimc_b200/synth.py::vegp_deciduous) and writes plain CSV fixtures that tests/test_r_parity.py compares with the
# CUDA path and with the in-repo oracle:
#     vegp.csv, soilp.csv          the parameter lists (one row per field; vectors comma-separated)
#     step_in.csv / step_out.csv   previn before / after ONE runonestep from paraminit on the first weather row
#     spin_out.csv                 previn after spinup(steps = 50)
#     runmodel_out.csv             data.frame returned by runmodel on the first 72 hours (reqhgt = NA, steps = 50)
# Copy out_dir to tests/golden/r_reference/ and the parity statements of this repository stop being "vs in-repo oracle" only.
suppressMessages({library(microctools); library(microclimc)})
args <- commandArgs(trailingOnly = TRUE)
out <- if (length(args) >= 1) args[1] else "r_reference"
dir.create(out, showWarnings = FALSE, recursive = TRUE)
m <- 20; sm <- 10; hgt <- 15; PAIt <- 4
zz <- ((1:m) - 0.5) / m
bell <- exp(-0.5 * ((zz - 0.7) / 0.18)^2) + 0.15 * exp(-0.5 * ((zz - 0.1) / 0.08)^2) + 0.05
vegp <- list(hgt = hgt, PAI = PAIt * bell / sum(bell), x = 1.2, lw = 0.05, cd = 0.2, iw = seq(0.8, 0.3, length.out = m),
             hgtg = 0.05 * hgt, zm0 = 0.004, pLAI = rep(0.8, m), refls = 0.4, refg = 0.15, refw = 0.15, reflp = 0.4,
             vegem = 0.97, gsmax = 0.33, q50 = 100, thickw = seq(0.4, 0.01, length.out = m) * 0.1 + 0.002, cpw = 2500,
             phw = 700, kwood = 0.3, clump = 0)
soilp <- soilinit("Loam", m = sm)
wlist <- function(l, file) {
  con <- file(file, "w")
  for (k in names(l)) writeLines(paste0(k, ",", paste(format(as.numeric(unlist(l[[k]])), digits = 17), collapse = ",")), con)
  close(con)
}
wlist(vegp, file.path(out, "vegp.csv")); wlist(soilp, file.path(out, "soilp.csv"))
data(weather, package = "microclimc")
w <- weather
tme <- as.POSIXlt(w$obs_time, format = "%Y-%m-%d %H:%M:%S", tz = "UTC")
lat <- 50; long <- -5
cv <- w[1, ]
previn <- paraminit(m, sm, hgt, cv$temp, cv$windspeed, cv$relhum, mean(w$temp), cv$swrad)
wlist(previn, file.path(out, "step_in.csv"))
one <- runonestep(cv, previn, vegp, soilp, timestep = 3600, tme = tme[1], lat = lat, long = long, edgedist = 100, reqhgt = NA,
                  theta = 0.3, thetap = 0.3)
wlist(one, file.path(out, "step_out.csv"))
sp <- spinup(w, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, theta = 0.3, thetap = 0.3, plotout = FALSE, steps = 50)
wlist(sp, file.path(out, "spin_out.csv"))
rm <- runmodel(w[1:72, ], vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, theta = 0.3, steps = 50, plotout = FALSE)
write.csv(rm, file.path(out, "runmodel_out.csv"), row.names = FALSE)
cat("wrote", out, "\n")
