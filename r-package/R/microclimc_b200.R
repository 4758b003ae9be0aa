# R wrappers with the reference's signatures (ilyamaclean/microclimc: R/code.R:566, 619, 687-689, 982-985, 1125-1129;
# R/NicheMapRcode.R:717-718) over the .Call shim of src/shim.c.  They only shape inputs into the library's
# column-interleaved arrays; all arithmetic runs on the GPU (there is no CPU path).  The Python module
# microclimc_b200/api.py is the same layer for hosts without R and is what the test-suite exercises.

.VEG_SCALARS <- c("hgt","x","lw","cd","hgtg","zm0","refls","refg","refw","reflp","vegem","gsmax","q50","cpw","phw","kwood","clump")
.SOIL_SCALARS <- c("Smax","Smin","alpha","n","Ksat","Vq","Vm","Vo","Mc","rho","b","psi_e")
.STATE_SCALARS <- c("tabove","zabove","relhum","tair","tsoil","pk","H","G","psi_m")
.STATE_ARRAYS <- c("tc","tleaf","uz","z","rh","Rabs","gv","gha","L","Rswin","Rlwin")

.jday <- function(tme) {           # astronomical Julian day (microctools::jday semantics, integer day)
  y <- tme$year + 1900; mo <- tme$mon + 1; d <- tme$mday
  a <- (14 - mo) %/% 12; yy <- y + 4800 - a; mm <- mo + 12 * a - 3
  d + (153 * mm + 2) %/% 5 + 365 * yy + yy %/% 4 - yy %/% 100 + yy %/% 400 - 32045
}
.pack_vegp <- function(vl) {       # list of vegp lists -> matrix ncol x NVEG
  t(sapply(vl, function(v) c(unlist(v[.VEG_SCALARS]), as.vector(as.matrix(v$PAI)[, 1]), as.vector(as.matrix(v$pLAI)[, 1]), v$iw, v$thickw)))
}
.pack_soilp <- function(sl) t(sapply(sl, function(s) c(unlist(s[.SOIL_SCALARS]), s$z)))
.state_list <- function(st, m, sm, col = 1) {
  out <- as.list(st[col, seq_along(.STATE_SCALARS)]); names(out) <- .STATE_SCALARS
  o <- length(.STATE_SCALARS)
  for (k in .STATE_ARRAYS) { out[[k]] <- st[col, o + 1:m]; o <- o + m }
  out$gt <- st[col, o + 1:(m + 1)]; o <- o + m + 1
  out$soiltc <- st[col, o + 1:sm]; o <- o + sm
  out$sz <- st[col, o + 1:sm]
  out
}
.state_row <- function(previn) c(unlist(previn[.STATE_SCALARS]), unlist(previn[.STATE_ARRAYS]), previn$gt, previn$soiltc, previn$sz)
.cfg <- function(timestep, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, spin, char = 0)
  c(timestep, ifelse(is.na(edgedist), NaN, edgedist), ifelse(is.na(reqhgt), NaN, reqhgt), zu, merid, dst, n, as.numeric(metopen), windhgt,
    zlafact, surfwet, spin, char)
.forcing <- function(climdata, theta) {
  tme <- as.POSIXlt(climdata$obs_time, tz = "UTC")
  dp <- climdata$difrad / climdata$swrad; dp[!is.finite(dp)] <- 0.5                       # code.R:1166-1167
  rbind(climdata$temp, climdata$relhum, climdata$pres, climdata$windspeed, climdata$skyem, climdata$swrad, dp,
        rep_len(theta, length(dp)), .jday(tme), tme$hour + tme$min / 60 + tme$sec / 3600)   # MC_F_N x T
}

soilinit <- function(soiltype, m = 10, sdepth = 2, reqdepth = NA) {                       # code.R:619-627 (host side)
  sel <- which(soilparams$Soil.type == soiltype)
  soilp <- as.list(soilparams[sel, ])
  z <- (sdepth / m^1.2) * c(1:m)^1.2
  if (is.na(reqdepth) == F) z[abs(z - reqdepth) == min(abs(z - reqdepth))][1] <- reqdepth
  soilp$z <- z
  soilp
}
paraminit <- function(m, sm, hgt, tair, u, relhum, tsoil, Rsw) {                          # code.R:566-585, on the device
  ctx <- .Call(mcb_create, m, sm, 1L)
  .Call(mcb_set_columns, ctx, matrix(0, 1, 17 + 4 * m), matrix(0, 1, 12 + sm), 0, 0)
  .Call(mcb_paraminit, ctx, matrix(c(hgt, tair, u, relhum, tsoil, Rsw), 1, 6))
  .state_list(.Call(mcb_get_state, ctx, 1L, 9L + 12L * m + 1L + 2L * sm), m, sm)
}
runonestep <- function(climvars, previn, vegp, soilp, timestep, tme, lat, long, edgedist = 1000, sdepth = 2, reqhgt = NA, zu = 2,
                       theta = 0.3, thetap = 0.3, merid = 0, dst = 0, n = 0.6, metopen = TRUE, windhgt = 2, zlafact = 1, surfwet = 1) {
  m <- length(previn$tc); sm <- length(previn$soiltc)
  if (m < 3) stop("At least three canopy layers must be provided")
  if (min(vegp$PAI) == 0) warning("Some PAI values are 0. Zero PAI values set to 0.0001 to enable model to run\n")
  ctx <- .Call(mcb_create, m, sm, 1L)
  .Call(mcb_set_columns, ctx, .pack_vegp(list(vegp)), .pack_soilp(list(soilp)), lat, long)
  .Call(mcb_set_config, ctx, .cfg(timestep, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, 0))
  .Call(mcb_set_state, ctx, matrix(.state_row(previn), 1))
  cv <- matrix(c(climvars$tair, climvars$relhum, climvars$pk, climvars$u, climvars$tsoil, climvars$skyem, climvars$Rsw,
                 ifelse(is.na(climvars$dp), NaN, climvars$dp)), 1, 8)                     # `$u` partial-matches `u2` (quirk Q4)
  .Call(mcb_run_onestep, ctx, cv, .jday(tme), tme$hour + tme$min / 60 + tme$sec / 3600, theta[1], thetap[1])
  .state_list(.Call(mcb_get_state, ctx, 1L, 9L + 12L * m + 1L + 2L * sm), m, sm)
}
runmodel_batch <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, merid = 0,
                           dst = 0, n = 0.6, steps = 200, tsoil = NA, metopen = TRUE, windhgt = 2, zlafact = 1, surfwet = 1,
                           fscale = NULL, foffset = NULL, sample = NULL, n_gpus = 1L) {
  # vegp / soilp: lists of the reference's lists, one per column
  char <- is.character(reqhgt)
  if (char && !(reqhgt %in% c("all", "allclim"))) stop("input reqhgt no recognised\n")      # code.R:1136-1139
  if (metopen == F && any(windhgt < sapply(vegp, function(v) v$hgt))) stop("When metopen is FALSE, windhgt must be above canopy\n")
  m <- length(vegp[[1]]$thickw); sm <- length(soilp[[1]]$z); ncol <- length(vegp)
  ctx <- .Call(mcb_create, m, sm, as.integer(n_gpus))
  .Call(mcb_set_columns, ctx, .pack_vegp(vegp), .pack_soilp(rep_len(soilp, ncol)), rep_len(lat, ncol), rep_len(long, ncol))
  f <- .forcing(climdata, theta)
  .Call(mcb_set_forcing, ctx, f, fscale, foffset, NULL)
  tme <- as.POSIXlt(climdata$obs_time, tz = "UTC")
  timestep <- round(as.numeric(tme[2]) - as.numeric(tme[1]), 0)                            # code.R:1162
  .Call(mcb_set_config, ctx, .cfg(timestep, edgedist, if (char) NA else reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet,
                                  steps, as.numeric(char)))
  if (is.null(sample)) sample <- seq_len(ncol)
  r <- .Call(mcb_run_transient, ctx, 0, ncol(f), if (all(is.na(tsoil))) NULL else rep_len(tsoil, ncol), as.numeric(sample), ncol)
  list(series = array(r[[1]], c(length(sample), 8, ncol(f))), stats = r[[2]], flags = r[[3]],
       state = .Call(mcb_get_state, ctx, as.integer(ncol), 9L + 12L * m + 1L + 2L * sm))
}
runmodel <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, merid = 0,
                     dst = 0, n = 0.6, steps = 200, plotout = FALSE, plotsteps = 100, tsoil = NA, metopen = TRUE, windhgt = 2,
                     zlafact = 1, surfwet = 1, previn = NA, snow = NA) {
  if (!is.na(snow[1]) && snow[1] == 1) {                                                 # .snowsort, code.R:918-933 (quirk Q3)
    vegp$lw <- vegp$lw * 1.5; vegp$zm0 <- 0.024; vegp$refls <- 0.85; vegp$refg <- 0.8; vegp$refw <- 0.7; vegp$reflp <- 0.8
    vegp$gsmax <- 10; vegp$q50 <- 1
  }
  r <- runmodel_batch(climdata, list(vegp), list(soilp), lat, long, edgedist, reqhgt, sdepth, zu, theta, merid, dst, n, steps, tsoil,
                      metopen, windhgt, zlafact, surfwet)
  s <- r$series[1, , ]
  data.frame(obs_time = climdata$obs_time, reftemp = climdata$temp, tout = s[1, ], tleaf = s[2, ], relhum = s[3, ], SWin = s[4, ],
             LWin = s[5, ], H = s[7, ], L = s[6, ], G = s[8, ])                           # code.R:1324-1325
}
spinup <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, thetap = 0.3,
                   merid = 0, dst = 0, n = 0.6, plotout = TRUE, steps = 200, metopen = TRUE, windhgt = 2, zlafact = 1, surfwet = 1) {
  m <- length(vegp$thickw); sm <- length(soilp$z)
  ctx <- .Call(mcb_create, m, sm, 1L)
  .Call(mcb_set_columns, ctx, .pack_vegp(list(vegp)), .pack_soilp(list(soilp)), lat, long)
  .Call(mcb_set_forcing, ctx, .forcing(climdata, theta[1]), NULL, NULL, NULL)
  tme <- as.POSIXlt(climdata$obs_time, tz = "UTC")
  .Call(mcb_set_config, ctx, .cfg(round(as.numeric(tme[2]) - as.numeric(tme[1]), 0), edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt,
                                  zlafact, surfwet, steps))
  .Call(mcb_spinup, ctx, as.integer(steps))
  .state_list(.Call(mcb_get_state, ctx, 1L, 9L + 12L * m + 1L + 2L * sm), m, sm)
}
runmodelS_batch <- function(climdata, vegp, soilp, nmrout, reqhgt, lat, long, metopen = TRUE, windhgt = 2, surfwet = 1, groundem = 0.95,
                            n_gpus = 1L) {
  if (metopen == F && any(windhgt < sapply(vegp, function(v) v$hgt))) stop("When metopen is FALSE, windhgt must be above canopy\n")
  m <- length(vegp[[1]]$thickw); sm <- length(soilp[[1]]$z); ncol <- length(vegp)
  tme <- as.POSIXlt(climdata$obs_time, tz = "UTC")
  clim <- rbind(climdata$temp, climdata$relhum, climdata$pres, climdata$windspeed, climdata$swrad, climdata$difrad, climdata$skyem,
                .jday(tme), tme$hour + tme$min / 60 + tme$sec / 3600)
  snow <- nmrout$snow                                                                    # partial match of `snowtemp` (quirk Q17)
  sn <- if (max(nmrout$metout$SNOWDEP) > 0) t(as.matrix(snow[, 3:10])) else matrix(0, 8, ncol(clim))
  nmr <- rbind(nmrout$soiltemps$D0cm, nmrout$soilmoist$WC0cm, nmrout$metout$SNOWDEP, sn)
  ctx <- .Call(mcb_create, m, sm, as.integer(n_gpus))
  .Call(mcb_set_columns, ctx, .pack_vegp(vegp), .pack_soilp(rep_len(soilp, ncol)), rep_len(lat, ncol), rep_len(long, ncol))
  out <- .Call(mcb_run_steady, ctx, clim, nmr, c(reqhgt, as.numeric(metopen), windhgt, surfwet, groundem), NULL, ncol)
  array(out, c(ncol, 7, ncol(clim)))
}
runmodelS <- function(climdata, vegp, soilp, nmrout, reqhgt, lat, long, metopen = TRUE, windhgt = 2, surfwet = 1, groundem = 0.95) {
  o <- runmodelS_batch(climdata, list(vegp), list(soilp), nmrout, reqhgt, lat, long, metopen, windhgt, surfwet, groundem)[1, , ]
  data.frame(obs_time = climdata$obs_time, Tref = climdata$temp, Tloc = o[1, ], tleaf = o[2, ], RHref = climdata$relhum, RHloc = o[3, ],
             RSWloc = o[4, ], RLWloc = o[5, ], windspeed = o[6, ], dp = o[7, ])           # NMR.R:864-865
}
leafem <- function(tc, vegem) vegem * 5.67 * 10^-8 * (tc + 273.15)^4                      # code.R:59-62
