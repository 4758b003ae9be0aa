# R wrappers with the reference's signatures (ilyamaclean/microclimc: R/code.R:566, 619, 687-689, 982-985, 1125-1129;
# R/NicheMapRcode.R:717-718) over the .Call shim of src/shim.c.  They only shape inputs into the library's
# column-interleaved arrays and outputs back into the reference's lists / data.frames; all arithmetic runs on the GPU
# (there is no CPU path).  microclimc_b200/api.py is the same layer for hosts without R; tests/test_r_shim.py replays the
# argument shapes built here (R column-major) through the compiled shim.
#
# Layout rule used throughout: the library's C arrays are [row][column] with the column index fastest, which is an R
# matrix of dim (columns x rows); three-index C arrays [a][b][c] are R arrays of dim c(c, b, a).

.VEG_SCALARS <- c("hgt","x","lw","cd","hgtg","zm0","refls","refg","refw","reflp","vegem","gsmax","q50","cpw","phw","kwood","clump")
.SOIL_SCALARS <- c("Smax","Smin","alpha","n","Ksat","Vq","Vm","Vo","Mc","rho","b","psi_e")
.STATE_SCALARS <- c("tabove","zabove","relhum","tair","tsoil","pk","H","G","psi_m")
.STATE_ARRAYS <- c("tc","tleaf","uz","z","rh","Rabs","gv","gha","L","Rswin","Rlwin")
.nstate <- function(m, sm) 9L + 12L * m + 1L + 2L * sm

.jday <- function(tme) {           # astronomical Julian day (microctools::jday semantics, integer day)
  y <- tme$year + 1900; mo <- tme$mon + 1; d <- tme$mday
  a <- (14 - mo) %/% 12; yy <- y + 4800 - a; mm <- mo + 12 * a - 3
  d + (153 * mm + 2) %/% 5 + 365 * yy + yy %/% 4 - yy %/% 100 + yy %/% 400 - 32045
}
.lt <- function(tme) tme$hour + tme$min / 60 + tme$sec / 3600
.first_col <- function(x) if (is.matrix(x)) as.vector(x[, 1]) else as.vector(x)
.pack_vegp <- function(vl) {       # list of vegp lists -> matrix ncol x NVEG (time-varying fields: their first time step, .vegpsort(vegp, 1))
  t(sapply(vl, function(v) { v$zm0 <- v$zm0[1]; c(unlist(v[.VEG_SCALARS]), .first_col(v$PAI), .first_col(v$pLAI), v$iw, v$thickw) }))
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
  tme <- as.POSIXlt(climdata$obs_time, format = "%Y-%m-%d %H:%M:%S", tz = "UTC")
  dp <- climdata$difrad / climdata$swrad; dp[!is.finite(dp)] <- 0.5                       # code.R:1166-1167
  rbind(climdata$temp, climdata$relhum, climdata$pres, climdata$windspeed, climdata$skyem, climdata$swrad, dp,
        rep_len(theta, length(dp)), .jday(tme), .lt(tme))                                  # MC_F_N x T
}
# .vegpsort (code.R:904-916) for every time step at once: array c(ncolv, 2m+1, T) = the library's vegt [T][2m+1][ncolv], or NULL when
# nothing varies in time.  `zm0_fixed`: .snowsort's zm0 (code.R:922) overrides the series when snow[1] == 1.
.vegt <- function(vl, nT, zm0_fixed = NA) {
  varies <- function(v) is.matrix(v$PAI) || is.matrix(v$pLAI) || length(v$zm0) > 1
  if (!any(sapply(vl, varies))) return(NULL)
  m <- length(vl[[1]]$thickw)
  out <- array(0, c(length(vl), 2 * m + 1, nT))
  for (k in seq_along(vl)) {
    v <- vl[[k]]
    P <- if (is.matrix(v$PAI)) v$PAI else matrix(v$PAI, m, nT)
    pl <- if (is.matrix(v$pLAI)) v$pLAI else matrix(v$pLAI, m, nT)
    if (ncol(P) != nT || ncol(pl) != nT || (length(v$zm0) > 1 && length(v$zm0) != nT))
      stop("time-varying vegp must have one column / element per time step of climdata\n")
    out[k, 1:m, ] <- P; out[k, m + 1:m, ] <- pl
    out[k, 2 * m + 1, ] <- if (is.na(zm0_fixed)) rep_len(v$zm0, nT) else zm0_fixed
  }
  out
}
.snowsort <- function(vegp, snow, i = 1) {                                                 # code.R:918-933
  if (class(snow)[1] != "logical" && snow[i] == 1) {
    vegp$lw <- vegp$lw * 1.5; vegp$zm0 <- 0.024; vegp$refls <- 0.85; vegp$refg <- 0.8; vegp$refw <- 0.7; vegp$reflp <- 0.8
    vegp$gsmax <- 10; vegp$q50 <- 1
  }
  vegp
}
# runmodel's theta argument (code.R:1093-1100, 1146-1154) -> list(th = forcing column, thetaz = sm x T matrix or NULL)
.theta <- function(theta, nT, sm) {
  if (is.matrix(theta)) {
    if (nrow(theta) == 1) return(list(th = as.vector(theta), thetaz = NULL))
    if (nrow(theta) != sm || ncol(theta) != nT) stop("theta matrix must have one row per soil layer and one column per time step\n")
    return(list(th = theta[1, ], thetaz = theta))
  }
  if (length(theta) == 1) return(list(th = rep(theta, nT), thetaz = NULL))
  if (length(theta) == sm) return(list(th = rep(theta[1], nT), thetaz = matrix(rep(theta, nT), ncol = nT)))
  if (length(theta) == nT) return(list(th = theta, thetaz = NULL))
  stop("theta must be a single value, a vector over soil layers or time steps, or a matrix [soil layers x time steps]\n")
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
  .state_list(.Call(mcb_get_state, ctx, 1L, .nstate(m, sm)), m, sm)
}
runonestep <- function(climvars, previn, vegp, soilp, timestep, tme, lat, long, edgedist = 1000, sdepth = 2, reqhgt = NA, zu = 2,
                       theta = 0.3, thetap = 0.3, merid = 0, dst = 0, n = 0.6, metopen = TRUE, windhgt = 2, zlafact = 1, surfwet = 1) {
  m <- length(previn$tc); sm <- length(previn$soiltc)
  if (m < 3) stop("At least three canopy layers must be provided")
  if (min(vegp$PAI) == 0) warning("Some PAI values are 0. Zero PAI values set to 0.0001 to enable model to run\n")
  if (length(theta) > 1) stop("runonestep on the GPU takes a single theta; soil moisture by depth goes through runmodel(theta = matrix)\n")
  ctx <- .Call(mcb_create, m, sm, 1L)
  .Call(mcb_set_columns, ctx, .pack_vegp(list(vegp)), .pack_soilp(list(soilp)), lat, long)
  .Call(mcb_set_config, ctx, .cfg(timestep, edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet, 0))
  .Call(mcb_set_state, ctx, matrix(.state_row(previn), 1))
  cv <- matrix(c(climvars$tair, climvars$relhum, climvars$pk, climvars$u, climvars$tsoil, climvars$skyem, climvars$Rsw,
                 ifelse(is.na(climvars$dp), NaN, climvars$dp)), 1, 8)                     # `$u` partial-matches `u2` (quirk Q4)
  tme <- as.POSIXlt(tme, tz = "UTC")
  .Call(mcb_run_onestep, ctx, cv, .jday(tme), .lt(tme), theta[1], thetap[1])
  .state_list(.Call(mcb_get_state, ctx, 1L, .nstate(m, sm)), m, sm)
}
# Batched runmodel over stacked columns.  vegp / soilp: lists of the reference's lists, one per column (soilp is recycled).
# previn: NA (spin up, code.R:1155-1159) or a list of state lists, one per column.  full = TRUE (or a character reqhgt) also returns the
# whole state of the sampled columns at every step: array c(nsample, NSTATE, T).
runmodel_batch <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, merid = 0,
                           dst = 0, n = 0.6, steps = 200, tsoil = NA, metopen = TRUE, windhgt = 2, zlafact = 1, surfwet = 1,
                           previn = NA, snow = NA, fscale = NULL, foffset = NULL, sample = NULL, full = FALSE, n_gpus = 1L) {
  char <- is.character(reqhgt)
  if (char && !(reqhgt %in% c("all", "allclim"))) stop("input reqhgt no recognised\n")      # code.R:1136-1139
  if (metopen == F && any(windhgt < sapply(vegp, function(v) v$hgt))) stop("When metopen is FALSE, windhgt must be above canopy\n")
  m <- length(vegp[[1]]$thickw); ncol <- length(vegp)
  soilp <- rep_len(soilp, ncol); sm <- length(soilp[[1]]$z)
  if (!char && !is.na(reqhgt) && reqhgt < 0) {                                             # code.R:1130-1135
    soilp <- lapply(soilp, function(s) {
      dif <- s$z + reqhgt
      sel <- which(abs(dif) == min(abs(dif)))[1]
      s$z[sel] <- -reqhgt
      if (dif[sel] < 0 & sel != length(s$z)) s$z[sel + 1] <- s$z[sel + 1] - dif[sel]
      s })
  }
  snowed <- class(snow)[1] != "logical" && snow[1] == 1                                    # quirk Q3: snow[1] decides for the whole run
  vegp <- lapply(vegp, .snowsort, snow = snow, i = 1)
  tme <- as.POSIXlt(climdata$obs_time, format = "%Y-%m-%d %H:%M:%S", tz = "UTC")
  nT <- length(tme)
  th <- .theta(theta, nT, sm)
  ctx <- .Call(mcb_create, m, sm, as.integer(n_gpus))
  .Call(mcb_set_columns, ctx, .pack_vegp(vegp), .pack_soilp(soilp), rep_len(lat, ncol), rep_len(long, ncol))
  .Call(mcb_set_forcing, ctx, .forcing(climdata, th$th), fscale, foffset, NULL)
  vt <- .vegt(vegp, nT, if (snowed) 0.024 else NA)
  if (!is.null(vt)) .Call(mcb_set_vegt, ctx, vt, nT, dim(vt)[1])
  if (!is.null(th$thetaz)) .Call(mcb_set_thetaz, ctx, th$thetaz, nT)
  timestep <- round(as.numeric(tme[2]) - as.numeric(tme[1]), 0)                            # code.R:1162
  spin <- if (class(previn)[1] == "logical") steps else 0
  .Call(mcb_set_config, ctx, .cfg(timestep, edgedist, if (char) NA else reqhgt, zu, merid, dst, n, metopen, windhgt, zlafact, surfwet,
                                  spin, as.numeric(char)))
  if (spin == 0) {
    if (!is.null(previn$tc)) previn <- list(previn)                                        # a single state list
    .Call(mcb_set_state, ctx, t(sapply(previn, .state_row)))
  }
  if (is.null(sample)) sample <- seq_len(ncol)
  want_full <- full || char
  r <- .Call(mcb_run_transient, ctx, 0, nT, if (all(is.na(tsoil))) NULL else rep_len(as.numeric(tsoil), ncol), as.numeric(sample), ncol,
             as.integer(want_full), .nstate(m, sm))
  list(series = array(r[[1]], c(length(sample), 8, nT)), stats = r[[2]], flags = r[[3]],
       full = if (want_full) array(r[[4]], c(length(sample), .nstate(m, sm), nT)) else NULL,
       state = .Call(mcb_get_state, ctx, as.integer(ncol), .nstate(m, sm)), m = m, sm = sm, sample = sample)
}
# The data.frames runmodel returns for reqhgt "allclim" / "all" (code.R:1267-1322) from the full-state rows of one column
.all_tables <- function(full, obs_time, m, sm, with_fluxes) {            # full: NSTATE x T
  rows <- function(first, n) t(full[first + seq_len(n) - 1, , drop = FALSE])
  o <- 9L
  tc <- rows(o + 1, m); tleaf <- rows(o + m + 1, m); uz <- rows(o + 2 * m + 1, m); rh <- rows(o + 4 * m + 1, m)
  gv <- rows(o + 6 * m + 1, m); gha <- rows(o + 7 * m + 1, m); L <- rows(o + 8 * m + 1, m); Rsw <- rows(o + 9 * m + 1, m); Rlw <- rows(o + 10 * m + 1, m)
  gt <- rows(o + 11 * m + 1, m + 1); soiltc <- rows(o + 12 * m + 2, sm)
  nT <- ncol(full)
  z <- full[o + 3 * m + seq_len(m), nT]; sz <- full[o + 12 * m + 1 + sm + seq_len(sm), nT]; zabove <- full[2, nT]
  df <- function(x, nms) { d <- data.frame(obs_time = obs_time, x); names(d) <- c("obs_time", nms); d }
  zn <- round(z, 3)
  out <- list(Airtemp = df(cbind(tc, full[1, ], full[9, ]), c(paste0("Temp_", zn, "m"), paste0("Temp_", round(zabove, 3), "m"), "psi_m")),
              Leaftemp = df(tleaf, paste0("Temp_", zn, "m")),
              Soiltemp = df(soiltc, paste0("Temp_", round(sz, 3), "m")),
              Windspeed = df(uz, paste0("u_", zn)),
              Relhum = df(rh, paste0("relh_", zn)))
  if (with_fluxes) {
    out$Conductivity <- df(cbind(gha, gt, gv), c(paste0("Leafbound_", zn), paste0("Air_", 0:m), paste0("Leafvap_", zn)))
    out$Fluxes <- df(cbind(Rsw, Rlw, L, full[7, ], full[8, ]),
                     c(paste0("SWin_", zn), paste0("Lwin_", zn), paste0("Latent_", zn, "W/m^2"), "Sensible", "Ground"))
  }
  out
}
runmodel <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, merid = 0,
                     dst = 0, n = 0.6, steps = 200, plotout = FALSE, plotsteps = 100, tsoil = NA, metopen = TRUE, windhgt = 2,
                     zlafact = 1, surfwet = 1, previn = NA, snow = NA) {
  r <- runmodel_batch(climdata, list(vegp), list(soilp), lat, long, edgedist, reqhgt, sdepth, zu, theta, merid, dst, n, steps, tsoil,
                      metopen, windhgt, zlafact, surfwet, previn = previn, snow = snow)
  if (is.character(reqhgt)) return(.all_tables(matrix(r$full[1, , ], nrow = dim(r$full)[2]), climdata$obs_time, r$m, r$sm, reqhgt == "all"))
  s <- matrix(r$series[1, , ], nrow = 8)
  data.frame(obs_time = climdata$obs_time, reftemp = climdata$temp, tout = s[1, ], tleaf = s[2, ], relhum = s[3, ], SWin = s[4, ],
             LWin = s[5, ], H = s[7, ], L = s[6, ], G = s[8, ])                           # code.R:1324-1325
}
spinup <- function(climdata, vegp, soilp, lat, long, edgedist = 100, reqhgt = NA, sdepth = 2, zu = 2, theta = 0.3, thetap = 0.3,
                   merid = 0, dst = 0, n = 0.6, plotout = TRUE, steps = 200, metopen = TRUE, windhgt = 2, zlafact = 1, surfwet = 1) {
  if (!isTRUE(all.equal(theta, thetap))) stop("spinup on the GPU needs thetap == theta (runmodel always calls it that way, code.R:1157-1158)\n")
  m <- length(vegp$thickw); sm <- length(soilp$z)
  ctx <- .Call(mcb_create, m, sm, 1L)
  .Call(mcb_set_columns, ctx, .pack_vegp(list(vegp)), .pack_soilp(list(soilp)), lat, long)      # .vegpsort(vegp, 1), code.R:1004
  f <- .forcing(climdata, theta[1])
  .Call(mcb_set_forcing, ctx, f, NULL, NULL, NULL)
  if (length(theta) > 1) {
    if (length(theta) != sm) stop("theta must be a single value or one value per soil layer\n")
    .Call(mcb_set_thetaz, ctx, matrix(rep(theta, ncol(f)), ncol = ncol(f)), ncol(f))
  }
  tme <- as.POSIXlt(climdata$obs_time, format = "%Y-%m-%d %H:%M:%S", tz = "UTC")
  .Call(mcb_set_config, ctx, .cfg(round(as.numeric(tme[2]) - as.numeric(tme[1]), 0), edgedist, reqhgt, zu, merid, dst, n, metopen, windhgt,
                                  zlafact, surfwet, steps))
  .Call(mcb_spinup, ctx, as.integer(steps))
  .state_list(.Call(mcb_get_state, ctx, 1L, .nstate(m, sm)), m, sm)
}
runmodelS_batch <- function(climdata, vegp, soilp, nmrout, reqhgt, lat, long, metopen = TRUE, windhgt = 2, surfwet = 1, groundem = 0.95,
                            n_gpus = 1L) {
  if (metopen == F && any(windhgt < sapply(vegp, function(v) v$hgt))) stop("When metopen is FALSE, windhgt must be above canopy\n")
  m <- length(vegp[[1]]$thickw); ncol <- length(vegp)
  soilp <- rep_len(soilp, ncol); sm <- length(soilp[[1]]$z)
  tme <- as.POSIXlt(climdata$obs_time)
  clim <- rbind(climdata$temp, climdata$relhum, climdata$pres, climdata$windspeed, climdata$swrad, climdata$difrad, climdata$skyem,
                .jday(tme), .lt(tme))
  nT <- ncol(clim)
  snow <- nmrout$snow                                                                    # partial match of `snowtemp` (quirk Q17)
  sn <- if (max(nmrout$metout$SNOWDEP) > 0) t(as.matrix(snow[, 3:10])) else matrix(0, 8, nT)
  nmr <- rbind(nmrout$soiltemps$D0cm, nmrout$soilmoist$WC0cm, nmrout$metout$SNOWDEP, sn)
  ctx <- .Call(mcb_create, m, sm, as.integer(n_gpus))
  .Call(mcb_set_columns, ctx, .pack_vegp(vegp), .pack_soilp(soilp), rep_len(lat, ncol), rep_len(long, ncol))
  vt <- .vegt(vegp, nT)                                                                  # vegp$PAI as m x T matrix (NMR.R:730-735)
  if (!is.null(vt)) .Call(mcb_set_vegt, ctx, vt, nT, dim(vt)[1])
  out <- .Call(mcb_run_steady, ctx, clim, nmr, c(reqhgt, as.numeric(metopen), windhgt, surfwet, groundem), NULL, ncol)
  list(out = array(out, c(ncol, 7, nT)), snowdep = nmrout$metout$SNOWDEP)
}
runmodelS <- function(climdata, vegp, soilp, nmrout, reqhgt, lat, long, metopen = TRUE, windhgt = 2, surfwet = 1, groundem = 0.95) {
  r <- runmodelS_batch(climdata, list(vegp), list(soilp), nmrout, reqhgt, lat, long, metopen, windhgt, surfwet, groundem)
  o <- matrix(r$out[1, , ], nrow = 7)
  tleaf <- o[2, ]
  sbs <- which(reqhgt < r$snowdep / 100)                                                 # quirk Q16 (NMR.R:590-594, 668): tleaf2 <- tz2[sbs]
  if (length(sbs) > 0) { tz2 <- o[1, sbs]; tleaf[sbs] <- tz2[sbs] }                      # double-indexes the already subset vector (NA beyond it)
  data.frame(obs_time = climdata$obs_time, Tref = climdata$temp, Tloc = o[1, ], tleaf = tleaf, RHref = climdata$relhum, RHloc = o[3, ],
             RSWloc = o[4, ], RLWloc = o[5, ], windspeed = o[6, ], dp = o[7, ])           # NMR.R:864-865
}
leafem <- function(tc, vegem) vegem * 5.67 * 10^-8 * (tc + 273.15)^4                      # code.R:59-62
