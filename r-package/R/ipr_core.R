# Replacement compute cores for the reference package (apply as the hunks in INTEGRATION.md).
# Everything the reference does before and after these regions stays as it is.

# progress(done, total) for the .Call routines: the reference shows a pbapply timer bar around its per-date
# loop (R/IDW.R:350-352, R/Cressman.R:338-340); the engine reports finished 128-date chunks instead of dates.
.ipr_progress <- function() {
  if (!isTRUE(getOption("InterpolateR.progress", interactive()))) return(NULL)
  pbapply::pboptions(type = "timer", use_lb = FALSE, style = 1, char = "=")
  pb <- NULL
  function(done, total) {
    if (is.null(pb)) pb <<- pbapply::startpb(0, total)
    pbapply::setpb(pb, done)
    if (done >= total) pbapply::closepb(pb)
    invisible(NULL)
  }
}

.ipr_devices <- function(devices) if (is.null(devices)) NULL else as.integer(devices)

# one row per sorted unique date (first occurrence, like match() in idw_predict), one column per training
# station in the order of the distance matrix
.ipr_obs_matrix <- function(train_data, estaciones) {
  td <- train_data[order(train_data$Date), ]
  td <- td[!duplicated(td$Date), ]
  obs <- as.matrix(td[, estaciones, with = FALSE])
  storage.mode(obs) <- "double"
  obs
}

# --- replaces R/IDW.R:249-355 -------------------------------------------------------------------
# inputs available at that point of IDW(): r_zero, train_data, Points_Train (Cod-sorted, with X, Y),
# estaciones, Dates_extracted, p.  Produces `Ensamble` exactly as the reference leaves it at :355.
.ipr_idw_core <- function(r_zero, train_data, Points_Train, estaciones, Dates_extracted, p,
                          devices = getOption("InterpolateR.devices", NULL)) {
  grid_xy <- terra::as.data.frame(r_zero, xy = TRUE, cells = TRUE)[, c("x", "y", "cell")]
  grid_xy <- grid_xy[order(grid_xy$cell), ]
  obs <- .ipr_obs_matrix(train_data, estaciones)
  message("Analysis in progress. Please wait...")
  vals <- .Call(C_ipr_idw, as.double(grid_xy$x), as.double(grid_xy$y),
                as.double(Points_Train$X), as.double(Points_Train$Y), obs, as.double(p),
                .ipr_devices(devices), .ipr_progress(),
                isTRUE(getOption("InterpolateR.pinned_result", FALSE)))
  Ensamble <- terra::rast(r_zero, nlyrs = length(Dates_extracted))
  terra::values(Ensamble) <- vals          # n_cells x n_dates, one layer per date
  Ensamble
}

# --- replaces R/Cressman.R:268-345 --------------------------------------------------------------
# inputs: spl_layer, train_data, Points_Train (ordered by ID), search_radius_m (sorted unique),
# Dates_extracted.  Produces the unnamed list `Ensamble` of per-radius stacks as at :343-345.
.ipr_cressman_core <- function(spl_layer, train_data, Points_Train, search_radius_m, Dates_extracted,
                               devices = getOption("InterpolateR.devices", NULL)) {
  if (terra::is.lonlat(spl_layer)) {
    stop("shapefile must use a projected CRS (meters). Reproject before calling Cressman.")
  }
  grid_xy <- terra::as.data.frame(spl_layer, xy = TRUE, cells = TRUE, na.rm = FALSE)[, c("x", "y", "cell")]
  grid_xy <- grid_xy[order(grid_xy$cell), ]
  obs <- .ipr_obs_matrix(train_data, Points_Train$Cod)
  message("Analysis in progress. Please wait...")
  arr <- .Call(C_ipr_cressman, as.double(grid_xy$x), as.double(grid_xy$y),
               as.double(Points_Train$X), as.double(Points_Train$Y), obs, as.double(search_radius_m),
               .ipr_devices(devices), .ipr_progress(),
               isTRUE(getOption("InterpolateR.pinned_result", FALSE)))
  lapply(seq_along(search_radius_m), function(i) {
    r <- terra::rast(spl_layer, nlyrs = length(Dates_extracted))
    terra::values(r) <- arr[, , i]
    r
  })
}

# --- SURVEY 8(f1): validation without reading the stack back --------------------------------------
# What terra::extract(Ensamble, Points_VectTest) returns inside .validate()
# (R/Intr_Connections_module.R:220-223) — a data.frame(ID, <one column per layer>) — computed by evaluating the
# engine on the cells that hold the test stations only (a few hundred cells instead of a 29 GB stack).
# `template` is the one-layer raster the stack is built on (r_zero / spl_layer), `core` one of the two
# closures below, `n_round` what the caller applies to the stack (IDW: terra::app(round), Cressman: round()).
# A point outside the raster gets a row of NA, as with terra::extract.
.ipr_extract_cells <- function(template, test_cords, layer_names, core, n_round = NULL) {
  cells <- terra::cellFromXY(template, as.matrix(test_cords[, c("X", "Y")]))
  inside <- !is.na(cells)
  ucells <- unique(cells[inside])
  out <- matrix(NA_real_, nrow = length(cells), ncol = length(layer_names))
  if (length(ucells) > 0) {
    xy <- terra::xyFromCell(template, ucells)
    vals <- core(as.double(xy[, 1]), as.double(xy[, 2]))      # n_ucells x n_dates
    if (!is.null(n_round)) vals <- round(vals, n_round)
    out[inside, ] <- vals[match(cells[inside], ucells), , drop = FALSE]
  }
  colnames(out) <- layer_names
  data.frame(ID = seq_along(cells), out, check.names = FALSE)
}

.ipr_idw_cells_core <- function(train_data, Points_Train, estaciones, p,
                                devices = getOption("InterpolateR.devices", NULL)) {
  obs <- .ipr_obs_matrix(train_data, estaciones)
  function(x, y) {
    .Call(C_ipr_idw, x, y, as.double(Points_Train$X), as.double(Points_Train$Y), obs, as.double(p),
          .ipr_devices(devices), NULL, FALSE)
  }
}

# one radius of Cressman (the list returned by Cressman() is validated radius by radius, R/Cressman.R:361-372)
.ipr_cressman_cells_core <- function(train_data, Points_Train, radius_m,
                                     devices = getOption("InterpolateR.devices", NULL)) {
  obs <- .ipr_obs_matrix(train_data, Points_Train$Cod)
  function(x, y) {
    arr <- .Call(C_ipr_cressman, x, y, as.double(Points_Train$X), as.double(Points_Train$Y), obs,
                 as.double(radius_m), .ipr_devices(devices), NULL, FALSE)
    matrix(arr, nrow = length(x))
  }
}

# --- SURVEY 8(f3), first slice: the IDW fallback of Kriging_Ordinary -------------------------------
# perform_kriging() decides per date whether to krige or to fall back to IDW (flat variogram, singular or
# ill-conditioned gamma matrix, R/Kriging_Ordinary.R:428-462).  The dates that fall back are collected and
# evaluated in one call; `obs` holds their rows (n_dates x n_stations, NA = station not available that day).
# Returns the n_cells x n_dates matrix of layer values (data_Kriging order), including the constant layers of
# the prelude (:406-423).
.ipr_kriging_idw_fallback <- function(data_Kriging, Points_Train, obs,
                                      devices = getOption("InterpolateR.devices", NULL)) {
  storage.mode(obs) <- "double"
  .Call(C_ipr_idw_kriging, as.double(data_Kriging$X), as.double(data_Kriging$Y),
        as.double(Points_Train$X), as.double(Points_Train$Y), obs, .ipr_devices(devices), NULL, FALSE)
}

# --- SURVEY 8(f3): the kriging branch ---------------------------------------------------------------
# Replaces the prediction step of perform_kriging() for ALL dates at once.  process_day() keeps what it does
# before it (R/Kriging_Ordinary.R:510-540: the empirical variogram and fit_variogram per date) and the decision of
# :428-462 (flat variogram, det == 0, kappa > 1e12 -> IDW); it then only RECORDS the outcome per date:
#   params: data.frame(model = <0 fallback | 1 exponential | 2 spherical | 3 gaussian | 4 linear>, nugget, sill, range)
# and the layers of all dates come from one call.  `obs`: n_dates x n_stations (NA = not available that day), columns
# aligned with Points_Train.  Returns the n_cells x n_dates matrix of layer values in data_Kriging order: the
# constant layers of the prelude (:406-423), the IDW fallback (:465-476) and max(0, sum(weights * values))
# (:478-497; mean(values) where the system is exactly singular, :494).
.ipr_kriging_core <- function(data_Kriging, Points_Train, obs, params,
                              devices = getOption("InterpolateR.devices", NULL)) {
  storage.mode(obs) <- "double"
  .Call(C_ipr_kriging, as.double(data_Kriging$X), as.double(data_Kriging$Y),
        as.double(Points_Train$X), as.double(Points_Train$Y), obs, as.integer(params$model),
        as.double(params$nugget), as.double(params$sill), as.double(params$range), .ipr_devices(devices))
}

# returns the device memory the engine keeps between calls (call it after a large job if other GPU work follows)
ipr_release <- function() invisible(.Call(C_ipr_release))
