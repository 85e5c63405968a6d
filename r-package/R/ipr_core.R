# Replacement compute cores for the reference package (apply as the two hunks in INTEGRATION.md).
# Everything the reference does before and after these regions stays as it is.

# --- replaces R/IDW.R:249-355 -------------------------------------------------------------------
# inputs available at that point of IDW(): r_zero, train_data, Points_Train (Cod-sorted, with X, Y),
# estaciones, Dates_extracted, p.  Produces `Ensamble` exactly as the reference leaves it at :355.
.ipr_idw_core <- function(r_zero, train_data, Points_Train, estaciones, Dates_extracted, p,
                          devices = getOption("InterpolateR.devices", NULL)) {
  grid_xy <- terra::as.data.frame(r_zero, xy = TRUE, cells = TRUE)[, c("x", "y", "cell")]
  grid_xy <- grid_xy[order(grid_xy$cell), ]
  # one row per sorted unique date (first occurrence, like match() in idw_predict), one column per
  # training station in the order of the distance matrix
  td <- train_data[order(train_data$Date), ]
  td <- td[!duplicated(td$Date), ]
  obs <- as.matrix(td[, estaciones, with = FALSE])
  storage.mode(obs) <- "double"
  vals <- .Call(C_ipr_idw, as.double(grid_xy$x), as.double(grid_xy$y),
                as.double(Points_Train$X), as.double(Points_Train$Y), obs, as.double(p),
                if (is.null(devices)) NULL else as.integer(devices))
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
  estaciones <- Points_Train$Cod
  td <- train_data[order(train_data$Date), ]
  td <- td[!duplicated(td$Date), ]
  obs <- as.matrix(td[, estaciones, with = FALSE])
  storage.mode(obs) <- "double"
  arr <- .Call(C_ipr_cressman, as.double(grid_xy$x), as.double(grid_xy$y),
               as.double(Points_Train$X), as.double(Points_Train$Y), obs, as.double(search_radius_m),
               if (is.null(devices)) NULL else as.integer(devices))
  lapply(seq_along(search_radius_m), function(i) {
    r <- terra::rast(spl_layer, nlyrs = length(Dates_extracted))
    terra::values(r) <- arr[, , i]
    r
  })
}
