#!/usr/bin/env Rscript
# Turnkey golden generator from the REAL package (VERDICT r1 item 5a).
#
#   Rscript tests/golden/make_golden.R [outdir = tests/golden/r]
#
# Needs R with InterpolateR (the unmodified reference: R CMD INSTALL /root/reference) and terra.  Neither exists
# in the build container nor on the GPU box, so this script has never been run there; the day a box has them,
# one command freezes the reference's own outputs and `pytest tests/test_r_golden.py` pins the oracle (and, on a
# GPU, the CUDA path) to them.  Until then DESIGN.md and the oracle header keep saying "parity unpinned".
#
# Every array is written as raw little-endian doubles (exact; `<name>.f64`), described by manifest.csv
# (case, name, nrow, ncol); tables (gof / categorical metrics) additionally as CSV with 17 significant digits.
#
# Cases (call shapes of tests/testthat/test-IDW.R:17-24 and test-Cressman.R:19-31):
#   pk_idw_m001      IDW(BD_Obs, BD_Coord, shp, grid_resolution = 5, p = 2, n_round = NULL, stat_validation = "M001")
#   pk_idw_m004_r1   IDW(..., n_round = 1, stat_validation = "M004", Rain_threshold = <5 classes>)   (test-IDW.R:17)
#   pk_crs_m001      Cressman(..., grid_resolution = 5, search_radius = c(50, 20, 10), stat_validation = "M001",
#                             Rain_threshold = <5 classes>)
#   syn_a            a seeded synthetic network built below (30 stations, 24 dates, 10 % NA, a station on a cell
#                    centre, an all-NA and an all-zero date): IDW p = 2, IDW p = 3, Cressman radii 5/10/20/50 km,
#                    stat_validation = two stations
suppressPackageStartupMessages({
  library(InterpolateR)
  library(terra)
  library(data.table)
})

args <- commandArgs(trailingOnly = TRUE)
outdir <- if (length(args) >= 1) args[1] else file.path("tests", "golden", "r")
dir.create(outdir, recursive = TRUE, showWarnings = FALSE)
manifest <- data.frame(case = character(), name = character(), nrow = integer(), ncol = integer())

put <- function(case, name, m) {
  m <- as.matrix(m)
  storage.mode(m) <- "double"
  con <- file(file.path(outdir, sprintf("%s.%s.f64", case, name)), "wb")
  writeBin(as.vector(m), con, size = 8, endian = "little")   # column-major
  close(con)
  manifest[nrow(manifest) + 1, ] <<- list(case, name, nrow(m), ncol(m))
}
put_table <- function(case, name, tab) {
  tab <- as.data.frame(tab)
  num <- vapply(tab, is.numeric, logical(1))
  tab[num] <- lapply(tab[num], function(v) formatC(v, digits = 17, format = "g"))
  write.csv(tab, file.path(outdir, sprintf("%s.%s.csv", case, name)), row.names = FALSE, quote = TRUE)
}
put_text <- function(case, name, x) writeLines(as.character(x), file.path(outdir, sprintf("%s.%s.txt", case, name)))

put_inputs <- function(case, BD_Obs, BD_Coord, shp) {
  st <- setdiff(names(BD_Obs), "Date")
  put(case, "obs", as.matrix(as.data.frame(BD_Obs)[, st]))            # n_dates x n_stations, BD_Obs column order
  put_text(case, "stations", st)
  put_text(case, "dates", as.character(BD_Obs$Date))
  crd <- as.data.frame(BD_Coord)
  put_text(case, "coord_cod", crd$Cod)
  put(case, "coord_xy", cbind(crd$X, crd$Y))
  e <- terra::ext(shp)
  put(case, "shp_ext", matrix(c(e$xmin, e$xmax, e$ymin, e$ymax), nrow = 1))
}
put_stack <- function(case, name, r) {
  put(case, paste0(name, ".values"), terra::values(r))                 # n_cells x n_layers, terra cell order
  put(case, paste0(name, ".xy"), terra::xyFromCell(r, seq_len(terra::ncell(r))))
  e <- terra::ext(r)
  put(case, paste0(name, ".geom"), matrix(c(terra::ncol(r), terra::nrow(r), e$xmin, e$xmax, e$ymin, e$ymax), nrow = 1))
  put_text(case, paste0(name, ".names"), names(r))
}
put_validation <- function(case, name, v) {
  if (is.data.frame(v)) {
    put_table(case, paste0(name, ".gof"), v)
  } else {
    put_table(case, paste0(name, ".gof"), v$gof)
    put_table(case, paste0(name, ".categorical"), v$categorical_metrics)
  }
}

Rain_threshold <- list(no_rain = c(0, 1), light_rain = c(1, 5), moderate_rain = c(5, 20),
                       heavy_rain = c(20, 40), extremely_rain = c(40, Inf))

# ---------------------------------------------------------------- packaged data (BASELINE configs[0], [1])
data("BD_Obs", package = "InterpolateR")
data("BD_Coord", package = "InterpolateR")
shp <- terra::vect(system.file("extdata", "study_area.shp", package = "InterpolateR"))

put_inputs("pk", BD_Obs, BD_Coord, shp)
o <- IDW(BD_Obs, BD_Coord, shp, grid_resolution = 5, p = 2, n_round = NULL, training = 1,
         stat_validation = "M001", Rain_threshold = NULL, save_model = FALSE)
put_stack("pk_idw_m001", "ensamble", o$Ensamble)
put_validation("pk_idw_m001", "validation", o$Validation)

o <- IDW(BD_Obs, BD_Coord, shp, grid_resolution = 5, p = 2, n_round = 1, training = 1,
         Rain_threshold = Rain_threshold, stat_validation = "M004", save_model = FALSE)
put_stack("pk_idw_m004_r1", "ensamble", o$Ensamble)
put_validation("pk_idw_m004_r1", "validation", o$Validation)

o <- Cressman(BD_Obs, BD_Coord, shp, grid_resolution = 5, search_radius = c(50, 20, 10), training = 1,
              stat_validation = "M001", Rain_threshold = Rain_threshold, save_model = FALSE)
put_text("pk_crs_m001", "radius_names", names(o$Ensamble))
for (i in seq_along(o$Ensamble)) {
  put_stack("pk_crs_m001", sprintf("ensamble%d", i), o$Ensamble[[i]])
  put_validation("pk_crs_m001", sprintf("validation%d", i), o$Validation[[i]])
}

# ---------------------------------------------------------------- a small synthetic network
set.seed(42)
n_st <- 30; n_d <- 24
xmin <- 200000; ymin <- 9000000; W <- 40000; H <- 30000
X <- xmin + runif(n_st) * W
Y <- ymin + runif(n_st) * H
X[1] <- xmin + 7.5 * 1000; Y[1] <- ymin + H - 3.5 * 1000          # exactly a cell centre of the 1 km grid
cod <- sprintf("S%03d", seq_len(n_st))
obs <- matrix(ifelse(runif(n_d * n_st) < 0.45, round(rgamma(n_d * n_st, 0.8, scale = 8), 1), 0), n_d, n_st)
obs[runif(n_d * n_st) < 0.1] <- NA
obs[4, ] <- NA                                                      # all-NA date
obs[8, !is.na(obs[8, ])] <- 0                                       # all-zero date
syn_obs <- data.table::data.table(Date = as.Date("2020-01-01") + 0:(n_d - 1), obs)
data.table::setnames(syn_obs, c("Date", cod))
syn_coord <- data.table::data.table(Cod = cod, X = X, Y = Y, Z = 100 + seq_len(n_st))
syn_shp <- terra::vect(sprintf("POLYGON ((%f %f, %f %f, %f %f, %f %f, %f %f))", xmin, ymin, xmin + W, ymin,
                               xmin + W, ymin + H, xmin, ymin + H, xmin, ymin), crs = "EPSG:32717")
put_inputs("syn_a", syn_obs, syn_coord, syn_shp)
sv <- c("S005", "S017")
put_text("syn_a", "stat_validation", sv)
for (pw in c(2, 3)) {
  o <- IDW(syn_obs, syn_coord, syn_shp, grid_resolution = 1, p = pw, n_round = NULL, training = 1,
           stat_validation = sv, Rain_threshold = NULL, save_model = FALSE)
  put_stack("syn_a", sprintf("idw_p%d", pw), o$Ensamble)
  put_validation("syn_a", sprintf("idw_p%d.validation", pw), o$Validation)
}
o <- Cressman(syn_obs, syn_coord, syn_shp, grid_resolution = 1, search_radius = c(5, 10, 20, 50), training = 1,
              stat_validation = sv, Rain_threshold = NULL, save_model = FALSE)
put_text("syn_a", "crs.radius_names", names(o$Ensamble))
for (i in seq_along(o$Ensamble)) {
  put_stack("syn_a", sprintf("crs%d", i), o$Ensamble[[i]])
  put_validation("syn_a", sprintf("crs%d.validation", i), o$Validation[[i]])
}

write.csv(manifest, file.path(outdir, "manifest.csv"), row.names = FALSE)
writeLines(c(sprintf("R %s", getRversion()), sprintf("InterpolateR %s", utils::packageVersion("InterpolateR")),
             sprintf("terra %s", utils::packageVersion("terra")), format(Sys.time(), tz = "UTC", usetz = TRUE)),
           file.path(outdir, "versions.txt"))
cat("golden vectors of the real package written to", outdir, "\n")
