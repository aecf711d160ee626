# testthat file for a maintainer with R, RMINC, its test data and a B200: NOT executed in this repository's CI (R is
# absent from the build image; the Python mirror in tests/ runs the same comparisons against the restated reference).
# It compares every routine of rminc_gpu.R with RMINC's own CPU result on RMINC's test data, in the style of
# tests/testthat/test_mincLm.R, test_vertexLm.R, test_anatLm.R, test_mincAnova.R, test_mincFDR.R and test_mincTFCE.R.
#
#   library(RMINC); dyn.load("rmincgpu.so"); source("rminc_gpu.R"); rmincgpu_install(); testthat::test_file("test_rminc_gpu.R")
library(testthat)
context("rminc-b200 against RMINC")

dataPath <- file.path(tempdir(), "rminctestdata/")
if (!dir.exists(dataPath)) getRMINCTestData(tempdir())
gf <- read.csv(file.path(dataPath, "test_data_set.csv"))
mask <- file.path(dataPath, "testminc-mask.mnc")
tol <- 1e-9

test_that("mincLm: every column of every voxel, masked rows zero", {
  ref <- verboseRun(mincLm(jacobians_fixed_2 ~ Sex + Age, gf, mask = mask))
  plm <- RMINC:::parseLmFormula(jacobians_fixed_2 ~ Sex + Age, gf, NULL)
  got <- mincLm_c_wrapper(plm, mask, 1, 99999999)
  inside <- mincGetVolume(mask) > 0.5
  expect_equal(unclass(got)[inside, ], unclass(ref)[inside, ], tolerance = tol, check.attributes = FALSE)
  expect_true(all(got[!inside, ] == 0))
})

test_that("minc2_model's own .Call, files read by the library's MINC 2 reader (libminc-written, chunked + deflated files)", {
  # the one check this repository cannot run itself: real libminc / libhdf5 files through minc2_reader.cpp
  files <- as.character(gf$jacobians_fixed_2)
  X <- model.matrix(~ Sex + Age, gf)
  ref <- .Call("minc2_model", files, matrix(NA), X, NULL, 1, mask, 1, 99999999, NULL, NULL, "lm", PACKAGE = "RMINC")
  got <- .Call("rmincgpu_minc2_model", files, matrix(NA), X, NULL, 1, mask, 1, 99999999, NULL, NULL, "lm", PACKAGE = "rmincgpu")
  inside <- mincGetVolume(mask) > 0.5
  expect_equal(got[inside, ], ref[inside, ], tolerance = tol, check.attributes = FALSE)
  old <- options(RMINC_GPU_READER = TRUE)
  plm <- RMINC:::parseLmFormula(jacobians_fixed_2 ~ Sex + Age, gf, NULL)
  expect_equal(mincLm_c_wrapper(plm, mask, 1, 99999999)[inside, ], ref[inside, ], tolerance = tol, check.attributes = FALSE)
  options(old)
})

test_that("vertexLm / anatLm native loop keeps its signature", {
  Y <- matrix(rnorm(500 * nrow(gf)), ncol = nrow(gf))
  X <- model.matrix(~ Sex + Age, gf)
  ref <- .Call("vertex_lm_loop", Y, matrix(NA), X, PACKAGE = "RMINC")
  expect_equal(vertex_lm_loop_gpu(Y, matrix(NA), X), ref, tolerance = tol)
  w <- runif(nrow(gf), 0.5, 2)
  refw <- .Call("vertex_wlm_loop", Y, matrix(NA), X, w, PACKAGE = "RMINC")
  expect_equal(vertex_wlm_loop_gpu(Y, matrix(NA), X, w), refw, tolerance = tol)
})

test_that("mincAnova F columns", {
  ref <- verboseRun(mincAnova(jacobians_fixed_2 ~ Sex * Age, gf, mask = mask))
  X <- model.matrix(~ Sex * Age, gf)
  got <- per_voxel_anova_gpu(as.character(gf$jacobians_fixed_2), X, mask)
  inside <- mincGetVolume(mask) > 0.5
  expect_equal(got[inside, ], unclass(ref)[inside, ], tolerance = tol, check.attributes = FALSE)
  old <- options(RMINC_GPU_READER = TRUE)             # per_voxel_anova's own arguments, files read by the library
  got2 <- per_voxel_anova_gpu(as.character(gf$jacobians_fixed_2), X, mask)
  options(old)
  expect_equal(got2[inside, ], unclass(ref)[inside, ], tolerance = tol, check.attributes = FALSE)
})

test_that("mincRandomize: same null as the R loop for the same seed", {
  lmod <- verboseRun(mincLm(jacobians_fixed_2 ~ Sex, gf, mask = mask))
  set.seed(20); ref <- verboseRun(RMINC:::mincRandomize_core(lmod, R = 8, alternative = "two.sided"))
  set.seed(20); got <- mincRandomize_core(lmod, R = 8, alternative = "two.sided")
  expect_equal(got, ref, tolerance = tol, check.attributes = FALSE)
  expect_equal(dim(got), c(8L, 2L))
})

test_that("mincFDR thresholds and q-values of one column", {
  lmod <- verboseRun(mincLm(jacobians_fixed_2 ~ Sex, gf, mask = mask))
  ref <- verboseRun(mincFDR(lmod, mask = mask))
  inside <- mincGetVolume(mask) > 0.5
  res <- rmincgpu_fdr_column(lmod[, "tvalue-SexM"], "t", attr(lmod, "df")[[3]], as.double(inside))
  expect_equal(res$qvalue[inside], ref[inside, "qvalue-tvalue-SexM"], tolerance = 1e-8)
  expect_equal(res$thresholds, attr(ref, "thresholds")[, "tvalue-SexM"], tolerance = 1e-6, check.attributes = FALSE)
})

test_that("TFCE: graph routine equals RMINC's, volume routine equals vertexTFCE on neighbour_list (test_mincTFCE.R:153-170)", {
  vol <- mincGetVolume(gf$jacobians_fixed_2[1])
  sizes <- minc.dimensions.sizes(gf$jacobians_fixed_2[1])
  nl <- RMINC:::neighbour_list(sizes[3], sizes[2], sizes[1], 6)
  ref <- vertexTFCE(as.numeric(vol), nl, weights = .001)
  expect_equal(graph_tfce_gpu(as.numeric(vol), nl, weights = rep(.001, length(vol))), ref, tolerance = 1e-9)
  hmax <- max(abs(vol))
  got <- .Call("rmincgpu_volume_tfce", as.numeric(vol), as.double(sizes), 6L, .001, hmax / 100, .5, 2, 0L, PACKAGE = "rmincgpu")
  expect_false(any(abs(got - ref) > 1e-4 * max(abs(ref))))     # thresholds differ by the floating residue of the R loop
})

test_that("RMINC's own mincLm reaches the GPU after rmincgpu_install()", {
  # the advisor's finding: sourcing rminc_gpu.R alone leaves RMINC::mincLm on the CPU .Call("minc2_model")
  launches <- function() .Call("rmincgpu_launch_count", PACKAGE = "rmincgpu")
  before <- launches()
  got <- verboseRun(mincLm(jacobians_fixed_2 ~ Sex + Age, gf, mask = mask))
  expect_gt(launches(), before)
  rmincgpu_uninstall()
  ref <- verboseRun(mincLm(jacobians_fixed_2 ~ Sex + Age, gf, mask = mask))
  rmincgpu_install()
  inside <- mincGetVolume(mask) > 0.5
  expect_equal(unclass(got)[inside, ], unclass(ref)[inside, ], tolerance = tol, check.attributes = FALSE)
})

test_that("one upload: mincLm, mincRandomize and mincFDR on the resident dataset", {
  ds <- rmincgpu_dataset(gf$jacobians_fixed_2, mask = mask)
  X <- model.matrix(~ Sex + Age, gf)
  ref <- verboseRun(mincLm(jacobians_fixed_2 ~ Sex + Age, gf, mask = mask))
  inside <- mincGetVolume(mask) > 0.5
  expect_equal(rmincgpu_dataset_lm(ds, X)[inside, ], unclass(ref)[inside, ], tolerance = tol, check.attributes = FALSE)
  set.seed(1); d <- rmincgpu_dataset_randomize(ds, ref, R = 20)
  expect_equal(dim(d), c(20L, 3L))
  q <- rmincgpu_dataset_fdr(ds, which(colnames(ref) == "tvalue-Age"), "t", attr(ref, "df")[[2]])
  expect_true(all(q$qvalue[inside] >= 0 & q$qvalue[inside] <= 1))
  rmincgpu_dataset_free(ds)
})
