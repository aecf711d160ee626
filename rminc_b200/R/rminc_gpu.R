# rminc_gpu.R -- R wrappers that make the GPU path a drop-in inside RMINC.
#
# Either add this file to RMINC's R/ directory (its definitions then replace RMINC's own inside the namespace), or
# source() it after library(RMINC) AND call rmincgpu_install(): RMINC::mincLm, mincRandomize and mincTFCE resolve
# mincLm_c_wrapper / mincRandomize_core inside namespace:RMINC, so definitions that only live in the global
# environment are never reached by them -- rmincgpu_install() assigns them into the namespace
# (utils::assignInNamespace) and rmincgpu_uninstall() puts RMINC's own back.  It re-defines the call sites of the
# native per-voxel OLS loop and nothing else; formula handling, attributes, classes and every downstream consumer
# (mincFDR, AIC/BIC, mincWriteVolume, ...) are RMINC's own.
#
#   mincLm_c_wrapper   R/minc_voxel_statistics.R:695-712   -> .Call("rmincgpu_lm")
#   vertex_lm_loop     call sites R/minc_vertex_statistics.R:408-414, R/minc_anatomy.R:1144-1150
#                                                           -> .Call("rmincgpu_vertex_lm_loop")
#   mincRandomize_core R/minc_voxel_statistics.R:1441-1531 -> .Call("rmincgpu_randomize")
#
# options(RMINC_GPU_N = k) shards voxels over k GPUs of the box (default 1); options(RMINC_GPU_DEVICE = d) picks the
# device of the single-GPU routines (default 0).

.rmincgpu_saved <- new.env()

# Make RMINC's own mincLm / mincRandomize / mincTFCE.mincLm use the GPU: replace the two internal functions they call.
rmincgpu_install <- function() {
  for (nm in c("mincLm_c_wrapper", "mincRandomize_core")) {
    if (is.null(.rmincgpu_saved[[nm]])) .rmincgpu_saved[[nm]] <- get(nm, envir = asNamespace("RMINC"))
    utils::assignInNamespace(nm, get(nm, envir = globalenv()), ns = "RMINC")
  }
  invisible(TRUE)
}
rmincgpu_uninstall <- function() {
  for (nm in ls(.rmincgpu_saved)) utils::assignInNamespace(nm, .rmincgpu_saved[[nm]], ns = "RMINC")
  invisible(TRUE)
}

rmincgpu_n <- function(parallel = NULL) {
  if (!is.null(parallel)) return(as.integer(parallel[2]))
  as.integer(getOption("RMINC_GPU_N", 1L))
}

# Stage the n volumes as the columns of a V x n double matrix, in file dimension order -- the
# same element order minc2_model uses for its output rows (src/modelling_functions.c:1059).
rmincgpu_stage <- function(filenames) {
  first <- mincGetVolume(filenames[1])
  Y <- matrix(0, nrow = length(first), ncol = length(filenames))
  Y[, 1] <- first
  for (j in seq_along(filenames)[-1]) Y[, j] <- mincGetVolume(filenames[j])
  Y
}

# Replacement for RMINC:::mincLm_c_wrapper.  `mask` may be a file name or NULL.
# options(RMINC_GPU_READER = TRUE): the .Call keeps minc2_model's own eleven arguments (R/minc_voxel_statistics.R:695-712)
# and the library reads the MINC 2 files itself -- all host cores, uint16 / float32 voxels shipped in their stored type --
# instead of staging doubles through mincGetVolume; the GPU count comes from options(RMINC_GPU_N).
mincLm_c_wrapper <- function(plm, mask, mask_min, mask_max, n_gpus = rmincgpu_n()) {
  if (isTRUE(getOption("RMINC_GPU_READER", FALSE))) {
    old <- options(RMINC_GPU_N = as.integer(n_gpus)); on.exit(options(old))
    have_mask <- if (is.null(mask)) 0 else 1
    mm <- if (is.logical(plm$mmatrix)) plm$mmatrix else as.matrix(plm$mmatrix)
    right <- if (is.logical(plm$data.matrix.right)) plm$data.matrix.right else as.character(plm$data.matrix.right)
    return(.Call("rmincgpu_minc2_model", as.character(plm$data.matrix.left), right, mm, NULL, as.double(have_mask),
                 if (is.null(mask)) character(0) else as.character(mask), as.double(mask_min), as.double(mask_max), NULL, NULL,
                 "lm", PACKAGE = "rmincgpu"))
  }
  Y <- rmincgpu_stage(as.character(plm$data.matrix.left))
  m <- if (is.null(mask)) NULL else as.double(mincGetVolume(mask))
  if (!is.logical(plm$data.matrix.right)) {
    # a file term on the right-hand side: per-voxel design [mmatrix | z_v] (one GPU)
    Z <- rmincgpu_stage(as.character(plm$data.matrix.right))
    mm <- if (is.logical(plm$mmatrix)) plm$mmatrix else as.matrix(plm$mmatrix)
    return(.Call("rmincgpu_lm_cov", Y, Z, mm, m, as.double(mask_min), as.double(mask_max), PACKAGE = "rmincgpu"))
  }
  .Call("rmincgpu_lm", Y, as.matrix(plm$mmatrix), m, as.double(mask_min), as.double(mask_max),
        as.integer(n_gpus), PACKAGE = "rmincgpu")
}

# mincLm(parallel = c("local", N)) on the GPU path means N GPUs, not N forked processes: a CUDA
# context must never be forked.  RMINC's own mincLm is run with parallel = NULL -- its sequential branch
# (R/minc_voxel_statistics.R:841-847) calls mincLm_c_wrapper, which rmincgpu_install() has pointed at the GPU wrapper
# above -- and n_gpus = N travels in options(RMINC_GPU_N); the masked rows come back as 0 in every column, as the
# reference's parallel branch produces (:898-905).
mincLm_gpu <- function(formula, data = NULL, subset = NULL, mask = NULL, maskval = NULL, parallel = NULL) {
  if (!identical(get("mincLm_c_wrapper", envir = asNamespace("RMINC")), get("mincLm_c_wrapper", envir = globalenv())))
    stop("call rmincgpu_install() first: RMINC::mincLm still resolves RMINC's CPU mincLm_c_wrapper")
  old <- options(RMINC_GPU_N = rmincgpu_n(parallel)); on.exit(options(old))
  cl <- match.call(); cl[[1]] <- quote(RMINC::mincLm); cl$parallel <- NULL
  eval(cl, parent.frame())
}

# vertexLm / anatLm: only the routine name changes.
vertex_lm_loop_gpu <- function(data.matrix.left, data.matrix.right, mmatrix)
  .Call("rmincgpu_vertex_lm_loop", data.matrix.left, data.matrix.right, mmatrix, PACKAGE = "rmincgpu")

# anatLm(weights = w): .Call("vertex_wlm_loop", ...) at R/minc_anatomy.R:1135-1142
vertex_wlm_loop_gpu <- function(data.matrix.left, data.matrix.right, mmatrix, weights)
  .Call("rmincgpu_vertex_wlm_loop", data.matrix.left, data.matrix.right, mmatrix, as.double(weights),
        PACKAGE = "rmincgpu")

# Replacement for RMINC:::mincRandomize_core (identity post_proc only).  The permutation indices
# are drawn here with R's own sample(), so a seed gives the same null distribution as the
# reference; the refits happen on the GPU with Y resident instead of update(lmod, data = shuf).
mincRandomize_core <- function(x, R = 500, replace = FALSE, parallel = NULL,
                               columns = grep("tvalue-", colnames(x)),
                               alternative = c("two.sided", "greater"), post_proc = identity, ...) {
  alternative <- match.arg(alternative)
  if (!identical(post_proc, identity))
    stop("a post_proc runs in its own routine: mincTFCE_mincLm_gpu() for mincTFCE.matrix, vertexTFCE_randomize_gpu() for vertexTFCE.matrix")
  lmod <- x
  n <- nrow(attr(lmod, "data"))
  idx <- vapply(seq_len(R), function(i) sample(seq_len(n), replace = replace), integer(n))
  Y <- rmincgpu_stage(attr(lmod, "filenames"))
  cl <- attr(lmod, "call")
  mask <- eval(cl$mask, parent.frame()); maskval <- eval(cl$maskval, parent.frame())
  m <- if (is.null(mask)) NULL else as.double(mincGetVolume(mask))
  lo <- if (is.null(maskval)) 1 else maskval
  hi <- if (is.null(maskval)) 99999999 else maskval
  # all predictor columns move together (:1457-1474): the permuted design is the model matrix
  # with its rows permuted; a resample that empties a factor level makes the refit singular and
  # the call fails, as the reference's rbind of ragged rows would.
  dist <- .Call("rmincgpu_randomize", Y, attr(lmod, "model"), m, as.double(lo), as.double(hi), idx,
                as.integer(columns), alternative == "two.sided", rmincgpu_n(parallel), PACKAGE = "rmincgpu")
  colnames(dist) <- colnames(lmod)[columns]
  dist
}

# The same permutation loop on volumes kept in their stored type.  `stored` is what the staging routine of
# INTEGRATION.md section 5 returns: list(U = raw vector (V x n samples), dtype = 1L (uint16) / 2L (float32), V,
# scale, offset (nslices x n), voxel_min, slice_len).  2 (or 4) bytes per sample cross the host link, once.
mincRandomize_stored <- function(lmod, stored, R = 500, replace = FALSE, parallel = NULL,
                                 columns = grep("tvalue-", colnames(lmod)),
                                 alternative = c("two.sided", "greater"), mask = NULL, maskval = NULL) {
  alternative <- match.arg(alternative)
  n <- nrow(attr(lmod, "data"))
  idx <- vapply(seq_len(R), function(i) sample(seq_len(n), replace = replace), integer(n))
  m <- if (is.null(mask)) NULL else as.double(mincGetVolume(mask))
  lo <- if (is.null(maskval)) 1 else maskval
  hi <- if (is.null(maskval)) 99999999 else maskval
  dist <- .Call("rmincgpu_randomize_stored", stored$U, as.integer(stored$dtype), as.double(stored$V), stored$scale,
                stored$offset, as.double(stored$voxel_min), as.double(stored$slice_len), attr(lmod, "model"), m,
                as.double(lo), as.double(hi), idx, as.integer(columns), alternative == "two.sided",
                rmincgpu_n(parallel), PACKAGE = "rmincgpu")
  colnames(dist) <- colnames(lmod)[columns]
  dist
}

# mincAnova / vertexAnova / anatAnova: the native loops (.Call("per_voxel_anova", ...),
# R/minc_voxel_statistics.R:557-566; .Call("vertex_anova_loop", ...), R/minc_vertex_statistics.R:314-320)
# become one routine taking the staged matrix and model.matrix's "assign" attribute.
per_voxel_anova_gpu <- function(filenames, mmatrix, mask = NULL, mask_min = 1, mask_max = 99999999) {
  if (isTRUE(getOption("RMINC_GPU_READER", FALSE)))      # per_voxel_anova's own arguments; the library reads the files
    return(.Call("rmincgpu_per_voxel_anova", as.character(filenames), mmatrix, as.integer(attr(mmatrix, "assign")),
                 as.double(if (is.null(mask)) 0 else 1), if (is.null(mask)) character(0) else as.character(mask),
                 as.double(mask_min), as.double(mask_max), PACKAGE = "rmincgpu"))
  Y <- rmincgpu_stage(filenames)
  m <- if (is.null(mask)) NULL else as.double(mincGetVolume(mask))
  .Call("rmincgpu_anova", Y, mmatrix, as.integer(attr(mmatrix, "assign")), m, as.double(mask_min),
        as.double(mask_max), PACKAGE = "rmincgpu")
}
vertex_anova_loop_gpu <- function(data.matrix, mmatrix)
  .Call("rmincgpu_anova", data.matrix, mmatrix, as.integer(attr(mmatrix, "assign")), NULL, 1, 99999999,
        PACKAGE = "rmincgpu")


# mincFDR's inner loop body on the GPU (R/minc_FDR.R:385-505): p-values, p.adjust(., "fdr") and the thresholds of one
# column.  stat_type "t" / "F"; df as in attr(buffer, "df")[[column]]; mask a numeric vector or NULL.
# method: mincFDR's (R/minc_FDR.R:429-438).  For "pFDR" / "fastqvalue" / "qvalue" pi0 is estimated as fast.qvalue does
# (R/minc_misc.R:437-451): mean(p >= lambda) / (1 - lambda) from the device, smooth.spline(df = 3) here.
rmincgpu_fdr_column <- function(stat, stat_type, df, mask = NULL, method = "FDR") {
  code <- c(FDR = 0L, p.adjust = 0L, BY = 1L, pFDR = 2L, fastqvalue = 2L, qvalue = 3L)[[method]]
  type <- if (stat_type == "F") 2L else 1L
  m <- if (is.null(mask)) NULL else as.double(mask)
  pi0 <- 1
  if (code >= 2L) {
    lambda <- if (code == 2L) seq(0, 0.9, 0.05) else seq(0.05, 0.95, 0.05)
    frac <- .Call("rmincgpu_fdr_pi0", as.double(stat), type, as.double(df), m, as.double(lambda), PACKAGE = "rmincgpu")
    spi0 <- smooth.spline(lambda, frac / (1 - lambda), df = 3)
    pi0 <- min(predict(spi0, x = max(lambda))$y, 1)
  }
  res <- .Call("rmincgpu_fdr_method", as.double(stat), type, as.double(df), m, code, as.double(pi0), PACKAGE = "rmincgpu")
  list(qvalue = res[[1]], thresholds = res[[2]], pi0 = pi0)
}


# vertexTFCE: the Rcpp exports graph_tfce_wqu / graph_tfce (R/RcppExports.R:32-38) and the randomisation of
# vertexTFCE.vertexLm (R/minc_vertex_statistics.R:892-948) on the GPU.  `adjacencies`: the 0-based adjacency list
# vertexTFCE.numeric builds (lapply(igraph::as_adj_list(surface), `-`, 1)).
graph_tfce_gpu <- function(map, adjacencies, E = .5, H = 2, nsteps = 100, weights = rep(1, NROW(map)),
                           side = c("both", "positive", "negative")) {
  side <- match(match.arg(side), c("both", "positive", "negative")) - 1L
  .Call("rmincgpu_graph_tfce", map, lapply(adjacencies, as.integer), as.double(E), as.double(H), as.integer(nsteps),
        as.double(weights), side, PACKAGE = "rmincgpu")
}

vertexTFCE_randomize_gpu <- function(lmod, adjacencies, R = 500, alternative = c("two.sided", "greater"), E = .5, H = 2,
                                     nsteps = 100, weights = NULL, side = c("both", "positive", "negative"),
                                     replace = FALSE, column = 1) {
  alternative <- match.arg(alternative)
  side <- match(match.arg(side), c("both", "positive", "negative")) - 1L
  columns <- grep("tvalue-", colnames(lmod))
  n <- nrow(attr(lmod, "data"))
  idx <- vapply(seq_len(R), function(i) sample(seq_len(n), replace = replace), integer(n))
  Y <- vertexTable(attr(lmod, "filenames"), column = column)
  w <- if (is.null(weights)) rep(1, nrow(Y)) else as.double(weights)
  dist <- .Call("rmincgpu_randomize_tfce", Y, attr(lmod, "model"), idx, as.integer(columns), alternative == "two.sided",
                lapply(adjacencies, as.integer), w, as.double(E), as.double(H), as.integer(nsteps), side,
                PACKAGE = "rmincgpu")
  colnames(dist) <- colnames(lmod)[columns]
  dist
}


# mincTFCE without the temporary files and the external TFCE program (R/minc_voxel_statistics.R:1186-1283): the tree's
# graph TFCE on the voxel grid.  x: a mincSingleDim vector or a matrix with a likeVolume; NA / NaN -> 0 as
# mincTFCE.matrix does.
mincTFCE_gpu <- function(x, d = 0.1, E = .5, H = 2.0, side = c("both", "positive", "negative"),
                         like_volume = likeVolume(x), connectivity = 6L) {
  side <- match(match.arg(side), c("both", "positive", "negative")) - 1L
  sizes <- minc.dimensions.sizes(like_volume)
  vv <- prod(abs(minc.separation.sizes(like_volume)))
  vals <- unclass(x); vals[!is.finite(vals) & is.na(vals)] <- 0
  run <- function(m, step) .Call("rmincgpu_volume_tfce", m, as.double(sizes), as.integer(connectivity), as.double(vv),
                                 as.double(step), as.double(E), as.double(H), side, PACKAGE = "rmincgpu")
  if (is.matrix(vals) && length(d) > 1) {
    out <- mapply(function(j, step) run(as.double(vals[, j]), step), seq_len(ncol(vals)), d)
  } else {
    storage.mode(vals) <- "double"
    out <- run(vals, d[1])
  }
  if (is.matrix(vals)) colnames(out) <- colnames(x)
  likeVolume(out) <- like_volume
  out
}

# mincTFCE.mincLm (:1382-1439): enhanced t columns + the null of their maxima, all R permutations in one call.
mincTFCE_mincLm_gpu <- function(lmod, R = 500, alternative = c("two.sided", "greater"), d = 0.1, E = .5, H = 2.0,
                                side = c("both", "positive", "negative"), replace = FALSE, parallel = NULL,
                                connectivity = 6L) {
  alternative <- match.arg(alternative); side_name <- match.arg(side)
  side <- match(side_name, c("both", "positive", "negative")) - 1L
  columns <- grep("tvalue-", colnames(lmod))
  like_vol <- likeVolume(lmod)
  original <- mincTFCE_gpu(lmod[, columns], d = d, E = E, H = H, side = side_name, like_volume = like_vol,
                           connectivity = connectivity)
  n <- nrow(attr(lmod, "data"))
  idx <- vapply(seq_len(R), function(i) sample(seq_len(n), replace = replace), integer(n))
  Y <- rmincgpu_stage(attr(lmod, "filenames"))
  cl <- attr(lmod, "call")
  mask <- eval(cl$mask, parent.frame()); maskval <- eval(cl$maskval, parent.frame())
  m <- if (is.null(mask)) NULL else as.double(mincGetVolume(mask))
  lo <- if (is.null(maskval)) 1 else maskval
  hi <- if (is.null(maskval)) 99999999 else maskval
  dist <- .Call("rmincgpu_randomize_volume_tfce", Y, attr(lmod, "model"), m, as.double(lo), as.double(hi), idx,
                as.integer(columns), alternative == "two.sided", as.double(minc.dimensions.sizes(like_vol)),
                as.integer(connectivity), as.double(prod(abs(minc.separation.sizes(like_vol)))), as.double(d),
                as.double(E), as.double(H), side, rmincgpu_n(parallel), PACKAGE = "rmincgpu")
  colnames(dist) <- colnames(lmod)[columns]
  structure(list(tfce = original, randomization_dist = dist,
                 args = list(call = cl, side = side_name, alternative = alternative)),
            class = c("mincTFCE_randomization", "minc_randomization"))
}


# ---- the resident dataset: Y is staged and uploaded ONCE (sharded over n_gpus devices) and stays in GPU memory behind an
# external pointer; mincLm, mincAnova, mincRandomize and mincFDR then run on the resident shards.  The reference re-reads
# every file for each of them (R/minc_voxel_statistics.R:1477).  The handle frees itself when it is garbage collected;
# rmincgpu_dataset_free() does it at once.
rmincgpu_dataset <- function(filenames, mask = NULL, maskval = NULL, n_gpus = rmincgpu_n()) {
  Y <- rmincgpu_stage(as.character(filenames))
  m <- if (is.null(mask)) NULL else as.double(mincGetVolume(mask))
  lo <- if (is.null(maskval)) 1 else maskval
  hi <- if (is.null(maskval)) 99999999 else maskval
  h <- .Call("rmincgpu_dataset_create", Y, m, as.double(lo), as.double(hi), as.integer(n_gpus), PACKAGE = "rmincgpu")
  structure(list(handle = h, filenames = as.character(filenames), mask = mask), class = "rmincgpu_dataset")
}
rmincgpu_dataset_free <- function(ds) invisible(.Call("rmincgpu_dataset_free", ds$handle, PACKAGE = "rmincgpu"))

# the native step of mincLm / mincAnova on the handle (the matrices RMINC's R layer then decorates)
rmincgpu_dataset_lm <- function(ds, mmatrix) .Call("rmincgpu_dataset_lm", ds$handle, as.matrix(mmatrix), PACKAGE = "rmincgpu")
rmincgpu_dataset_anova <- function(ds, mmatrix)
  .Call("rmincgpu_dataset_anova", ds$handle, mmatrix, as.integer(attr(mmatrix, "assign")), PACKAGE = "rmincgpu")

# mincRandomize on the handle: lmod is the mincLm object fitted on the same files
rmincgpu_dataset_randomize <- function(ds, lmod, R = 500, replace = FALSE, columns = grep("tvalue-", colnames(lmod)),
                                       alternative = c("two.sided", "greater")) {
  alternative <- match.arg(alternative)
  n <- nrow(attr(lmod, "data"))
  idx <- vapply(seq_len(R), function(i) sample(seq_len(n), replace = replace), integer(n))
  dist <- .Call("rmincgpu_dataset_randomize", ds$handle, attr(lmod, "model"), idx, as.integer(columns),
                alternative == "two.sided", PACKAGE = "rmincgpu")
  colnames(dist) <- colnames(lmod)[columns]
  dist
}

# mincFDR's per-column step on the result of the LAST fit on the handle (column: its 1-based column number)
# method as mincFDR's (R/minc_FDR.R:429-438): "FDR" / "p.adjust", "BY", "pFDR" / "fastqvalue", "qvalue"; pi0 for the last two
rmincgpu_dataset_fdr <- function(ds, column, stat_type, df, method = "FDR", pi0 = 1) {
  st <- if (stat_type == "F") 2L else 1L
  if (method %in% c("FDR", "p.adjust"))
    res <- .Call("rmincgpu_dataset_fdr", ds$handle, as.integer(column), st, as.double(df), PACKAGE = "rmincgpu")
  else
    res <- .Call("rmincgpu_dataset_fdr_method", ds$handle, as.integer(column), st, as.double(df),
                 switch(method, BY = 1L, pFDR = 2L, fastqvalue = 2L, qvalue = 3L, stop("unknown FDR method")), as.double(pi0),
                 PACKAGE = "rmincgpu")
  list(qvalue = res[[1]], thresholds = res[[2]])
}
