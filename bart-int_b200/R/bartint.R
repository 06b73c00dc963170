# bartint.R -- R side of the drop-in (SOURCE ONLY here: no R in this image; see INTEGRATION.md).
#
# source() this file AFTER src/BARTInt.R / src/survey_design/bartMean.R.  It keeps every outer signature
# (mainBARTInt, BARTInt, BART_posterior, computeBART) and replaces only the hot-path calls:
#   sampleIntegrals(model, dim, measure)            src/BARTInt.R:204-223  -> exact leaf-volume kernel
#   predict(model, X)  (S3 method predict.bart)     src/BARTInt.R:312,380; bartMean.R:47,74 -> full-matrix kernel
#   predict(model, X) + colVars(.) * weights        src/BARTInt.R:312-314  -> bartint_candidate_var()
#   rowMeans(predict(model, X))                     src/survey_design/bartMean.R:74-76 -> bartint_draw_means()
# Statements about package internals that could not be checked here (no R) are tagged [dbarts-mem] / [BART-mem] /
# [bartMachine-mem]; every forest built from a package's public tree export is CHECKED against that package's own
# predict() on a few training rows before it is used (bartint_check_forest).
dyn.load(Sys.getenv("BARTINT_R_SO", "bartint_r.so"))

bartint_measure_code <- c(uniform = 0L, gaussian = 1L, exponential = 2L)

# All GPUs of the box behind this one R process: bartint_init(0) opens the library's device group (worker thread,
# stream and NCCL communicator per GPU).  After it, bartint_design_step() shards its X matrix over the GPUs.
bartint_init <- function(n_gpus = 0L) .Call("bartint_R_init", as.integer(n_gpus))

# ---- forest handles ---------------------------------------------------------------------------------------------
# Handles are cached in this environment, keyed by the identity of the fitted sampler (a dbarts sampler is a
# reference object: its external pointer and the number of saved draws identify the posterior).  R copies lists by
# value, so nothing can be cached on `model` itself.  bartint_release(model) frees the HBM copy at once; otherwise the
# external pointer's finalizer does when R collects it.
.bartint_cache <- new.env(parent = emptyenv())

bartint_key <- function(model) {
  fit <- model$fit
  if (is.null(fit) || is.null(fit$pointer)) return(NULL)          # not a dbarts sampler: no identity to key on
  paste(format(fit$pointer), fit$control@n.samples, dim(fit$state[[1]]@savedTreeFits)[3], sep = "/")
}

bartint_release <- function(model = NULL) {
  keys <- if (is.null(model)) ls(.bartint_cache) else bartint_key(model)
  for (k in keys) if (!is.null(.bartint_cache[[k]])) {
    .Call("bartint_R_release", .bartint_cache[[k]])
    rm(list = k, envir = .bartint_cache)
  }
  invisible(NULL)
}

# n_check rows of the training data: our predictions must equal the package's own, else the export convention
# (leaf scale / offset / routing) is not what this file assumes and we stop instead of returning shifted numbers.
bartint_check_forest <- function(h, model, x_check, tol = 1e-8) {
  ours <- .Call("bartint_R_eval", h, as.matrix(x_check) + 0, NULL, 4L)[[3]]
  ref <- predict(model, x_check)
  if (!isTRUE(all.equal(dim(ours), dim(ref))) || max(abs(ours - ref)) > tol * max(1, max(abs(ref))))
    stop("bartint: the exported forest does not reproduce predict() of this package version; see INTEGRATION.md")
  invisible(TRUE)
}

bartint_forest <- function(model, handle = NULL) {
  if (!is.null(handle)) return(handle)
  # --- BART::wbart / gbart fits (README.md:202: "swap the BART package") ---
  if (!is.null(model$treedraws)) {
    if (inherits(model, "pbart") || inherits(model, "lbart"))
      stop("bartint: probit / logit BART fits return the latent sum; apply pnorm / plogis yourself via bartint_forest(handle=)")
    # predict.wbart adds fit$mu (wbart) or fit$offset (gbart) [BART-mem]; the shim takes the first that is not NULL
    return(.Call("bartint_R_forest_wbart", model$treedraws$trees, model$treedraws$cutpoints,
                 model$offset, model$mu, model$binaryOffset))
  }
  # --- bartMachine fits [bartMachine-mem] ---
  if (inherits(model, "bartMachine")) return(bartint_forest_bartmachine(model))
  # --- dbarts ---
  fit <- model$fit
  if (!fit$control@keepTrees) stop("bartint needs bart(..., keeptrees = TRUE)")
  key <- bartint_key(model)
  if (!is.null(key) && !is.null(.bartint_cache[[key]])) return(.bartint_cache[[key]])
  state <- fit$state[[1]]
  y <- fit$data@y
  h <- if (is.character(state@savedTrees)) {
    # dbarts 0.9-8 (README.md:102-106): tree strings + per-observation fits, what getTree() parses (BARTInt.R:92-133)
    .Call("bartint_R_forest", state@savedTrees, state@savedTreeFits, fit$data@x,
          dbarts:::createCutPoints(fit), min(y), max(y))
  } else {
    bartint_forest_dbarts_trees(model)
  }
  if (!is.null(key)) assign(key, h, envir = .bartint_cache)
  h
}

# dbarts >= 0.9-9: the saved state is no longer strings; the public export is fit$getTrees() [dbarts-mem]: one row per
# node in pre-order, columns sample, tree, n, var (1-based, -1 = leaf), value (split value | leaf prediction).
bartint_forest_dbarts_trees <- function(model) {
  fit <- model$fit
  tr <- fit$getTrees()
  if (!is.null(tr$chain)) tr <- tr[tr$chain == 1, ]
  S <- max(tr$sample); m <- max(tr$tree)
  id <- (tr$sample - 1L) * m + (tr$tree - 1L)                     # tree index t + s*m, rows already grouped by it
  off <- c(0, cumsum(tabulate(id + 1L, S * m)))
  y <- fit$data@y
  # leaf values are per-tree contributions on the ORIGINAL scale with the mean spread over the trees [dbarts-mem]; the
  # check below stops if this version of dbarts does otherwise
  h <- .Call("bartint_R_forest_values", as.integer(S), as.integer(m), ncol(fit$data@x), as.double(off),
             as.integer(ifelse(tr$var < 0, -1L, tr$var - 1L)), as.double(tr$value), NULL, NULL, 0L, 1.0, 0.0)
  bartint_check_forest(h, model, fit$data@x[seq_len(min(5L, nrow(fit$data@x))), , drop = FALSE])
  h
}

# bartMachine [bartMachine-mem]: extract_raw_node_data(bm, g) gives, for Gibbs sample g, a list of m root nodes with
# fields isLeaf, splitAttributeM (0-based), splitValue, y_pred, left, right; a row goes left iff x <= splitValue.
bartint_forest_bartmachine <- function(bm) {
  S <- bm$num_iterations_after_burn_in; m <- bm$num_trees
  var <- integer(0); val <- numeric(0); off <- 0
  flat <- function(node) {
    if (isTRUE(node$isLeaf) || is.null(node$left)) {
      var[length(var) + 1L] <<- -1L; val[length(val) + 1L] <<- node$y_pred
    } else {
      var[length(var) + 1L] <<- as.integer(node$splitAttributeM); val[length(val) + 1L] <<- node$splitValue
      flat(node$left); flat(node$right)
    }
  }
  for (g in seq_len(S)) for (root in bartMachine::extract_raw_node_data(bm, g = g)) {
    flat(root); off <- c(off, length(var))
  }
  h <- .Call("bartint_R_forest_values", as.integer(S), as.integer(m), as.integer(bm$p), as.double(off), var, val,
             NULL, NULL, 0L, 1.0, 0.0)
  bartint_check_forest(h, bm, bm$X[seq_len(min(5L, nrow(bm$X))), , drop = FALSE])
  h
}

# drop-in for src/BARTInt.R:204-223.  The reference integrates only draws 1..ntree (nDraw is read from
# dim(treeFits)[2], SURVEY quirk B1); all_draws = TRUE integrates every saved draw.
sampleIntegrals <- function(model, dim, measure, all_draws = FALSE, handle = NULL) {
  h <- bartint_forest(model, handle)
  n_used <- if (all_draws) 0L
            else if (!is.null(model$treedraws)) as.integer(strsplit(trimws(sub("\n.*", "", model$treedraws$trees)), " +")[[1]][2])
            else dim(model$fit$state[[1]]@treeFits)[2]
  .Call("bartint_R_exact_integrals", h, bartint_measure_code[[measure]], as.integer(n_used))
}

# S x N matrix exactly like predict.bart
bartint_predict <- function(model, X, handle = NULL) {
  .Call("bartint_R_eval", bartint_forest(model, handle), as.matrix(X) + 0, NULL, 4L)[[3]]
}

# The S3 method the reference's predict(model, X) calls dispatch to (src/BARTInt.R:312,380; bartMean.R:47,74).  Defined
# in the global environment it shadows dbarts' registered method, so the reference runs unmodified; anything this
# file cannot serve (no saved trees, extra arguments such as offset / type) goes to dbarts' own method.
predict.bart <- function(object, newdata, ...) {
  if (missing(newdata) || length(list(...)) > 0L || is.null(object$fit) || !object$fit$control@keepTrees ||
      !isTRUE(getOption("bartint.predict", TRUE)))
    return(getFromNamespace("predict.bart", "dbarts")(object, newdata, ...))
  bartint_predict(object, newdata)
}

# colVars(predict(model, X)) * weights, without the S x N matrix
bartint_candidate_var <- function(model, X, weights = 1, handle = NULL) {
  .Call("bartint_R_eval", bartint_forest(model, handle), as.matrix(X) + 0, as.double(weights), 2L)[[2]]
}

# rowMeans(predict(model, X)), without the S x N matrix
bartint_draw_means <- function(model, X, handle = NULL) {
  .Call("bartint_R_eval", bartint_forest(model, handle), as.matrix(X) + 0, NULL, 1L)[[1]]
}

# Batch design: the k candidates of largest variance, order(var, decreasing = TRUE)[1:k] (1-based); k = 1 is the first
# element of which(var == max(var)) (BARTInt.R:315,387).  The random tie break of the reference stays in R.
bartint_topk <- function(values, k = 1L) .Call("bartint_R_topk", as.double(values), as.integer(k)) + 1L

# The two-line patches inside the reference (shown in INTEGRATION.md):
#   src/BARTInt.R:312-314   fValues <- predict(model, candidateSet); var <- colVars(fValues) * weights
#                        -> var <- bartint_candidate_var(model, candidateSet, weights)
#   bartMean.R:47,50        likewise with candidateX
#   bartMean.R:74-82        pred <- predict(model, fullData); rowMeans(pred); mean(pred)
#                        -> rm <- bartint_draw_means(model, fullData); meanValue[i] <- mean(rm);
#                           standardDeviation[i] <- sum((rm - meanValue[i])^2) / (length(rm) - 1)

# One library call for the body of the sequential loop (src/BARTInt.R:259-262 + 312-314; with X also
# src/survey_design/bartMean.R:74-82): exact integrals (scaled units) + their rescaled mean/sd, per-candidate
# variance * weights, per-draw means.  The export of the saved trees is pipelined against the kernels; after
# bartint_init() the rows of X are sharded over all GPUs.  (dbarts 0.9-8 state only; other packages: the calls above.)
bartint_design_step <- function(model, candidateSet = NULL, weights = NULL, measure = "uniform", X = NULL,
                                all_draws = FALSE) {
  fit <- model$fit
  state <- fit$state[[1]]
  y <- fit$data@y
  n_used <- if (all_draws) 0L else dim(state@treeFits)[2]
  r <- .Call("bartint_R_design_step", state@savedTrees, state@savedTreeFits, fit$data@x,
             dbarts:::createCutPoints(fit), min(y), max(y),
             if (is.null(candidateSet)) NULL else as.matrix(candidateSet) + 0,
             if (is.null(weights)) NULL else as.double(weights),
             if (is.null(X)) NULL else as.matrix(X) + 0, bartint_measure_code[[measure]], as.integer(n_used))
  names(r) <- c("integrals", "integral_mean_sd", "cand_var", "draw_mean", "draw_mean_var")
  r
}
