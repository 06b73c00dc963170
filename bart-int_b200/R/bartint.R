# bartint.R -- R side of the drop-in (SOURCE ONLY here: no R in this image; see INTEGRATION.md).
#
# source() this file AFTER src/BARTInt.R / src/survey_design/bartMean.R.  It keeps every outer signature
# (mainBARTInt, BARTInt, BART_posterior, computeBART) and replaces only the hot-path calls:
#   sampleIntegrals(model, dim, measure)            src/BARTInt.R:204-223  -> exact leaf-volume kernel
#   predict(model, X) + colVars(.) * weights        src/BARTInt.R:312-314  -> bartint_candidate_var()
#   rowMeans(predict(model, X))                     src/survey_design/bartMean.R:74-76 -> bartint_draw_means()
dyn.load(Sys.getenv("BARTINT_R_SO", "bartint_r.so"))

bartint_measure_code <- c(uniform = 0L, gaussian = 1L, exponential = 2L)

# export the dbarts 0.9-8 sampler state once per fitted model (cached on the model object's environment)
bartint_forest <- function(model) {
  if (!is.null(model$treedraws))      # a BART::wbart / pbart fit (README.md:202: "swap the BART package")
    return(.Call("bartint_R_forest_wbart", model$treedraws$trees, model$treedraws$cutpoints,
                 if (is.null(model$offset)) 0 else model$offset))
  fit <- model$fit
  if (!fit$control@keepTrees) stop("bartint needs bart(..., keeptrees = TRUE)")
  key <- "bartint_handle"
  if (!is.null(attr(model, key))) return(attr(model, key))
  state <- fit$state[[1]]
  y <- fit$data@y
  .Call("bartint_R_forest", state@savedTrees, state@savedTreeFits, fit$data@x,
        dbarts:::createCutPoints(fit), min(y), max(y))
}

# drop-in for src/BARTInt.R:204-223.  The reference integrates only draws 1..ntree (nDraw is read from
# dim(treeFits)[2], SURVEY quirk B1); all_draws = TRUE integrates every saved draw.
sampleIntegrals <- function(model, dim, measure, all_draws = FALSE) {
  h <- bartint_forest(model)
  n_used <- if (all_draws) 0L else dim(model$fit$state[[1]]@treeFits)[2]
  .Call("bartint_R_exact_integrals", h, bartint_measure_code[[measure]], as.integer(n_used))
}

# S x N matrix exactly like predict.bart (only when the caller really wants the matrix)
bartint_predict <- function(model, X) {
  .Call("bartint_R_eval", bartint_forest(model), as.matrix(X) + 0, NULL, 4L)[[3]]
}

# colVars(predict(model, X)) * weights, without the S x N matrix
bartint_candidate_var <- function(model, X, weights = 1) {
  .Call("bartint_R_eval", bartint_forest(model), as.matrix(X) + 0, as.double(weights), 2L)[[2]]
}

# rowMeans(predict(model, X)), without the S x N matrix
bartint_draw_means <- function(model, X) {
  .Call("bartint_R_eval", bartint_forest(model), as.matrix(X) + 0, NULL, 1L)[[1]]
}

# The two-line patches inside the reference (shown in INTEGRATION.md):
#   src/BARTInt.R:312-314   fValues <- predict(model, candidateSet); var <- colVars(fValues) * weights
#                        -> var <- bartint_candidate_var(model, candidateSet, weights)
#   bartMean.R:47,50        likewise with candidateX
#   bartMean.R:74-82        pred <- predict(model, fullData); rowMeans(pred); mean(pred)
#                        -> rm <- bartint_draw_means(model, fullData); meanValue[i] <- mean(rm);
#                           standardDeviation[i] <- sum((rm - meanValue[i])^2) / (length(rm) - 1)

# One library call for the body of the sequential loop (src/BARTInt.R:259-262 + 312-314; with X also
# src/survey_design/bartMean.R:74-82): exact integrals (scaled units) + their rescaled mean/sd, per-candidate
# variance * weights, per-draw means.  The export of the saved trees is pipelined against the kernels.
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
