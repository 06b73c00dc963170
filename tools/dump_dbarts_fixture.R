#!/usr/bin/env Rscript
# Dumps a REAL dbarts 0.9-8 posterior + the reference's own outputs for the hot path, to pin the oracle
# (oracle/oracle_np.py is "parity unpinned" until this has been run on a box that has R; this image has none).
#
#   Rscript tools/dump_dbarts_fixture.R /path/to/BART-Int tests/golden/dbarts_fixture.json
#
# Needs: R, dbarts 0.9-8 (README.md:102-106 of the reference), data.tree, matrixStats, mvtnorm, msm, lhs, jsonlite.
args <- commandArgs(TRUE)
ref <- if (length(args) >= 1) args[1] else "."
out <- if (length(args) >= 2) args[2] else "dbarts_fixture.json"
suppressMessages({ library(dbarts); library(matrixStats); library(jsonlite) })
source(file.path(ref, "src/BARTInt.R"))          # sampleIntegrals, getTree, ...

set.seed(20260101)
dim <- 2; n <- 40
trainX <- replicate(dim, runif(n))
trainY <- exp(-rowSums((trainX - 0.5)^2) * 4)
sink("/dev/null")
model <- bart(trainX, trainY, keeptrees = TRUE, keepevery = 5L, nskip = 200, ndpost = 300, ntree = 12, k = 2, sigest = 0.1)
sink()
state <- model$fit$state[[1]]
X <- rbind(replicate(dim, runif(64)), trainX[1:4, ], c(0, 0), c(1, 1), c(0.5, 0.5))
pred <- predict(model, X)                                      # S x N  (BARTInt.R:312)
fixture <- list(
  dbarts_version = as.character(packageVersion("dbarts")),
  m = dim(state@savedTrees)[1], S = dim(state@savedTrees)[2], n = n, d = dim,
  saved_trees = as.vector(state@savedTrees),                   # [tree, sample] column-major
  saved_tree_fits = as.vector(state@savedTreeFits),            # [obs, tree, sample] column-major
  x_train = as.vector(model$fit$data@x), y_train = as.vector(model$fit$data@y),
  cut_points = dbarts:::createCutPoints(model$fit),
  y_min = min(trainY), y_max = max(trainY),
  X = as.vector(X), N = nrow(X),
  predict = as.vector(pred),                                   # S x N column-major
  col_vars = colVars(pred), row_means = rowMeans(pred),
  sample_integrals_uniform = sampleIntegrals(model, dim, "uniform"),
  sample_integrals_gaussian = sampleIntegrals(model, dim, "gaussian"),
  sample_integrals_exponential = sampleIntegrals(model, dim, "exponential")
)
write(toJSON(fixture, digits = NA, auto_unbox = TRUE), out)
cat("wrote", out, "\n")
