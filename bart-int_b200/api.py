"""Host-side mirror of the reference's R interface for the posterior-evaluation hot path.

Same names, argument meaning and return shapes as the R functions they stand for, so the parity
tests read like the reference's call sites.  Everything numerical goes through libbartint_cuda.so;
there is no CPU path here.

  R (reference)                                   here
  ----------------------------------------------  -----------------------------------------------
  model <- bart(..., keeptrees=TRUE)              BartModel (state container; the MCMC fit itself
     src/BARTInt.R:256,282,363                    is out of scope -- `bart=` is a pluggable callable,
                                                  default: synthetic posterior from the BART prior)
  sampleIntegrals(model, dim, measure)  :204-223  sampleIntegrals(model, dim, measure)
  predict(model, X)                     :312,380  predict(model, X)              -> (S, N) ndarray
  colVars(fValues) * weights            :314,386  bartint_candidate_var(model, X, weights)
  rowMeans(predict(model, X))  bartMean.R:76,82   bartint_draw_means(model, X)
  BARTInt(...) :225, BART_posterior(...) :329,    BARTInt, BART_posterior, mainBARTInt
  mainBARTInt(...) :398
  computeBART(...)             bartMean.R:16      computeBART
  (loop body :259-262 + :312-314 in one call)     bartint_design_step(model, candidateSet, weights, measure)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import synth
from .forest import FlatForest, Forest


@dataclass
class DbartsSamplerState:
    """What the reference reads from `model$fit` (a dbartsSampler RefClass), BARTInt.R:108-116,122,191,216."""
    saved_trees: np.ndarray       # chr [m, S]      state[[1]]@savedTrees[treeNum, sampleNum]
    saved_tree_fits: np.ndarray   # dbl [n, m, S]   state[[1]]@savedTreeFits[, treeNum, sampleNum]
    x: np.ndarray                 # dbl [n, d]      data@x
    y: np.ndarray                 # dbl [n]         data@y
    cut_points: list              # dbarts:::createCutPoints(sampler)

    @property
    def n_tree(self) -> int:      # ncol(state[[1]]@treeFits), BARTInt.R:191
        return int(self.saved_trees.shape[0])

    @property
    def n_draws(self) -> int:
        return int(self.saved_trees.shape[1])


@dataclass
class BartModel:
    """The `bart` object the reference passes around (`model`); `fit` is `model$fit`."""
    fit: DbartsSamplerState
    y_min: float
    y_max: float
    _forest: Forest | None = field(default=None, repr=False)

    def forest(self) -> Forest:
        """Export the saved trees to HBM once per model (getTree for every tree, BARTInt.R:92-133)."""
        if self._forest is None:
            self._forest = Forest.from_dbarts(self.fit.saved_trees, self.fit.saved_tree_fits, self.fit.x,
                                              self.fit.cut_points, self.y_min, self.y_max)
        return self._forest

    def close(self):
        if self._forest is not None:
            self._forest.close()
            self._forest = None


def synthetic_bart(trainX, trainY, ndpost=5000, keepevery=5, ntree=50, k=2.0, seed=0, numcut=100) -> BartModel:
    """Stand-in for `bart(trainX, trainY, keeptrees=TRUE, keepevery, ndpost, ntree, k)`.

    Keeps the reference's bookkeeping (S = ndpost %/% keepevery saved draws, cut points from the training
    range, response range) but draws the trees from the BART prior instead of running the MCMC sampler,
    and returns them in dbarts' 0.9-8 saved-state format so the exporter is on the path.
    """
    trainX = np.asarray(trainX, np.float64)
    if trainX.ndim == 1:
        trainX = trainX.reshape(-1, 1)
    trainY = np.asarray(trainY, np.float64).ravel()
    S = ndpost // keepevery
    cut_lists = synth.uniform_cutpoints(trainX, numcut)
    ff = synth.prior_forest(seed, S, ntree, trainX, trainY, numcut=numcut, k=k, cut_lists=cut_lists)
    strings, fits = ff.to_dbarts_state(trainX)
    state = DbartsSamplerState(strings, fits, trainX, trainY, cut_lists)
    return BartModel(state, float(trainY.min()), float(trainY.max()))


def model_from_flat(ff: FlatForest, x_train, y_train=None) -> BartModel:
    x_train = np.asarray(x_train, np.float64)
    strings, fits = ff.to_dbarts_state(x_train)
    y = np.zeros(x_train.shape[0]) if y_train is None else np.asarray(y_train, np.float64)
    return BartModel(DbartsSamplerState(strings, fits, x_train, y, ff.cut_lists()), ff.y_min, ff.y_max)


# ---------------------------------------------------------------------------------------------
# drop-in entry points
# ---------------------------------------------------------------------------------------------
def sampleIntegrals(model: BartModel, dim: int, measure: str, all_draws: bool = False) -> np.ndarray:
    """BARTInt.R:204-223.  Per-draw exact integrals in SCALED units (the caller rescales, :260).

    Bug-compatible by default: the reference takes nDraw <- dim(treeFits)[2], the number of TREES, so only
    draws 1..ntree are integrated (quirk B1).  all_draws=True integrates every saved draw.
    """
    f = model.forest()
    if dim != f.d:
        raise ValueError(f"dim={dim} but the model was fit on {f.d} variables")
    n_used = f.S if all_draws else min(f.m, f.S)
    return f.exact_integrals(measure, n_draws_used=n_used)


def predict(model: BartModel, X) -> np.ndarray:
    """predict(model, X): (S, N) matrix in original units, draws in rows (BARTInt.R:312; bartMean.R:47,74)."""
    return model.forest().eval(X, want=("full_pred",))["full_pred"]


def bartint_candidate_var(model: BartModel, X, weights=1.0) -> np.ndarray:
    """colVars(predict(model, X)) * weights without materialising the S x N matrix (BARTInt.R:312-314)."""
    w = np.asarray(weights, np.float64).ravel()
    return model.forest().eval(X, weights=w, want=("point_var",))["point_var"]


def bartint_draw_means(model: BartModel, X) -> np.ndarray:
    """rowMeans(predict(model, X)) without materialising the S x N matrix (bartMean.R:74-76)."""
    return model.forest().eval(X, want=("draw_mean",))["draw_mean"]


def bartint_design_step(model: BartModel, candidateSet=None, weights=1.0, measure="uniform", X=None,
                        all_draws: bool = False) -> dict:
    """One sequential-design step in one library call (bartint_design_step_dbarts098): what BARTInt.R:259-262 and
    :312-314 (and, when X is given, bartMean.R:74-82) compute from a freshly fitted model, with the tree export
    pipelined against the evaluation.  Returns the per-draw integrals in scaled units (`integrals`, the value of
    sampleIntegrals -- first ntree draws unless all_draws, quirk B1), `integral_mean_sd` (rescaled, :260-262),
    `cand_var` (= colVars(predict(model, candidateSet)) * weights) and `draw_mean` / `draw_mean_var`."""
    from .forest import StagedMatrix, design_step_dbarts
    fit = model.fit
    sc = StagedMatrix(candidateSet) if candidateSet is not None else None
    sx = StagedMatrix(X) if X is not None else None
    try:
        return design_step_dbarts(fit.saved_trees, fit.saved_tree_fits, fit.x, fit.cut_points, model.y_min, model.y_max,
                                  mc=sx, candidates=sc, weights=weights, measure=measure,
                                  n_draws_integral=None if all_draws else min(fit.n_tree, fit.n_draws))
    finally:
        for h in (sc, sx):
            if h is not None:
                h.close()


def r_sample_one(v: np.ndarray, rng) -> int:
    """R's sample(v, 1) INCLUDING its scalar quirk (B2): a length-1 numeric v samples from 1:v.
    v holds 1-based indices; returns a 1-based index."""
    v = np.asarray(v)
    if v.size == 1:
        return int(rng.integers(1, int(v[0]) + 1))
    return int(v[rng.integers(0, v.size)])


def _candidates(measure: str, dim: int, n: int, rng, refcompat: bool = True):
    """BARTInt.R:299-309: candidate set and weights.

    refcompat=True returns the weights OBJECT the reference builds, quirks included; `_pick` applies
    R's recycling to it:
      gaussian     `dtnorm(candidateSet, ...)` is a C x d MATRIX, not a length-C vector;
      exponential  `colMeans(dexp(candidateSet))` has length d, not C.
    refcompat=False returns the per-candidate density (product over the coordinates), length C -- what the paper's text
    describes."""
    X = synth.sample_measure(rng, measure, n, dim)
    if measure == "uniform":
        return X, np.ones(1)
    if measure == "gaussian":
        from scipy.special import ndtr
        dens = np.exp(-0.5 * (X - 0.5) ** 2) / math.sqrt(2 * math.pi) / (ndtr(0.5) - ndtr(-0.5))
        return X, (dens if refcompat else dens.prod(axis=1))
    return X, (np.exp(-X).mean(axis=0) if refcompat else np.exp(-X.sum(axis=1)))


def _pick(var_unweighted: np.ndarray, weights: np.ndarray, rng) -> int:
    """`var <- colVars(fValues) * weights; index <- sample(which(var == max(var)), 1)` (BARTInt.R:314-315) with R's
    recycling of whatever shape `weights` has; returns the 1-based index R would use in `candidateSet[index, ]`.

    length 1 or C: plain product.  Length d (exponential quirk): weights[i mod d].  C x d matrix (gaussian quirk): the
    product is a C x d matrix (var recycled down the columns) and `which` returns LINEAR indices into it; an index
    beyond C makes `candidateSet[index, ]` fail in R ("subscript out of bounds") -- here the candidate of that row is
    taken instead, which is what R picks whenever it does not fail."""
    C_ = var_unweighted.size
    w = np.asarray(weights, np.float64)
    if w.ndim == 2:                                   # C x d, column-major linear indices like R
        prod = (var_unweighted[:, None] * w).ravel(order="F")
    else:
        prod = var_unweighted * np.resize(w, C_)
    lin = r_sample_one(np.nonzero(prod == prod.max())[0] + 1, rng)       # quirk B2 inside
    return (lin - 1) % C_ + 1


def BARTInt(dim, trainX, trainY, numNewTraining, FUN, sequential, measure, save_posterior=False,
            save_posterior_dir="results/genz", save_posterior_filename="default", bart=None, rng=None,
            all_draws=False):
    """BARTInt.R:225-327 with the `bart()` fit injected (`bart(trainX, trainY, ndpost=...) -> BartModel`)."""
    rng = np.random.default_rng(0) if rng is None else rng
    bart = synthetic_bart if bart is None else bart
    trainX = np.asarray(trainX, np.float64).reshape(-1, dim)
    trainY = np.asarray(trainY, np.float64).ravel()
    meanValue = np.zeros(numNewTraining)
    standardDeviation = np.zeros(numNewTraining)
    posterior = []
    model = None
    for i in range(1, numNewTraining + 1):
        ymin, ymax = trainY.min(), trainY.max()                                  # :250-251, :277-278
        if model is not None:
            model.close()
        model = bart(trainX, trainY, ndpost=10000 if i == 1 else 5000, keepevery=5, ntree=50, k=2.0,
                     seed=int(rng.integers(1 << 31)))                            # :256 / :282
        integrals = sampleIntegrals(model, dim, measure, all_draws=all_draws)    # :259 / :289
        res, mean, sd = model.forest().finalize_integrals(integrals)             # :260-262 / :290-292
        meanValue[i - 1], standardDeviation[i - 1] = mean, sd
        if save_posterior:
            posterior.append(res)
        if i == 1:
            if numNewTraining == 1:                                              # :267
                break
            continue
        candidateSet, weights = _candidates(measure, dim, 100, rng)              # :299-309
        if sequential:
            var = bartint_candidate_var(model, candidateSet, 1.0)                # :312-314 (colVars; the weights follow)
            index = _pick(var, weights, rng)                                     # :314-315 (quirks B2, B4)
        else:
            index = int(rng.integers(1, 101))                                    # :318
        value = float(np.asarray(FUN(candidateSet[index - 1:index, :])).ravel()[0])   # :320
        trainX = np.vstack([trainX, candidateSet[index - 1]])                    # :321
        trainY = np.append(trainY, value)
    return {"meanValueBART": meanValue, "standardDeviationBART": standardDeviation,
            "trainData": np.column_stack([trainX, trainY]), "model": model, "posterior_samples": posterior}


def BART_posterior(dim, trainX, trainY, numNewTraining, FUN, sequential, measure="uniform", bart=None, rng=None):
    """BARTInt.R:329-396: same loop, every iteration selects a candidate, returns the last model."""
    rng = np.random.default_rng(0) if rng is None else rng
    bart = synthetic_bart if bart is None else bart
    trainX = np.asarray(trainX, np.float64).reshape(-1, dim)
    trainY = np.asarray(trainY, np.float64).ravel()
    model = None
    for _ in range(numNewTraining):
        if model is not None:
            model.close()
        model = bart(trainX, trainY, ndpost=5000, keepevery=5, ntree=50, k=2.0, seed=int(rng.integers(1 << 31)))
        candidateSet, _ = _candidates(measure, dim, 100, rng)
        if sequential:
            var = bartint_candidate_var(model, candidateSet)                     # :380-386
            index = r_sample_one(np.nonzero(var == var.max())[0] + 1, rng)
        else:
            index = int(rng.integers(1, 101))
        value = float(np.asarray(FUN(candidateSet[index - 1:index, :])).ravel()[0])
        trainX = np.vstack([trainX, candidateSet[index - 1]])
        trainY = np.append(trainY, value)
    return {"model": model, "trainData": np.column_stack([trainX, trainY])}


def mainBARTInt(dim, num_iterations, FUN, trainX, trainY, sequential=True, measure="uniform", save_posterior=False,
                save_posterior_dir="results/genz", save_posterior_filename="default", **kw):
    """BARTInt.R:398-422."""
    return BARTInt(dim, trainX, trainY, num_iterations, FUN, sequential, measure, save_posterior,
                   save_posterior_dir, save_posterior_filename, **kw)


def computeBART(nonOneHotTrainX, trainX, trainY, nonOneHotCandidateX, candidateX, candidateY, num_iterations,
                save_posterior=False, save_posterior_dir="results/survey_design", num_cv="default", bart=None,
                rng=None, weights=None):
    """bartMean.R:16-93.  `weights` replaces build_density(...) (a kernel density estimate, out of scope);
    default 1.  One fused device pass per predict() call site (:47 and :74)."""
    rng = np.random.default_rng(0) if rng is None else rng
    bart = synthetic_bart if bart is None else bart
    trainX = np.asarray(trainX, np.float64)
    trainY = np.asarray(trainY, np.float64).ravel()
    candidateX = np.asarray(candidateX, np.float64)
    candidateY = np.asarray(candidateY, np.float64).ravel()
    fullData = np.vstack([trainX, candidateX])                                   # :26
    weights = np.ones(candidateX.shape[0]) if weights is None else np.asarray(weights, np.float64).copy()
    meanValue = np.zeros(num_iterations)
    standardDeviation = np.zeros(num_iterations)
    posterior = []
    for i in range(num_iterations):
        model = bart(trainX, trainY, ndpost=5000, keepevery=3, ntree=50, k=2.0, seed=int(rng.integers(1 << 31)))  # :41-42
        var = bartint_candidate_var(model, candidateX, weights)                  # :47,50
        index = r_sample_one(np.nonzero(var == var.max())[0] + 1, rng)           # :51
        trainX = np.vstack([trainX, candidateX[index - 1]])                      # :72
        trainY = np.append(trainY, candidateY[index - 1])
        rm = bartint_draw_means(model, fullData)                                 # :74,76
        S = rm.size
        # mean(pred) == mean of the row means (all rows have the same length)       :80
        meanValue[i] = float(rm.mean())
        standardDeviation[i] = float(((rm - meanValue[i]) ** 2).sum() / (S - 1)) if S > 1 else float("nan")  # :82 (variance, B6)
        if save_posterior:
            posterior.append(rm)
        candidateX = np.delete(candidateX, index - 1, axis=0)                    # :85-87
        weights = np.delete(weights, index - 1)
        candidateY = np.delete(candidateY, index - 1)
        model.close()
    return {"meanValueBART": meanValue, "standardDeviationBART": standardDeviation,
            "trainData": np.column_stack([trainX, trainY]), "posterior_samples": posterior}
