"""Forest / Points handles over the C ABI, and the FlatForest (SoA) host container.

Vocabulary follows the reference: a *forest* is a dbarts posterior of S draws x m trees
(`model$fit$state[[1]]@savedTrees`, BARTInt.R:110); *points* are rows of the matrix handed to
`predict(model, X)` (BARTInt.R:312); *cut points* are `dbarts:::createCutPoints` (BARTInt.R:108).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


class BartIntError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"bartint {_lib.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def _check(status: int) -> None:
    if status != 0:
        raise BartIntError(status, _lib.load().bartint_last_error().decode("utf-8", "replace"))


def _p(arr, ctype):
    return None if arr is None else arr.ctypes.data_as(C.POINTER(ctype))


@dataclass
class FlatForest:
    """Flattened posterior (SoA).  Tree (draw s, tree t) is at linear index t + s*m.

    var == -1 marks a leaf; left/right are node indices relative to the tree's first node;
    mu is the leaf value in dbarts' scaled response units.
    """
    S: int
    m: int
    d: int
    tree_offset: np.ndarray
    var: np.ndarray
    cut: np.ndarray
    left: np.ndarray
    right: np.ndarray
    mu: np.ndarray
    cut_offsets: np.ndarray
    cut_values: np.ndarray
    y_min: float = -0.5
    y_max: float = 0.5
    meta: dict = field(default_factory=dict)
    route: int = 0      # 0: dbarts "left iff x <= cut"; 1: BART::wbart "left iff x < cut" (include/bartint.h)

    def __post_init__(self):
        self.tree_offset = np.ascontiguousarray(self.tree_offset, np.int64)
        for k in ("var", "cut", "left", "right", "cut_offsets"):
            setattr(self, k, np.ascontiguousarray(getattr(self, k), np.int32))
        self.mu = np.ascontiguousarray(self.mu, np.float64)
        self.cut_values = np.ascontiguousarray(self.cut_values, np.float64)

    @property
    def n_trees(self) -> int:
        return self.S * self.m

    @property
    def n_nodes(self) -> int:
        return int(self.tree_offset[-1])

    def cut_lists(self):
        return [self.cut_values[self.cut_offsets[j]:self.cut_offsets[j + 1]] for j in range(self.d)]

    def slice_draws(self, lo: int, hi: int) -> "FlatForest":
        """Sub-forest of draws [lo, hi) (used for draw-sharding across GPUs)."""
        a, b = int(self.tree_offset[lo * self.m]), int(self.tree_offset[hi * self.m])
        return FlatForest(hi - lo, self.m, self.d, self.tree_offset[lo * self.m:hi * self.m + 1] - a,
                          self.var[a:b], self.cut[a:b], self.left[a:b], self.right[a:b], self.mu[a:b],
                          self.cut_offsets, self.cut_values, self.y_min, self.y_max, dict(self.meta), self.route)

    # ---- dbarts 0.9-8 wire format, for exporter tests and synthetic "models" ----
    def tree_string(self, k: int) -> str:
        """Pre-order string of tree k: tree := "." | VAR " " CUT " " tree tree (0-based ints)."""
        a = int(self.tree_offset[k])
        out = []
        stack = [0]
        while stack:
            node = stack.pop()
            if self.var[a + node] < 0:
                out.append(".")
            else:
                out.append(f"{int(self.var[a + node])} {int(self.cut[a + node])} ")
                stack.append(int(self.right[a + node]))
                stack.append(int(self.left[a + node]))
        return "".join(out)

    def to_wbart_text(self) -> str:
        """The forest as BART::wbart's ``treedraws$trees`` text: "ndpost ntree p", then per tree the node count and one
        line "nid var cut theta" per node in pre-order (heap ids: root 1, children 2 nid / 2 nid + 1)."""
        out = [f"{self.S} {self.m} {self.d}"]
        for k in range(self.n_trees):
            a, b = int(self.tree_offset[k]), int(self.tree_offset[k + 1])
            out.append(str(b - a))
            stack = [(0, 1)]
            while stack:
                node, nid = stack.pop()
                if self.var[a + node] >= 0:
                    out.append(f"{nid} {int(self.var[a + node])} {int(self.cut[a + node])} 0")
                    stack.append((int(self.right[a + node]), 2 * nid + 1))
                    stack.append((int(self.left[a + node]), 2 * nid))
                else:
                    out.append(f"{nid} 0 0 {float(self.mu[a + node])!r}")
        return "\n".join(out) + "\n"

    def to_dbarts_state(self, x_train: np.ndarray):
        """(savedTrees [m, S] of str, savedTreeFits [n, m, S]) as dbarts would hold them for this forest:
        the fit of training row i under tree (t, s) is the mu of the leaf the row falls in."""
        x_train = np.asarray(x_train, np.float64)
        n = x_train.shape[0]
        K = self.n_trees
        strings = np.empty((self.m, self.S), dtype=object)
        for k in range(K):
            strings[k % self.m, k // self.m] = self.tree_string(k)
        # route every training row through every tree at once on the integer cut map (x > cut  <=>  code > cut index)
        codes = np.zeros((n, max(self.d, 1)), np.int32)
        for j, cuts in enumerate(self.cut_lists()):
            codes[:, j] = np.where(np.isnan(x_train[:, j]), 0, np.searchsorted(cuts, x_train[:, j], side="left"))
        base = self.tree_offset[:-1, None]
        node = np.broadcast_to(base, (K, n)).copy()
        rows = np.arange(n)[None, :]
        while True:
            v = self.var[node]
            internal = v >= 0
            if not internal.any():
                break
            go_right = codes[rows, np.where(internal, v, 0)] > self.cut[node]
            nxt = np.where(go_right, self.right[node], self.left[node]) + base
            node = np.where(internal, nxt, node)
        # [n, m, S] in R's column-major layout (n fastest) == C-ordered (S, m, n): a transposed view, no copy
        fits = self.mu[node].reshape(self.S, self.m, n).transpose(2, 1, 0)
        return strings, fits


class WireStrings:
    """savedTrees as the C ABI takes them: an array of K = m*S C strings in R's column-major order (tree index
    fastest).  R already holds its character matrix this way (CHARSXP pointers), so the .Call shim passes them
    without conversion; on the Python side the encoding is done once per model and cached here."""

    def __init__(self, tree_strings):
        tree_strings = np.asarray(tree_strings, dtype=object)
        self.m, self.S = tree_strings.shape
        self._enc = [str(v).encode("ascii") for v in tree_strings.T.ravel()]     # index t + s*m
        self.array = (C.c_char_p * max(len(self._enc), 1))(*self._enc)


    def slice_draws(self, lo: int, hi: int) -> "WireStrings":
        """The strings of draws [lo, hi) as a view (no re-encoding): what a rank passes when it exports its share."""
        v = WireStrings.__new__(WireStrings)
        v.m, v.S = self.m, hi - lo
        v._enc = self._enc                                   # keeps the byte strings alive
        n = max((hi - lo) * self.m, 1)
        v.array = (C.c_char_p * n).from_address(C.addressof(self.array) + lo * self.m * C.sizeof(C.c_char_p))
        v._parent = self
        return v


def topk(values, k: int):
    """order(values, decreasing=TRUE)[1:k] on the device (0-based indices; ties in index order, NaN never selected):
    the batch version of the reference's `which(var == max(var))` pick (BARTInt.R:315,387)."""
    v = np.ascontiguousarray(np.asarray(values, np.float64).ravel())
    idx = np.zeros(max(k, 1), np.int64)
    val = np.zeros(max(k, 1))
    _check(_lib.load().bartint_topk(_p(v, C.c_double), v.size, int(k), _p(idx, C.c_int64), _p(val, C.c_double)))
    return idx[:k], val[:k]


def init_group(n_gpus: int = 0) -> int:
    """bartint_init: open the single-process device group over n_gpus devices (0 = all).  Returns its size."""
    _check(_lib.load().bartint_init(int(n_gpus)))
    return int(_lib.load().bartint_group_size())


def group_size() -> int:
    return int(_lib.load().bartint_group_size())


def group_exchange() -> str:
    return _lib.load().bartint_group_exchange().decode()


def shutdown_group() -> None:
    _lib.load().bartint_shutdown()


class StagedMatrix:
    """A host matrix on its way to the device (asynchronous copy started at construction), independent of any forest:
    the candidate matrix / fullData exist before the model is fitted, so their transfer can overlap the export."""

    def __init__(self, X, device: int = 0):
        """device: a CUDA device index, or "group" / -1 for the device group opened by `init_group` (row shards, one
        per member)."""
        if device == "group":
            device = _lib.BARTINT_DEVICE_GROUP
        X = np.asarray(X, np.float64)
        self._X = np.asfortranarray(X)          # keep the host buffer alive until the copy has been consumed
        self.N, self.d = self._X.shape
        h = C.c_void_p()
        _check(_lib.load().bartint_stage_matrix(_p(self._X, C.c_double), self.N, self.d, device, C.byref(h)))
        self._h = h

    def flush(self) -> None:
        """Put the remaining row blocks on the bus now (by default only the first block travels until a consumer asks)."""
        _check(_lib.load().bartint_staged_flush(self._h))

    def points(self, forest: "Forest", stream: int = 0) -> "Points":
        h = C.c_void_p()
        _check(_lib.load().bartint_staged_points(forest._h, self._h, C.c_void_p(stream), C.byref(h)))
        return Points(forest, h.value, self.N)

    def close(self):
        if self._h:
            _lib.load().bartint_staged_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Points:
    """A point set binned to cut-point codes, resident in HBM (dbarts' integer cut map of x.test)."""

    def __init__(self, forest: "Forest", handle: int, N: int):
        self._forest = forest
        self._h = C.c_void_p(handle)
        self.N = N

    def codes(self) -> np.ndarray:
        """(N, d) uint8 codes: #{k : x > cut[j][k]} (for the bit-exactness tests)."""
        out = np.zeros((self._forest.d, self.N), np.uint8)
        _check(_lib.load().bartint_points_codes(self._h, _p(out, C.c_uint8)))
        return out.T.copy()

    def close(self):
        if self._h:
            _lib.load().bartint_points_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Forest:
    """A flattened posterior resident in HBM."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        info = (C.c_int64 * 12)()
        _check(_lib.load().bartint_forest_info(self._h, info))
        (self.S, self.m, self.d, self.n_nodes, self.n_leaves, self.max_depth, self.max_internal, tk, self.n_groups,
         self.table_bits, kind, self.gather_words) = [int(v) for v in info]
        self.table_kernel_ok = bool(tk)                                  # a fast kernel (table or gather) applies
        self.fast_kernel = {0: "none", 1: "table", 2: "gather"}[kind]
        self.route = int(_lib.load().bartint_forest_route(self._h))
        bs = (C.c_int64 * 16)()
        _check(_lib.load().bartint_forest_bitslice_info(self._h, bs))
        # the bit-sliced sums-only kernel: applies, predicate rows, points per tile, longest draw, fixed-point exponent
        self.bitslice_ok, self.bitslice_rows, self.bitslice_tile = bool(bs[0]), int(bs[1]), int(bs[2]) * 128
        self.bitslice_info = [int(v) for v in bs]
        self.y_min = self.y_max = None

    # ---- constructors ----
    @classmethod
    def from_soa(cls, ff: FlatForest) -> "Forest":
        lib = _lib.load()
        h = C.c_void_p()
        _check(lib.bartint_forest_from_soa_routed(
            ff.S, ff.m, ff.d, _p(ff.tree_offset, C.c_int64), _p(ff.var, C.c_int32), _p(ff.cut, C.c_int32),
            _p(ff.left, C.c_int32), _p(ff.right, C.c_int32), _p(ff.mu, C.c_double), _p(ff.cut_offsets, C.c_int32),
            _p(ff.cut_values, C.c_double), float(ff.y_min), float(ff.y_max), int(getattr(ff, "route", 0)), C.byref(h)))
        f = cls(h.value)
        f.y_min, f.y_max = float(ff.y_min), float(ff.y_max)
        return f

    @classmethod
    def from_wbart(cls, trees_text, cut_lists, offset: float = 0.0) -> "Forest":
        """Exporter for BART::wbart / pbart fits: ``fit$treedraws$trees`` (text), ``fit$treedraws$cutpoints`` and
        ``fit$offset``.  Routing follows tree::bn (left iff x < cut); predictions are offset + sum of leaf values."""
        lib = _lib.load()
        text = trees_text.encode() if isinstance(trees_text, str) else bytes(trees_text)
        d = len(cut_lists)
        cut_offsets = np.zeros(d + 1, np.int32)
        cut_offsets[1:] = np.cumsum([len(c) for c in cut_lists])
        cut_values = (np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0))
        cut_values = np.ascontiguousarray(cut_values)
        h = C.c_void_p()
        _check(lib.bartint_forest_from_wbart(text, len(text), d, _p(cut_offsets, C.c_int32), _p(cut_values, C.c_double),
                                             float(offset), C.byref(h)))
        f = cls(h.value)
        f.y_min, f.y_max = float(offset) - 0.5, float(offset) + 0.5
        return f

    @classmethod
    def from_value_splits(cls, S, m, d, tree_offset, var, value, left=None, right=None, route=0, leaf_scale=1.0,
                          offset=0.0) -> "Forest":
        """Exporter for trees given with split VALUES (dbarts >= 0.9-9 ``fit$getTrees()``, bartMachine raw node data):
        var[g] = -1 marks a leaf, value[g] is the split value or the leaf prediction; left/right None = pre-order."""
        toff = np.ascontiguousarray(tree_offset, np.int64)
        var = np.ascontiguousarray(var, np.int32)
        value = np.ascontiguousarray(value, np.float64)
        lft = None if left is None else np.ascontiguousarray(left, np.int32)
        rgt = None if right is None else np.ascontiguousarray(right, np.int32)
        h = C.c_void_p()
        _check(_lib.load().bartint_forest_from_value_splits(int(S), int(m), int(d), _p(toff, C.c_int64), _p(var, C.c_int32),
                                                            _p(value, C.c_double), _p(lft, C.c_int32), _p(rgt, C.c_int32),
                                                            int(route), float(leaf_scale), float(offset), C.byref(h)))
        f = cls(h.value)
        f.y_min, f.y_max = float(offset) - 0.5, float(offset) + 0.5
        return f

    def cutpoints(self):
        """The forest's cut-point lists (one ascending array per variable)."""
        off = np.zeros(self.d + 1, np.int32)
        _check(_lib.load().bartint_forest_cutpoints(self._h, _p(off, C.c_int32), None))
        val = np.zeros(max(int(off[-1]), 1))
        _check(_lib.load().bartint_forest_cutpoints(self._h, _p(off, C.c_int32), _p(val, C.c_double)))
        return [val[off[j]:off[j + 1]].copy() for j in range(self.d)]

    @classmethod
    def from_dbarts(cls, tree_strings, tree_fits, x_train, cut_lists, y_min, y_max) -> "Forest":
        """Exporter: dbarts 0.9-8 saved state -> HBM forest (getTree, BARTInt.R:92-133, for all trees).

        tree_strings [m, S] (R: savedTrees[treeNum, sampleNum]); tree_fits [n, m, S]; x_train [n, d].
        """
        lib = _lib.load()
        wire = tree_strings if isinstance(tree_strings, WireStrings) else WireStrings(tree_strings)
        m, S, arr = wire.m, wire.S, wire.array
        x_train = np.asarray(x_train, np.float64)
        n, d = x_train.shape
        fits = np.asarray(tree_fits, np.float64).reshape(n, m, S)
        if not fits.flags["F_CONTIGUOUS"]:
            fits = np.asfortranarray(fits)          # R arrays are column-major already; numpy ones may need a copy
        xt = np.asfortranarray(x_train)
        cut_offsets = np.zeros(d + 1, np.int32)
        cut_offsets[1:] = np.cumsum([len(c) for c in cut_lists])
        cut_values = (np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0))
        cut_values = np.ascontiguousarray(cut_values)
        h = C.c_void_p()
        _check(lib.bartint_forest_from_dbarts098(arr, _p(fits, C.c_double), _p(xt, C.c_double), n, d,
                                                 _p(cut_offsets, C.c_int32), _p(cut_values, C.c_double),
                                                 float(y_min), float(y_max), m, S, C.byref(h)))
        f = cls(h.value)
        f.y_min, f.y_max = float(y_min), float(y_max)
        f._cut_offsets, f._cut_values = cut_offsets, cut_values
        return f

    # ---- device image (replication between ranks) ----
    def image_size(self) -> tuple[int, int]:
        """(host header bytes, device blob bytes) of this forest's device image."""
        hb, db = C.c_int64(0), C.c_int64(0)
        _check(_lib.load().bartint_forest_image_size(self._h, C.byref(hb), C.byref(db)))
        return int(hb.value), int(db.value)

    def pack(self, d_blob: int, device_bytes: int, stream: int = 0) -> np.ndarray:
        """Copy the device arrays into the device buffer at address d_blob (asynchronously on `stream`) and return the
        host header (uint8 array).  Together they rebuild the forest elsewhere with Forest.unpack."""
        hb, _ = self.image_size()
        header = np.zeros(hb, np.uint8)
        _check(_lib.load().bartint_forest_pack(self._h, header.ctypes.data_as(C.c_void_p), hb, C.c_void_p(d_blob),
                                               int(device_bytes), C.c_void_p(stream)))
        return header

    @classmethod
    def unpack(cls, header: np.ndarray, d_blob: int, device_bytes: int, stream: int = 0) -> "Forest":
        """Forest on the current device from a (header, device blob) image made by pack() -- possibly on another rank."""
        header = np.ascontiguousarray(header, np.uint8)
        h = C.c_void_p()
        _check(_lib.load().bartint_forest_unpack(header.ctypes.data_as(C.c_void_p), header.nbytes, C.c_void_p(d_blob),
                                                 int(device_bytes), C.c_void_p(stream), C.byref(h)))
        return cls(h.value)

    def export_soa(self) -> dict:
        n = self.n_nodes
        off = np.zeros(self.S * self.m + 1, np.int64)
        var, cut, left, right = (np.zeros(n, np.int32) for _ in range(4))
        mu = np.zeros(n)
        _check(_lib.load().bartint_forest_export_soa(self._h, _p(off, C.c_int64), _p(var, C.c_int32),
                                                     _p(cut, C.c_int32), _p(left, C.c_int32), _p(right, C.c_int32),
                                                     _p(mu, C.c_double)))
        return dict(tree_offset=off, var=var, cut=cut, left=left, right=right, mu=mu)

    # ---- points ----
    def points(self, X) -> Points:
        """Bin a host matrix (N x d, any layout) on the device."""
        X = np.asarray(X, np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, self.d) if self.d else X.reshape(-1, 0)
        N, d = X.shape
        Xf = np.asfortranarray(X)
        h = C.c_void_p()
        _check(_lib.load().bartint_points_from_host(self._h, _p(Xf, C.c_double), N, d, C.byref(h)))
        return Points(self, h.value, N)

    def points_from_device(self, dptr: int, N: int, ld: int | None = None, stream: int = 0) -> Points:
        """Bin a column-major N x d fp64 matrix that already lives in HBM (dptr = device address)."""
        h = C.c_void_p()
        _check(_lib.load().bartint_points_from_device(self._h, C.c_void_p(dptr), N, self.d, ld or N,
                                                      C.c_void_p(stream), C.byref(h)))
        return Points(self, h.value, N)

    def generate_points(self, measure: str, N: int, seed: int, first_index: int = 0, d_x_out: int = 0,
                        stream: int = 0) -> Points:
        """Draw N points from the integration measure on the device (Philox4x32-10) and bin them on the fly."""
        h = C.c_void_p()
        _check(_lib.load().bartint_points_generate(self._h, _lib.MEASURES[measure], N, seed, first_index,
                                                   C.c_void_p(d_x_out or None), C.c_void_p(stream), C.byref(h)))
        return Points(self, h.value, N)

    # ---- evaluation ----
    def eval(self, pts, weights=None, want=("draw_mean", "point_mean", "point_var"), scaled=False, draws=None,
             force_walk=False, no_bitslice=False) -> dict:
        """predict + colVars*weights + rowMeans (+ the full S x N matrix when asked), fused on the device.

        `pts` is either a Points handle (codes resident in HBM) or a host matrix (N x d): the latter goes through the
        one-shot `bartint_eval`, which pipelines host->device copies against the evaluation for large N."""
        lo, hi = (0, self.S) if draws is None else draws
        nd = hi - lo
        host = not isinstance(pts, Points)
        if host:
            X = np.asarray(pts, np.float64)
            if X.ndim == 1:
                X = X.reshape(-1, self.d) if self.d else X.reshape(-1, 0)
            X = np.asfortranarray(X)
            N = X.shape[0]
        else:
            N = pts.N
        out = {}
        if "draw_mean" in want:
            out["draw_mean"] = np.zeros(nd)
        if "point_mean" in want:
            out["point_mean"] = np.zeros(N)
        if "point_var" in want:
            out["point_var"] = np.zeros(N)
        if "full_pred" in want:
            out["full_pred"] = np.zeros((nd, N), order="F")
        w = None
        if weights is not None:
            w = np.ascontiguousarray(np.asarray(weights, np.float64).ravel())
        flags = ((_lib.BARTINT_SCALED_UNITS if scaled else 0) | (_lib.BARTINT_FORCE_WALK if force_walk else 0) |
                 (_lib.BARTINT_NO_BITSLICE if no_bitslice else 0))
        outs = (_p(out.get("draw_mean"), C.c_double), _p(out.get("point_mean"), C.c_double),
                _p(out.get("point_var"), C.c_double), _p(out.get("full_pred"), C.c_double))
        wargs = (_p(w, C.c_double), 0 if w is None else w.size)
        if host and draws is None:
            _check(_lib.load().bartint_eval(self._h, _p(X, C.c_double), N, X.shape[1], flags, *wargs, *outs))
            return out
        own = host
        if own:
            pts = self.points(X)
        try:
            _check(_lib.load().bartint_eval_points(self._h, pts._h, flags, lo, hi, *wargs, *outs))
            return out
        finally:
            if own:
                pts.close()

    def eval_device(self, pts: Points, d_draw_sum: int = 0, d_point_mean: int = 0, d_point_m2: int = 0, draws=None,
                    stream: int = 0, force_walk: bool = False, no_bitslice: bool = False) -> None:
        """Asynchronous evaluation writing raw scaled-unit partials to device addresses (0 = not wanted)."""
        lo, hi = (0, self.S) if draws is None else draws
        flags = (_lib.BARTINT_FORCE_WALK if force_walk else 0) | (_lib.BARTINT_NO_BITSLICE if no_bitslice else 0)
        _check(_lib.load().bartint_eval_points_device(
            self._h, pts._h, flags, lo, hi, C.c_void_p(d_draw_sum or None),
            C.c_void_p(d_point_mean or None), C.c_void_p(d_point_m2 or None), C.c_void_p(stream)))

    def merge_moments_device(self, shard_draws, d_mean: int, d_m2: int, N: int, d_weights: int, weights_len: int,
                             d_mean_out: int, d_var_out: int, scaled=False, stream: int = 0) -> None:
        sd = np.ascontiguousarray(shard_draws, np.int32)
        _check(_lib.load().bartint_merge_moments_device(
            self._h, _lib.BARTINT_SCALED_UNITS if scaled else 0, sd.size, _p(sd, C.c_int32), C.c_void_p(d_mean),
            C.c_void_p(d_m2), N, C.c_void_p(d_weights or None), weights_len, C.c_void_p(d_mean_out or None),
            C.c_void_p(d_var_out or None), C.c_void_p(stream)))

    def finalize_draws_device(self, d_draw_sum: int, ndraws: int, n_points: int, d_draw_mean: int, d_mean_var: int = 0,
                              scaled=False, stream: int = 0) -> None:
        """rowMeans + {mean(pred), variance of row means} (bartMean.R:76,80,82) from per-draw sums, on device."""
        _check(_lib.load().bartint_finalize_draws_device(
            self._h, _lib.BARTINT_SCALED_UNITS if scaled else 0, C.c_void_p(d_draw_sum), ndraws, n_points,
            C.c_void_p(d_draw_mean), C.c_void_p(d_mean_var or None), C.c_void_p(stream)))

    def exact_integrals_device(self, d_out: int, measure="uniform", n_draws_used=None, d_rescaled: int = 0,
                               d_mean_sd: int = 0, stream: int = 0) -> None:
        """sampleIntegrals (+ the BARTInt.R:260-262 finalisation when d_mean_sd is given), asynchronous."""
        _check(_lib.load().bartint_exact_integrals_device(
            self._h, _lib.MEASURES[measure], None, None, int(n_draws_used or 0), C.c_void_p(d_out),
            C.c_void_p(d_rescaled or None), C.c_void_p(d_mean_sd or None), C.c_void_p(stream)))

    def last_eval_profile(self, with_moments: bool = False) -> dict:
        """Timing of the last evaluation of the given kind (per-draw sums only / with per-point outputs)."""
        ms, nl, tk = C.c_float(), C.c_int32(), C.c_int32()
        _check(_lib.load().bartint_last_eval_profile(self._h, int(with_moments), C.byref(ms), C.byref(nl), C.byref(tk)))
        return {"kernel_ms": ms.value, "launches": nl.value, "table_kernel": tk.value == 1,
                "kernel": {0: "walk", 1: self.fast_kernel, 2: "bitslice"}.get(tk.value, "?")}

    def leaf_indices(self, pts, draws=None, use_table_kernel=False) -> np.ndarray:
        """Leaf ordinals, shape (ndraws*m, N), row t + (s - lo)*m."""
        own = not isinstance(pts, Points)
        if own:
            pts = self.points(pts)
        try:
            lo, hi = (0, self.S) if draws is None else draws
            out = np.zeros(((hi - lo) * self.m, pts.N), np.int32)
            _check(_lib.load().bartint_leaf_indices(self._h, pts._h, lo, hi, int(use_table_kernel),
                                                    _p(out, C.c_int32)))
            return out
        finally:
            if own:
                pts.close()

    # ---- exact integral ----
    def exact_integrals(self, measure="uniform", n_draws_used=None, lo=None, hi=None) -> np.ndarray:
        n = self.S if not n_draws_used or n_draws_used > self.S else int(n_draws_used)
        out = np.zeros(n)
        lo_a = None if lo is None else np.ascontiguousarray(np.broadcast_to(np.asarray(lo, np.float64), (self.d,)))
        hi_a = None if hi is None else np.ascontiguousarray(np.broadcast_to(np.asarray(hi, np.float64), (self.d,)))
        _check(_lib.load().bartint_exact_integrals(self._h, _lib.MEASURES[measure], _p(lo_a, C.c_double),
                                                   _p(hi_a, C.c_double), n, _p(out, C.c_double)))
        return out

    def finalize_integrals_device(self, d_integrals: int, n: int, d_rescaled: int, d_mean_sd: int, stream: int = 0) -> None:
        """BARTInt.R:260-262 on device buffers (asynchronous): rescale, mean, sd."""
        _check(_lib.load().bartint_finalize_integrals_device(self._h, C.c_void_p(d_integrals), int(n),
                                                             C.c_void_p(d_rescaled or None), C.c_void_p(d_mean_sd),
                                                             C.c_void_p(stream)))

    def finalize_integrals(self, integrals_scaled):
        x = np.ascontiguousarray(integrals_scaled, np.float64)
        res = np.zeros_like(x)
        mean, sd = C.c_double(), C.c_double()
        _check(_lib.load().bartint_finalize_integrals(self._h, _p(x, C.c_double), x.size, _p(res, C.c_double),
                                                      C.byref(mean), C.byref(sd)))
        return res, mean.value, sd.value

    def close(self):
        if self._h:
            _lib.load().bartint_forest_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def design_step_dbarts(tree_strings, tree_fits, x_train, cut_lists, y_min, y_max, mc=None, candidates=None,
                       weights=None, measure="uniform", n_draws_integral=None, n_chunks=0,
                       want=("integrals", "integral_mean_sd", "cand_var", "draw_mean", "draw_mean_var")) -> dict:
    """One sequential-design step in one ABI call (bartint_design_step_dbarts098): exact integrals + their mean/sd
    (BARTInt.R:259-262), per-candidate variance (BARTInt.R:312-314) and per-draw means over the MC / population points
    (bartMean.R:74-82), with the exporter pipelined against the evaluation kernels over chunks of draws.

    mc / candidates: StagedMatrix (start their host->device copies before calling) or None.
    """
    lib = _lib.load()
    wire = tree_strings if isinstance(tree_strings, WireStrings) else WireStrings(tree_strings)
    m, S = wire.m, wire.S
    x_train = np.asarray(x_train, np.float64)
    n, d = x_train.shape
    fits = np.asarray(tree_fits, np.float64).reshape(n, m, S)
    if not fits.flags["F_CONTIGUOUS"]:
        fits = np.asfortranarray(fits)
    xt = np.asfortranarray(x_train)
    cut_offsets = np.zeros(d + 1, np.int32)
    cut_offsets[1:] = np.cumsum([len(c) for c in cut_lists])
    cut_values = np.ascontiguousarray(np.concatenate([np.asarray(c, np.float64) for c in cut_lists]) if d else np.zeros(0))
    n_int = S if not n_draws_integral else min(int(n_draws_integral), S)
    C_ = candidates.N if candidates is not None else 0
    out = {}
    if "integrals" in want:
        out["integrals"] = np.zeros(n_int)
    if "integral_mean_sd" in want:
        out["integral_mean_sd"] = np.zeros(2)
    if "cand_var" in want and candidates is not None:
        out["cand_var"] = np.zeros(C_)
    if "draw_mean" in want and mc is not None:
        out["draw_mean"] = np.zeros(S)
    if "draw_mean_var" in want and mc is not None:
        out["draw_mean_var"] = np.zeros(2)
    w = None if weights is None else np.ascontiguousarray(np.atleast_1d(np.asarray(weights, np.float64)))
    _check(lib.bartint_design_step_dbarts098(
        wire.array, _p(fits, C.c_double), _p(xt, C.c_double), n, d, _p(cut_offsets, C.c_int32), _p(cut_values, C.c_double),
        float(y_min), float(y_max), m, S, mc._h if mc is not None else None,
        candidates._h if candidates is not None else None, _p(w, C.c_double) if w is not None else None,
        0 if w is None else len(w), _lib.MEASURES[measure], 0 if not n_draws_integral else int(n_draws_integral),
        int(n_chunks), *(_p(out[k], C.c_double) if k in out else None
                         for k in ("integrals", "integral_mean_sd", "cand_var", "draw_mean", "draw_mean_var"))))
    return out
