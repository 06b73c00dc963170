"""The R-side binding (bart-int_b200/R/r_shim.c) compiled against include/bartint.h and driven through a minimal mock
of R's C API (tests/r_mock: REAL / INTEGER / STRING_ELT / dim attributes / external pointers / Rf_error as a longjmp).
R itself is absent from this image, so this is as close as the drop-in boundary gets to being executed here:

* without a GPU the shim runs up to the library's device check -- which comes AFTER the exporter has parsed every
  tree string and validated it against the cut-point lists, so the marshalling of savedTrees (column-major
  [tree, sample]), the cut-point list and the dims is exercised, and library errors arrive as R errors;
* with a GPU (BARTINT_TEST_RSHIM_GPU=1, see tools/verify_on_gpu.sh) the complete calls are compared with the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bartint_b200 import synth
from bartint_testlib import has_gpu
from oracle import oracle_np as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "r_mock")
SEXP = C.c_void_p


@pytest.fixture(scope="module")
def rmock():
    out = os.path.join(MOCK, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbartint_rmock.so")
    libdir = os.path.join(ROOT, "bart-int_b200")
    # -Werror: a shim that no longer matches include/bartint.h (argument count / pointer types) must not compile
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-O1", "-I", MOCK, "-I",
                    os.path.join(ROOT, "include"), "-o", so, os.path.join(MOCK, "r_mock.c"),
                    os.path.join(libdir, "R", "r_shim.c"), "-L", libdir, "-lbartint_cuda", f"-Wl,-rpath,{libdir}"],
                   check=True)
    lib = C.CDLL(so)
    for name, res, args in [
            ("mock_nil", SEXP, []), ("mock_real", SEXP, [C.POINTER(C.c_double), C.c_ssize_t]),
            ("mock_int", SEXP, [C.POINTER(C.c_int), C.c_ssize_t]), ("mock_set_dim", None, [SEXP, C.POINTER(C.c_int), C.c_int]),
            ("mock_strings", SEXP, [C.POINTER(C.c_char_p), C.c_ssize_t]), ("mock_list", SEXP, [C.c_ssize_t]),
            ("SET_VECTOR_ELT", SEXP, [SEXP, C.c_ssize_t, SEXP]), ("VECTOR_ELT", SEXP, [SEXP, C.c_ssize_t]),
            ("mock_type", C.c_uint, [SEXP]), ("mock_length", C.c_ssize_t, [SEXP]),
            ("mock_real_ptr", C.POINTER(C.c_double), [SEXP]), ("mock_dim_ptr", C.POINTER(C.c_int), [SEXP]),
            ("mock_last_error", C.c_char_p, []), ("mock_finalize", None, [SEXP]), ("mock_reset", None, []),
            ("mock_call", SEXP, [C.c_void_p, C.c_int, C.POINTER(SEXP)])]:
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    yield RMock(lib)
    lib.mock_reset()


class RMock:
    def __init__(self, lib):
        self.lib = lib

    def real(self, a, dim=True):
        a = np.asarray(a, np.float64)
        flat = np.ascontiguousarray(a.ravel(order="F"))           # R is column-major
        s = self.lib.mock_real(flat.ctypes.data_as(C.POINTER(C.c_double)), flat.size)
        if dim and a.ndim >= 2:
            dims = (C.c_int * a.ndim)(*a.shape)
            self.lib.mock_set_dim(s, dims, a.ndim)
        return s

    def integer(self, v):
        arr = (C.c_int * 1)(int(v))
        return self.lib.mock_int(arr, 1)

    def strings(self, mat):
        mat = np.asarray(mat, dtype=object)
        flat = [str(x).encode() for x in mat.ravel(order="F")]
        arr = (C.c_char_p * len(flat))(*flat)
        s = self.lib.mock_strings(arr, len(flat))
        if mat.ndim == 2:
            self.lib.mock_set_dim(s, (C.c_int * 2)(*mat.shape), 2)
        return s

    def list_of_reals(self, vectors):
        s = self.lib.mock_list(len(vectors))
        for i, v in enumerate(vectors):
            self.lib.SET_VECTOR_ELT(s, i, self.real(np.asarray(v, float), dim=False))
        return s

    def call(self, name, *args):
        """.Call(name, ...): the value, or raises RuntimeError carrying R's error message."""
        fn = C.cast(getattr(self.lib, name), C.c_void_p)
        arr = (SEXP * len(args))(*args)
        out = self.lib.mock_call(fn, len(args), arr)
        err = self.lib.mock_last_error().decode()
        if err:
            raise RuntimeError(err)
        return out

    def to_numpy(self, s):
        n = self.lib.mock_length(s)
        a = np.ctypeslib.as_array(self.lib.mock_real_ptr(s), shape=(n,)).copy() if n else np.zeros(0)
        d = self.lib.mock_dim_ptr(s)
        return a.reshape((d[0], d[1]), order="F") if d else a


def _state(seed=3, S=5, m=4, d=3, n=30):
    rng = np.random.default_rng(seed)
    x_train = rng.random((n, d))
    ff = synth.prior_forest(seed, S, m, x_train, synth.genz_gaussian(x_train))
    strings, fits = ff.to_dbarts_state(x_train)
    return ff, x_train, strings, np.asarray(fits)


def _forest_args(r, ff, x_train, strings, fits):
    return (r.strings(strings), r.real(fits), r.real(x_train), r.list_of_reals(ff.cut_lists()),
            r.real([ff.y_min], dim=False), r.real([ff.y_max], dim=False))


def test_shim_marshals_the_dbarts_state_up_to_the_device_check(rmock):
    ff, x_train, strings, fits = _state()
    if has_gpu():
        h = rmock.call("bartint_R_forest", *_forest_args(rmock, ff, x_train, strings, fits))
        assert rmock.lib.mock_type(h) == 22                      # EXTPTRSXP
        rmock.lib.mock_finalize(h)                               # what R's gc does: bartint_forest_destroy
        with pytest.raises(RuntimeError, match="already released"):
            rmock.call("bartint_R_exact_integrals", h, rmock.integer(0), rmock.integer(0))
    else:
        # every string parsed, every split checked against the cut lists, leaf values recovered -- then no device
        with pytest.raises(RuntimeError, match="^bartint: .*(device|CUDA)"):
            rmock.call("bartint_R_forest", *_forest_args(rmock, ff, x_train, strings, fits))


def test_library_errors_become_r_errors_with_the_right_coordinates(rmock):
    ff, x_train, strings, fits = _state()
    bad = strings.copy()
    bad[1, 3] = "0 1 . ."[:-1] + "x"                              # savedTrees[tree 2, sample 4] in R's 1-based terms
    with pytest.raises(RuntimeError, match=r"malformed dbarts tree string at tree 1 \(draw 3\)"):
        rmock.call("bartint_R_forest", *_forest_args(rmock, ff, x_train, bad, fits))
    # a split on the last variable with a cut index one past ITS list (the other lists are longer): the flattened
    # cut-point list has to be marshalled per variable for this to be caught
    cuts = [np.linspace(0, 1, 12)[1:-1], np.linspace(0, 1, 12)[1:-1], np.array([0.5])]
    ok, bad = strings.copy(), strings.copy()
    ok[:, :] = "2 0 . ."
    bad[:, :] = "2 0 . ."
    bad[0, 0] = "2 1 . ."
    args = lambda st: (rmock.strings(st), rmock.real(fits), rmock.real(x_train), rmock.list_of_reals(cuts),
                       rmock.real([0.0], dim=False), rmock.real([1.0], dim=False))
    with pytest.raises(RuntimeError, match="outside the cut-point table"):
        rmock.call("bartint_R_forest", *args(bad))
    if not has_gpu():
        with pytest.raises(RuntimeError, match="(device|CUDA)"):
            rmock.call("bartint_R_forest", *args(ok))


def test_wbart_binding(rmock):
    rng = np.random.default_rng(4)
    ff = synth.prior_forest(4, 3, 5, rng.random((30, 2)), None)
    text = ff.to_wbart_text()
    cuts = rmock.list_of_reals(ff.cut_lists())
    nil = rmock.lib.mock_nil()
    trees = lambda t: rmock.strings(np.array([t], dtype=object))
    with pytest.raises(RuntimeError, match="node count|treedraws|tree 0"):
        rmock.call("bartint_R_forest_wbart", trees("1 1 2\nx\n"), cuts, rmock.real([0.0], dim=False), nil, nil)
    with pytest.raises(RuntimeError, match="p = 2 but d = 1"):
        rmock.call("bartint_R_forest_wbart", trees(text), rmock.list_of_reals(ff.cut_lists()[:1]),
                   rmock.real([0.0], dim=False), nil, nil)
    # a plain wbart fit has $mu and no $offset (gbart: $offset, pbart: $binaryOffset); none of them is an error, never 0
    with pytest.raises(RuntimeError, match="none of .offset, .mu, .binaryOffset"):
        rmock.call("bartint_R_forest_wbart", trees(text), cuts, nil, nil, nil)
    for args in ((rmock.real([1.5], dim=False), nil, nil), (nil, rmock.real([1.5], dim=False), nil),
                 (nil, nil, rmock.real([1.5], dim=False))):
        if has_gpu():
            h = rmock.call("bartint_R_forest_wbart", trees(text), cuts, *args)
            X = rng.random((40, 2))
            out = rmock.call("bartint_R_eval", h, rmock.real(X), nil, rmock.integer(4))
            full = rmock.to_numpy(rmock.lib.VECTOR_ELT(out, 2))
            want = onp.predict(onp.forest_from_wbart(text, ff.cut_lists(), 1.5), X)        # offset + sum of leaf values
            np.testing.assert_allclose(full, want, rtol=0, atol=1e-12)
            rmock.lib.mock_finalize(h)
        else:
            with pytest.raises(RuntimeError, match="(device|CUDA)"):
                rmock.call("bartint_R_forest_wbart", trees(text), cuts, *args)


def test_value_split_binding(rmock):
    """bartint_R_forest_values: the marshalling of a getTrees()-style export (double offsets, 0-based var, pre-order)."""
    off = rmock.real([0.0, 3.0, 4.0], dim=False)
    var = np.array([0, -1, -1, -1], np.int32)
    val = rmock.real([0.4, -1.0, 2.0, 0.25], dim=False)
    ivec = lambda a: rmock.lib.mock_int(np.ascontiguousarray(a, np.int32).ctypes.data_as(C.POINTER(C.c_int)), len(a))
    nil = rmock.lib.mock_nil()
    args = lambda S, v: (rmock.integer(S), rmock.integer(2 // S), rmock.integer(1), off, ivec(v), val, nil, nil, rmock.integer(0),
                         rmock.real([1.0], dim=False), rmock.real([10.0], dim=False))
    with pytest.raises(RuntimeError, match="do not describe S . m trees"):
        rmock.call("bartint_R_forest_values", rmock.integer(3), rmock.integer(1), rmock.integer(1), off, ivec(var), val, nil, nil,
                   rmock.integer(0), rmock.real([1.0], dim=False), rmock.real([0.0], dim=False))
    with pytest.raises(RuntimeError, match="variable 3"):
        rmock.call("bartint_R_forest_values", *args(1, np.array([3, -1, -1, -1], np.int32)))
    if has_gpu():
        h = rmock.call("bartint_R_forest_values", *args(1, var))
        X = np.array([[0.4], [0.41], [0.0]])
        out = rmock.call("bartint_R_eval", h, rmock.real(X), nil, rmock.integer(4))
        full = rmock.to_numpy(rmock.lib.VECTOR_ELT(out, 2))
        np.testing.assert_allclose(full, [[10.0 - 1.0 + 0.25, 10.0 + 2.0 + 0.25, 10.0 - 1.0 + 0.25]], atol=1e-13)   # left iff x <= 0.4
        k = rmock.call("bartint_R_topk", rmock.real([0.1, 0.7, 0.7, np.nan], dim=False), rmock.integer(3))
        assert list(np.ctypeslib.as_array(C.cast(rmock.lib.mock_real_ptr(k), C.POINTER(C.c_int)), shape=(3,))) == [1, 2, 0]
        rmock.call("bartint_R_release", h)
        with pytest.raises(RuntimeError, match="already released"):
            rmock.call("bartint_R_eval", h, rmock.real(X), nil, rmock.integer(4))
        assert rmock.lib.mock_type(rmock.call("bartint_R_init", rmock.integer(1))) == 13          # INTSXP: the group size
    else:
        with pytest.raises(RuntimeError, match="(device|CUDA)"):
            rmock.call("bartint_R_forest_values", *args(1, var))
        with pytest.raises(RuntimeError, match="(device|CUDA)"):
            rmock.call("bartint_R_init", rmock.integer(0))


@pytest.mark.gpu
def test_shim_end_to_end_against_the_oracle(rmock):
    ff, x_train, strings, fits = _state(seed=8, S=12, m=9, d=4, n=60)
    rng = np.random.default_rng(1)
    X = rng.random((700, 4))
    P = onp.predict(ff, X)
    h = rmock.call("bartint_R_forest", *_forest_args(rmock, ff, x_train, strings, fits))
    integ = rmock.to_numpy(rmock.call("bartint_R_exact_integrals", h, rmock.integer(0), rmock.integer(0)))
    np.testing.assert_allclose(integ, onp.sample_integrals(ff, "uniform"), rtol=1e-12, atol=1e-15)
    w = rng.random(700)
    out = rmock.call("bartint_R_eval", h, rmock.real(X), rmock.real(w, dim=False), rmock.integer(7))
    rm, cv, full = (rmock.to_numpy(rmock.lib.VECTOR_ELT(out, i)) for i in range(3))
    np.testing.assert_allclose(full, P, rtol=1e-12)
    np.testing.assert_allclose(rm, onp.row_means(P), rtol=1e-12)
    np.testing.assert_allclose(cv, onp.col_vars(P) * w, rtol=1e-9, atol=1e-18)
    rmock.lib.mock_finalize(h)
    step = rmock.call("bartint_R_design_step", *_forest_args(rmock, ff, x_train, strings, fits), rmock.real(X[:50]),
                      rmock.real([1.0], dim=False), rmock.real(X), rmock.integer(0), rmock.integer(0))
    integ2, ims, cv2, dm, dmv = (rmock.to_numpy(rmock.lib.VECTOR_ELT(step, i)) for i in range(5))
    np.testing.assert_allclose(integ2, integ, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(cv2, onp.col_vars(P[:, :50]), rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(dm, onp.row_means(P), rtol=1e-12)
    mean, sd = onp.finalize_integrals(integ, ff.y_min, ff.y_max)[1:]
    np.testing.assert_allclose(ims, [mean, sd], rtol=1e-10)
    assert dmv.shape == (2,)
