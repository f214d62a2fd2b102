"""The plain-C oracle (oracle/vrr_oracle.c, built by __graft_entry__.build_oracle) against the numpy
oracle and the golden fixtures generated from the unmodified reference.  CPU only."""
import ctypes
import os

import numpy as np
import pytest

from oracle import attention_np as A
from oracle import tables_np as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(ROOT, "oracle", "_build", "libvrr_oracle.so")


@pytest.fixture(scope="module")
def clib():
    import __graft_entry__ as g
    g.build_oracle()
    assert os.path.isfile(LIB)
    return ctypes.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("length", [17, 65])
def test_c_relative_index_matches_reference_fixture(clib, length):
    want = np.load(os.path.join(GOLDEN, "tables.npz"))[f"rel_index_L{length}"]
    got = np.empty((length, length), np.int64)
    clib.vrr_c_relative_index(length, _p(got))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("g", [8, 14])
def test_c_grid_coords_and_poly_distance(clib, g):
    z = np.load(os.path.join(GOLDEN, "tables.npz"))
    tx, ty = np.empty(g * g, np.float32), np.empty(g * g, np.float32)
    clib.vrr_c_grid_coords(g, g, _p(tx), _p(ty))
    assert np.array_equal(tx, z[f"t_x_g{g}"]) and np.array_equal(ty, z[f"t_y_g{g}"])
    d = np.empty((g * g, g * g), np.int64)
    clib.vrr_c_poly_l1(g, _p(d))
    assert np.array_equal(d, T.poly_l1_distance(g * g))


def test_c_mixed_scramble(clib):
    H, N = 6, 64
    hs, ps = np.empty((H, N), np.int64), np.empty((H, N), np.int64)
    clib.vrr_c_mixed_scramble(H, N, _p(hs), _p(ps))
    hs_np, ps_np = T.mixed_scramble(H, N)
    assert np.array_equal(hs, hs_np) and np.array_equal(ps, ps_np)


@pytest.mark.parametrize("kind", ["none", "table", "poly", "poly_perhead"])
def test_c_attention_core_matches_numpy(clib, kind):
    rng = np.random.default_rng(3)
    B, H, N, D = 2, 3, 17, 8
    q, k, v = (np.ascontiguousarray(rng.standard_normal((B, H, N, D))) for _ in range(3))
    d_out = np.ascontiguousarray(rng.standard_normal((B, N, H * D)))
    scale = D ** -0.5
    mode, param, heads, length, grid, bias = 0, np.zeros(1), 0, 0, 0, None
    if kind == "table":
        param = np.ascontiguousarray(rng.standard_normal((H, 2 * N - 1)) * 0.5)
        mode, heads, length, bias = 1, H, 2 * N - 1, T.relative_bias(param, N)
    elif kind.startswith("poly"):
        shared = kind == "poly"
        param = np.ascontiguousarray(rng.standard_normal(4 if shared else (H, 4)) * 0.05)
        mode, heads, length, grid = 2, 1 if shared else H, 4, 4
        bias = T.poly_bias(param, N - 1, H).astype(np.float64)
    clib.vrr_c_attn_fwd.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    out, lse = np.empty((B, N, H * D)), np.empty((B, H, N))
    clib.vrr_c_attn_fwd(_p(q), _p(k), _p(v), B, H, N, D, scale, mode, _p(param), heads, length, grid, _p(out), _p(lse))
    o_np, p_np, _, _ = A.attention_forward(q, k, v, scale, bias)
    np.testing.assert_allclose(out, o_np, rtol=0, atol=1e-6 if kind.startswith("poly") else 1e-12)
    clib.vrr_c_attn_bwd.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
    dq, dk, dv = (np.empty((B, H, N, D)) for _ in range(3))
    d_bias = np.zeros((H, N, N))
    clib.vrr_c_attn_bwd(_p(q), _p(k), _p(v), _p(out), _p(d_out), _p(lse), B, H, N, D, scale, mode, _p(param), heads,
                        length, grid, _p(dq), _p(dk), _p(dv), _p(d_bias))
    g = A.attention_backward(d_out, q, k, v, scale, bias)
    tol = 1e-6 if kind.startswith("poly") else 1e-11
    for got, want in ((dq, g["dq"]), (dk, g["dk"]), (dv, g["dv"]), (d_bias, g["dbias"])):
        np.testing.assert_allclose(got, want, rtol=0, atol=tol)
