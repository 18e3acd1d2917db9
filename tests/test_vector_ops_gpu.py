"""Every b200mf_vec_* entry point (LinearAlgebra::distributed::Vector BLAS-1,
lac/vector_operations_internal.h:2140-2660) against torch, FP64 and FP32, at sizes that are not a
multiple of the block / grid-stride (1, 255, 1000, 100003, 3 M)."""
import ctypes as C

import pytest
import torch

from dealii_b200 import _lib as L

pytestmark = pytest.mark.gpu
SIZES = [1, 255, 1000, 100003, 3_000_017]


def ptr(t):
    return C.c_void_p(t.data_ptr())


def setup(number, n, k):
    dt = torch.float64 if number == "f64" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(n % 1000)
    vs = [torch.rand(n, dtype=dt, device="cuda", generator=g) - 0.3 for _ in range(k)]
    return L.load(), (L.F64 if number == "f64" else L.F32), vs, (1e-13 if number == "f64" else 2e-5)


@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("n", SIZES)
def test_updates(number, n):
    lib, code, (x, y, w, d), tol = setup(number, n, 4)

    def close(a, b):
        assert (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1.0)

    v = y.clone(); L.check(lib.b200mf_vec_set(code, ptr(v), 2.5, n, None)); close(v, torch.full_like(v, 2.5))
    v = y.clone(); L.check(lib.b200mf_vec_axpy(code, ptr(v), -0.7, ptr(x), n, None)); close(v, y - 0.7 * x)
    v = y.clone(); L.check(lib.b200mf_vec_sadd(code, ptr(v), 1.3, -0.7, ptr(x), n, None)); close(v, 1.3 * y - 0.7 * x)
    v = y.clone(); L.check(lib.b200mf_vec_sadd_xavbw(code, ptr(v), 1.3, -0.7, ptr(x), 0.25, ptr(w), n, None))
    close(v, 1.3 * y - 0.7 * x + 0.25 * w)
    v = y.clone(); L.check(lib.b200mf_vec_equ(code, ptr(v), 3.0, ptr(x), 0.0, None, n, None)); close(v, 3.0 * x)
    v = y.clone(); L.check(lib.b200mf_vec_equ(code, ptr(v), 3.0, ptr(x), -2.0, ptr(w), n, None)); close(v, 3.0 * x - 2.0 * w)
    v = y.clone(); L.check(lib.b200mf_vec_scale(code, ptr(v), 0.5, None, n, None)); close(v, 0.5 * y)
    v = y.clone(); L.check(lib.b200mf_vec_scale(code, ptr(v), 1.0, ptr(d), n, None)); close(v, y * d)
    v = y.clone(); L.check(lib.b200mf_vec_scale_by(code, ptr(v), ptr(d), ptr(x), n, None)); close(v, d * x)
    torch.cuda.synchronize()


@pytest.mark.parametrize("number", ["f64", "f32"])
@pytest.mark.parametrize("n", SIZES)
def test_reductions(number, n):
    lib, code, (x, y, w), tol = setup(number, n, 3)
    r = C.c_double()
    xd, yd, wd = x.double(), y.double(), w.double()

    def rel(a, b):
        assert abs(a - b) <= 50 * tol * max(abs(b), float(xd.abs().max() * yd.abs().max())), (a, b)

    L.check(lib.b200mf_vec_dot(code, ptr(x), ptr(y), n, C.byref(r), None)); rel(r.value, float(xd @ yd))
    L.check(lib.b200mf_vec_norm_sqr(code, ptr(x), n, C.byref(r), None)); rel(r.value, float(xd @ xd))
    L.check(lib.b200mf_vec_norm_2(code, ptr(x), n, C.byref(r), None)); rel(r.value, float(xd.norm()))
    L.check(lib.b200mf_vec_norm_1(code, ptr(x), n, C.byref(r), None)); rel(r.value, float(xd.abs().sum()))
    L.check(lib.b200mf_vec_norm_linfty(code, ptr(x), n, C.byref(r), None)); assert r.value == float(xd.abs().max())
    acc = torch.full((1,), 1.5, dtype=torch.float64, device="cuda")
    L.check(lib.b200mf_vec_dot_device(code, ptr(x), ptr(y), n, ptr(acc), None)); rel(float(acc) - 1.5, float(xd @ yd))
    v = y.clone()
    L.check(lib.b200mf_vec_add_and_dot(code, ptr(v), -0.4, ptr(x), ptr(w), n, C.byref(r), None))
    ref = y - 0.4 * x
    assert (v - ref).abs().max().item() <= tol * max(ref.abs().max().item(), 1.0)
    rel(r.value, float(ref.double() @ wd))
