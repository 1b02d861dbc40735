"""Row f4 on the GPU: b200va_stream (copy/scale/add/triad x f32/f64/f16/bf16) bit-exact
against the oracle through the C ABI -- ragged sizes, misaligned pointers, special values."""
import numpy as np
import pytest

import oracle
from conftest import has_gpu

pytestmark = pytest.mark.gpu
if has_gpu():
    import torch

    import k8s_gpu_hpa_b200 as pkg
    from k8s_gpu_hpa_b200 import capi, vector_add as va

OPS = ["copy", "scale", "add", "triad"]
DTS = ["f32", "f64", "f16", "bf16"]
SIZES = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 1023, 4097, 50000, (1 << 20) + 3, 5_000_011]


def host_inputs(dtype, n, seed):
    rng = np.random.default_rng(seed)
    if dtype == "f32":
        return (rng.standard_normal(n) * 4).astype(np.float32), (rng.standard_normal(n) * 4).astype(np.float32)
    if dtype == "f64":
        return rng.standard_normal(n) * 4, rng.standard_normal(n) * 4
    # half types: every bit pattern class incl. subnormals / inf / nan
    return (rng.integers(0, 1 << 16, n, dtype=np.uint32).astype(np.uint16),
            rng.integers(0, 1 << 16, n, dtype=np.uint32).astype(np.uint16))


def to_dev(x, dtype):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    if dtype == "f16":
        return t.view(torch.float16)
    if dtype == "bf16":
        return t.view(torch.bfloat16)
    return t


def to_host(t, dtype):
    if dtype in ("f16", "bf16"):
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


@pytest.mark.parametrize("dtype", DTS)
@pytest.mark.parametrize("op", OPS)
def test_stream_ops_ragged_sizes(op, dtype):
    nmax = max(SIZES)
    ha, hb = host_inputs(dtype, nmax, 11)
    a, b = to_dev(ha, dtype), to_dev(hb, dtype)
    s = 0.7001953125 if op in ("scale", "triad") else 0.0
    want = oracle.stream(op, dtype, ha, hb if op in ("add", "triad") else None, s)
    for n in SIZES:
        out = torch.zeros(n + 16, dtype=a.dtype, device="cuda")
        va.stream(op, a[:n], b[:n] if op in ("add", "triad") else None, out[:n], scalar=s)
        torch.cuda.synchronize()
        bad = oracle.first_mismatch_bits(to_host(out[:n], dtype), want[:n], dtype)
        assert bad < 0, f"{op} {dtype} n={n}: first mismatch at {bad}"
        assert bool((out[n:] == 0).all())


@pytest.mark.parametrize("dtype", DTS)
def test_stream_misaligned_and_aliased(dtype):
    n = 20_011
    ha, hb = host_inputs(dtype, n + 32, 12)
    a, b = to_dev(ha, dtype), to_dev(hb, dtype)
    for oa, ob, oc in [(0, 0, 0), (1, 1, 1), (3, 3, 3), (5, 5, 5), (1, 2, 3), (0, 1, 0), (7, 7, 0)]:
        out = torch.zeros(n + 32, dtype=a.dtype, device="cuda")
        va.stream("triad", a[oa:oa + n], b[ob:ob + n], out[oc:oc + n], scalar=-1.5)
        torch.cuda.synchronize()
        want = oracle.stream("triad", dtype, ha[oa:oa + n].copy(), hb[ob:ob + n].copy(), -1.5)
        assert oracle.first_mismatch_bits(to_host(out[oc:oc + n], dtype), want, dtype) == -1, (oa, ob, oc)
        assert bool((out[:oc] == 0).all()) and bool((out[oc + n:] == 0).all())
    x = to_dev(ha[:n], dtype)
    va.stream("triad", x, b[:n], x, scalar=2.0)                       # y = y + 2*x in place (axpy)
    assert oracle.first_mismatch_bits(to_host(x, dtype), oracle.stream("triad", dtype, ha[:n].copy(), hb[:n].copy(), 2.0), dtype) == -1


def test_stream_f32_add_is_bit_identical_to_the_hot_path():
    n = 3_000_017
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    va.fill_ctr(a, 0x0A)
    va.fill_ctr(b, 0x0B)
    c1, c2 = va.add(a, b), va.stream("add", a, b)
    assert bool((c1.view(torch.int32) == c2.view(torch.int32)).all())
    assert va.digest(c2) == oracle.ctr_vadd_digest(n)


def test_stream_special_values_f32_and_f64():
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "special_values.npz"))
    ua, ub = np.tile(g["a_bits"], 9), np.tile(g["b_bits"], 9)
    a, b = to_dev(ua.view(np.float32), "f32"), to_dev(ub.view(np.float32), "f32")
    for op, s in (("add", 0.0), ("triad", 1.0), ("triad", -0.5), ("scale", 2.0 ** -10)):
        got = to_host(va.stream(op, a, b if op != "scale" else None, scalar=s), "f32")
        want = oracle.stream(op, "f32", ua.view(np.float32), ub.view(np.float32) if op != "scale" else None, s)
        assert oracle.first_mismatch_bits(got, want, "f32") == -1, op
    da = np.array([0.0, -0.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, np.inf, -np.inf, np.nan, 1.0, 1 + 2 ** -52])
    xa, xb = np.repeat(da, len(da)), np.tile(da, len(da))
    for op, s in (("add", 0.0), ("triad", 3.0), ("scale", 0.5)):
        got = to_host(va.stream(op, to_dev(xa, "f64"), to_dev(xb, "f64") if op != "scale" else None, scalar=s), "f64")
        assert oracle.first_mismatch_bits(got, oracle.stream(op, "f64", xa, xb if op != "scale" else None, s), "f64") == -1, op


def test_stream_argument_errors():
    a = torch.zeros(64, dtype=torch.float32, device="cuda")
    L = capi.lib
    assert L.b200va_stream(9, 0, a.data_ptr(), a.data_ptr(), a.data_ptr(), 64, 0.0, None) == capi.ERR_VARIANT
    assert L.b200va_stream(2, 7, a.data_ptr(), a.data_ptr(), a.data_ptr(), 64, 0.0, None) == capi.ERR_VARIANT
    assert L.b200va_stream(2, 0, a.data_ptr(), None, a.data_ptr(), 64, 0.0, None) == capi.ERR_INVALID     # add needs b
    assert L.b200va_stream(0, 0, a.data_ptr(), None, a.data_ptr(), 64, 0.0, None) == capi.OK             # copy does not
    assert L.b200va_stream(2, 1, a.data_ptr() + 4, a.data_ptr(), a.data_ptr(), 4, 0.0, None) == capi.ERR_ALIGN  # f64 needs 8 B
    assert L.b200va_stream(2, 0, a.data_ptr(), a.data_ptr(), a.data_ptr() + 8, 32, 0.0, None) == capi.ERR_OVERLAP
