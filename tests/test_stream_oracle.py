"""Row f4 oracle pinned: the integer half/bfloat16 conversions against numpy/torch over
every 16-bit pattern and adversarial fp32 inputs, and the op semantics against numpy."""
import numpy as np
import pytest
import torch

import oracle


def test_half_conversions_match_numpy_exhaustively():
    h = np.arange(65536, dtype=np.uint16)
    fn = h.view(np.float16).astype(np.float32)
    assert oracle.first_mismatch(oracle.half_to_float(h), fn) == -1
    rng = np.random.default_rng(7)
    x = rng.integers(0, 1 << 32, 1_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    finite = fn[np.isfinite(fn)]
    x = np.concatenate([x, fn, np.nextafter(finite, np.float32(np.inf)), np.nextafter(finite, np.float32(-np.inf)),
                        (finite.astype(np.float64) * (1 + 2.0 ** -11)).astype(np.float32),      # exact ties
                        np.array([65504.0, 65519.996, 65520.0, 6e-8, 2.98e-8, 2.9802322e-8, 0.0, -0.0], np.float32)])
    with np.errstate(all="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    assert oracle.first_mismatch_bits(oracle.float_to_half(x), want, "f16") == -1


def test_bf16_conversions_match_torch():
    h = np.arange(65536, dtype=np.uint16)
    want = torch.from_numpy(h.view(np.int16).copy()).view(torch.bfloat16).float().numpy()
    assert oracle.first_mismatch(oracle.bf16_to_float(h), want) == -1
    rng = np.random.default_rng(8)
    x = rng.integers(0, 1 << 32, 1_000_000, dtype=np.uint64).astype(np.uint32)
    x = np.concatenate([x, (h.astype(np.uint32) << 16) | 0x8000, (h.astype(np.uint32) << 16) | 0x7FFF,
                        (h.astype(np.uint32) << 16) | 0x8001]).view(np.float32)
    want = torch.from_numpy(x.copy()).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert oracle.first_mismatch_bits(oracle.float_to_bf16(x), want, "bf16") == -1


@pytest.mark.parametrize("dtype,npdt", [("f32", np.float32), ("f64", np.float64)])
def test_native_ops_match_numpy(dtype, npdt):
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(10_001).astype(npdt), rng.standard_normal(10_001).astype(npdt)
    s = 3.0
    assert np.array_equal(oracle.stream("copy", dtype, a, None), a)
    assert np.array_equal(oracle.stream("scale", dtype, a, None, s), npdt(s) * a)
    assert np.array_equal(oracle.stream("add", dtype, a, b), a + b)
    # triad is a single fused rounding: compare with the exact result rounded once
    wide = np.longdouble if npdt is np.float64 else np.float64
    want = (a.astype(wide) + wide(s) * b.astype(wide)).astype(npdt)
    assert np.array_equal(oracle.stream("triad", dtype, a, b, s), want)


def test_f32_add_is_the_vectoradd_oracle():
    a, b = oracle.fill_rand(50000)
    assert np.array_equal(oracle.stream("add", "f32", a, b), oracle.vadd(a, b))


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_half_ops_widen_compute_round(dtype):
    rng = np.random.default_rng(4)
    to_f = oracle.half_to_float if dtype == "f16" else oracle.bf16_to_float
    to_h = oracle.float_to_half if dtype == "f16" else oracle.float_to_bf16
    a = rng.integers(0, 1 << 16, 50_000, dtype=np.uint32).astype(np.uint16)
    b = rng.integers(0, 1 << 16, 50_000, dtype=np.uint32).astype(np.uint16)
    fa, fb = to_f(a), to_f(b)
    with np.errstate(all="ignore"):
        assert oracle.first_mismatch_bits(oracle.stream("add", dtype, a, b), to_h(fa + fb), dtype) == -1
        assert oracle.first_mismatch_bits(oracle.stream("scale", dtype, a, None, 0.3), to_h(np.float32(0.3) * fa), dtype) == -1
    assert np.array_equal(oracle.stream("copy", dtype, a, None), a)
    # fp16 add computed through fp32 equals numpy's correctly rounded fp16 add (p' >= 2p+2)
    if dtype == "f16":
        with np.errstate(all="ignore"):
            want = (a.view(np.float16) + b.view(np.float16)).view(np.uint16)
        assert oracle.first_mismatch_bits(oracle.stream("add", dtype, a, b), want, dtype) == -1
