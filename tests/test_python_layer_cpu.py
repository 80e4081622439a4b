"""Host-side pieces of the Python mirror that need no device: how vectors reach the C ABI (`operators._vec_ptr`).
A Hermitian operator (`T = c64`, src/algorithms/mod.rs:167) takes its vectors as n complex numbers = 2 n interleaved doubles."""
import ctypes as C

import numpy as np
import pytest

from two_pass_lanczos_b200 import operators


def _read(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,)).copy()


def test_real_vectors_are_passed_as_contiguous_float64():
    v = np.arange(12, dtype=np.float32)[::2]  # strided, wrong dtype
    ptr, keep, is_torch = operators._vec_ptr(v)
    assert not is_torch and keep.dtype == np.float64 and keep.flags.c_contiguous and keep.shape == (6,)
    assert np.array_equal(_read(ptr, 6), np.arange(0, 12, 2, dtype=np.float64))


def test_complex_vectors_are_passed_interleaved():
    z = np.array([1 + 2j, -3.5 + 0.25j, 0 - 1j])
    ptr, keep, is_torch = operators._vec_ptr(z)
    assert not is_torch and keep.dtype == np.float64 and keep.shape == (6,)
    assert np.array_equal(_read(ptr, 6), np.array([1.0, 2.0, -3.5, 0.25, 0.0, -1.0]))
    # complex64 and non-contiguous input are converted, not reinterpreted
    z32 = np.array([1 + 2j, 9 + 9j, 3 - 4j], dtype=np.complex64)[::2]
    ptr, keep, _ = operators._vec_ptr(z32)
    assert np.array_equal(_read(ptr, 4), np.array([1.0, 2.0, 3.0, -4.0]))


def test_torch_vectors():
    torch = pytest.importorskip("torch")
    t = torch.arange(5, dtype=torch.float64)
    ptr, keep, is_torch = operators._vec_ptr(t)
    assert is_torch and keep.data_ptr() == t.data_ptr()  # consumed in place
    zc = torch.tensor([1 + 2j, 3 - 4j], dtype=torch.complex128)
    ptr, keep, is_torch = operators._vec_ptr(zc)
    assert is_torch and keep.dtype == torch.float64 and keep.tolist() == [1.0, 2.0, 3.0, -4.0]
    with pytest.raises(TypeError):
        operators._vec_ptr(torch.arange(3, dtype=torch.float32))
