"""The estimators' one-off transfers (csrc/transfer.cu, include/bhmm_b200.h group 4): a list of pageable host arrays in, one
array per trajectory out -- byte-exact round trips over ragged, empty and slot-straddling arrays, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    import torch
    assert torch.cuda.is_available()
    import bhmm_b200.engine as e
    return e


@pytest.mark.parametrize('threads', [1, 3, 8])
@pytest.mark.parametrize('dtype', [np.float64, np.int32])
def test_ragged_round_trip(eng, monkeypatch, threads, dtype):
    import torch
    from bhmm_b200._lib import lib, check
    monkeypatch.setenv('BHMM_B200_TRANSFER_THREADS', str(threads))
    rng = np.random.default_rng(5 + threads)
    # lengths around the 2 MiB staging slot, tiny ones, empty ones, and one array that spans many slots
    item = np.dtype(dtype).itemsize
    slot = (2 << 20) // item
    lengths = [1, 0, 7, slot - 1, slot, slot + 1, 0, 3 * slot + 17, 5, 12345, 9 * slot + 3, 0, 2]
    arrays = [np.ascontiguousarray(rng.integers(-2 ** 31, 2 ** 31, size=n).astype(dtype)) for n in lengths]
    total = int(sum(lengths))
    dst = torch.empty(total, dtype=torch.float64 if dtype == np.float64 else torch.int32, device='cuda')
    eng.upload_arrays(dst, arrays)
    ref = np.concatenate(arrays)
    assert np.array_equal(dst.cpu().numpy(), ref)
    # one flat destination
    back = eng.download_array(dst)
    assert back.dtype == ref.dtype and np.array_equal(back, ref)
    # ragged destinations through the C ABI
    outs = [np.full(n, 77, dtype=dtype) for n in lengths]
    ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    nbytes = np.asarray([o.nbytes for o in outs], dtype=np.int64)
    check(lib.bhmm_b200_download_ragged(ptrs, C.c_void_p(dst.data_ptr()), nbytes.ctypes.data_as(C.POINTER(C.c_longlong)),
                                        len(outs), threads, None))
    for o, a in zip(outs, arrays):
        assert np.array_equal(o, a)


def test_batch_from_list_equals_batch_from_concatenated(eng):
    """TrajectoryBatch(list of host arrays) goes through the mover; its observations equal the concatenation, for float
    observations, integer symbols and arrays that need a dtype conversion first."""
    rng = np.random.default_rng(2)
    obs = [rng.normal(size=n) for n in (300, 1, 4097, 70000)]
    b = eng.TrajectoryBatch(obs, 3)
    assert np.array_equal(b.obs.cpu().numpy(), np.concatenate(obs))
    b.close()
    sym = [rng.integers(0, 5, size=n).astype(np.int64) for n in (50, 100000, 3)]     # int64 -> int32 on the way
    b = eng.TrajectoryBatch(sym, 4)
    assert b.discrete and np.array_equal(b.obs.cpu().numpy(), np.concatenate(sym).astype(np.int32))
    b.close()
    f32 = [rng.normal(size=n).astype(np.float32) for n in (10, 2000)]
    b = eng.TrajectoryBatch(f32, 2)
    assert np.array_equal(b.obs.cpu().numpy(), np.concatenate(f32).astype(np.float64))
    b.close()


def test_transfer_rejects_bad_arguments(eng):
    import torch
    from bhmm_b200 import _lib
    dst = torch.empty(10, dtype=torch.float64, device='cuda')
    with pytest.raises(ValueError):
        eng.upload_arrays(dst, [np.zeros(9)])
    with pytest.raises(TypeError):
        eng.upload_arrays(dst, [np.zeros(20, dtype=np.float32)])
    rc = _lib.lib.bhmm_b200_upload_ragged(None, None, None, 3, 0, None)
    assert rc == _lib.ERR_INVALID
