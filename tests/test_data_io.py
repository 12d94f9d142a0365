"""Output path: PFM files byte-identical to the reference writer's, asynchronous writer."""
import importlib.util
import os

import numpy as np
import pytest

from mvsformer_b200 import data_io

REF_IO = "/root/reference/datasets/data_io.py"


def _array():
    rng = np.random.RandomState(3)
    return (rng.rand(12, 20).astype(np.float32) * 900.0 + 400.0)


def test_pfm_roundtrip_and_layout(tmp_path):
    a = _array()
    path = str(tmp_path / "d.pfm")
    data_io.save_pfm(path, a)
    raw = open(path, "rb").read()
    assert raw.startswith(b"Pf\n20 12\n-1.000000\n")                      # greyscale, width height, little endian
    assert len(raw) == len(b"Pf\n20 12\n-1.000000\n") + a.size * 4
    assert np.array_equal(np.frombuffer(raw[-a.size * 4:], "<f4").reshape(12, 20), a[::-1])   # rows stored bottom-up
    back, scale = data_io.read_pfm(path)
    assert scale == 1.0 and np.array_equal(back, a)
    with pytest.raises(Exception):
        data_io.save_pfm(path, a.astype(np.float64))


@pytest.mark.skipif(not os.path.exists(REF_IO), reason="reference not present (build container only)")
def test_pfm_bytes_equal_reference_writer(tmp_path):
    spec = importlib.util.spec_from_file_location("ref_data_io", REF_IO)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for arr in (_array(), np.stack([_array()] * 3, axis=-1), _array()[:, :, None]):
        p1, p2 = str(tmp_path / "a.pfm"), str(tmp_path / "b.pfm")
        ref.save_pfm(p1, arr)
        data_io.save_pfm(p2, arr)
        assert open(p1, "rb").read() == open(p2, "rb").read()
        assert np.array_equal(ref.read_pfm(p2)[0], data_io.read_pfm(p1)[0])


def test_async_writer(tmp_path):
    a = _array()
    with data_io.AsyncResultWriter(workers=2, max_pending=2) as w:
        for i in range(6):
            buf = a + i                                                     # caller reuses nothing: fresh arrays
            w.submit(str(tmp_path / "scan" / "depth_est" / ("%08d.pfm" % i)), buf,
                     str(tmp_path / "scan" / "confidence" / ("%08d.npy" % i)), np.stack([buf] * 4, -1))
    for i in range(6):
        back, _ = data_io.read_pfm(str(tmp_path / "scan" / "depth_est" / ("%08d.pfm" % i)))
        assert np.array_equal(back, a + i)
        assert np.load(str(tmp_path / "scan" / "confidence" / ("%08d.npy" % i))).shape == (12, 20, 4)
    bad = data_io.AsyncResultWriter(workers=1)
    bad.submit(str(tmp_path / "x.pfm"), np.zeros((2, 2, 2), np.float32))    # invalid shape -> error surfaces on close
    with pytest.raises(Exception):
        bad.close()
