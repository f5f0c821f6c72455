"""Data-ingestion conventions of the reference's catalyst driver (tests/catalyst/test_rec_script.py:
21-102 reader, :182-210 driver), SURVEY.md section 8 row f3: host logic against the restatement of
oracle/numpy_catalyst.py on a synthetic "file" (bit-exact), and the fused device pass
(selection + fftshift + normalisation, ptx_prepare_data) against the same."""
import numpy as np
import pytest

from oracle import numpy_catalyst as OC


def synthetic_file(nframes=40, ndet=64, nmodes=3, seed=0):
    rng = np.random.default_rng(seed)
    fid = {
        "data": rng.poisson(20.0, size=(nframes, ndet, ndet)).astype(np.float32),
        "positions_0": rng.uniform(-2e-6, 2e-6, size=(nframes, 2)).astype(np.float32),
        "positions_1": rng.uniform(-2e-6, 2e-6, size=(nframes, 2)).astype(np.float32),
        "initprobe": (rng.normal(size=(nmodes, ndet, ndet)) + 1j * rng.normal(size=(nmodes, ndet, ndet))).astype(np.complex64),
        "recprobe": (rng.normal(size=(nmodes, ndet, ndet)) + 1j * rng.normal(size=(nmodes, ndet, ndet))).astype(np.complex64),
    }
    attrs = {"detector_pixel_size": 75e-6, "detector_distance": 2.0, "incident_wavelength": 1.4,
             "rotation_angle": 12.5}
    return fid, attrs


OPTIONS = [
    {},
    {"use_original_positions": True, "swap_position_axes": False},
    {"use_original_probes": True, "swap_probe_axes": True, "data_fftshift": False},
    {"view_dims": (60, 90), "map_position_detector_pixel": 0.5},
]


@pytest.mark.parametrize("opts", OPTIONS)
def test_from_arrays_matches_reader_restatement(opts):
    from libtike.cufft.catalyst import PtychoDAO, driver_prepare
    fid, attrs = synthetic_file()
    o = dict({"view_dims": (300, 200)}, **opts)
    want = OC.h5_reader_arrays(fid, attrs, **o)
    dao = PtychoDAO.from_arrays(fid["data"], fid["positions_0"], fid["positions_1"], fid["initprobe"],
                                fid["recprobe"], attrs, pid=7, **o)
    for got, ref in zip((dao.data, dao.positions, dao.probes), want[:3]):
        assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref)
    assert dao.rotation_angle == want[3] and dao.pid == 7
    assert 0 < len(dao.ids) <= 40 and dao.positions.min() >= 0
    w = OC.driver_prepare(want[0], want[1], want[2], 2, o["view_dims"])
    g = driver_prepare(dao, 2, o["view_dims"])
    for got, ref in zip(g, w):
        assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref)
    assert g[1].shape == (1, o["view_dims"][0] + 64, o["view_dims"][1] + 64)


def test_reset_position_coordinates_is_mandatory():
    from libtike.cufft.catalyst import PtychoDAO
    fid, attrs = synthetic_file()
    with pytest.raises(ValueError):
        PtychoDAO.from_arrays(fid["data"], fid["positions_0"], fid["positions_1"], fid["initprobe"],
                              fid["recprobe"], attrs, reset_position_coordinates=False)


def test_h5_reader_needs_h5py():
    from libtike.cufft.catalyst import PtychoDAO
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py present")
    except ImportError:
        with pytest.raises(ImportError):
            PtychoDAO.h5_reader("/nonexistent/extracted_scan339.h5")


def test_h5_reader_lines_on_an_in_memory_file(monkeypatch):
    """h5py is absent from this image, so `PtychoDAO.h5_reader`'s own lines (test_rec_script.py:21-40:
    the scan id parsed out of the file name, the dataset names, the attributes) are exercised through a
    stand-in `h5py` module whose File is an in-memory mapping with `.attrs` -- the reader must hand
    exactly those datasets to `from_arrays`."""
    import sys
    import types
    from libtike.cufft.catalyst import PtychoDAO
    fid, attrs = synthetic_file()
    opened = []

    class File(dict):
        def __init__(self, name, mode):
            assert mode == "r"
            opened.append(name)
            super().__init__({("/" + k): v for k, v in fid.items()})
            self.attrs = dict(attrs)

        def __getitem__(self, key):  # h5py resolves 'data' and '/data' alike
            return super().__getitem__(key if key.startswith("/") else "/" + key)

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=File))
    name = "/data/run_12/extracted_scan339.h5"
    dao = PtychoDAO.h5_reader(name, view_dims=(300, 200))
    assert opened == [name]
    assert dao.pid == 339  # the second-to-last number in the name; the last is the 5 of ".h5" (test_rec_script.py:36)
    want = PtychoDAO.from_arrays(fid["data"], fid["positions_0"], fid["positions_1"], fid["initprobe"],
                                 fid["recprobe"], attrs, pid=339, view_dims=(300, 200))
    for a, b in ((dao.data, want.data), (dao.positions, want.positions), (dao.probes, want.probes)):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert dao.rotation_angle == want.rotation_angle
    assert PtychoDAO.h5_reader(name, pid=5, view_dims=(300, 200)).pid == 5


@pytest.mark.gpu
@pytest.mark.parametrize("ndet,shift", [(64, True), (128, True), (256, False)])
def test_device_data_preparation_is_bit_identical(ndet, shift):
    import torch
    from libtike.cufft.catalyst import PtychoDAO, driver_prepare, prepare_data_device
    fid, attrs = synthetic_file(nframes=50, ndet=ndet, seed=ndet)
    o = {"view_dims": (int(ndet * 0.8), int(ndet * 0.9)), "data_fftshift": shift}  # crops some positions
    want = OC.h5_reader_arrays(fid, attrs, **o)
    w = OC.driver_prepare(want[0], want[1], want[2], 2, o["view_dims"])
    if shift:
        dao = PtychoDAO.from_arrays(fid["data"], fid["positions_0"], fid["positions_1"], fid["initprobe"],
                                    fid["recprobe"], attrs, defer_data=True, **o)
        assert dao.data.shape[0] == 50 and len(dao.ids) < 50  # raw frames kept, filter recorded
        g = driver_prepare(dao, 2, o["view_dims"], device_data=True)
        assert isinstance(g[0], torch.Tensor) and g[0].is_cuda
        assert np.array_equal(g[0].cpu().numpy(), w[0])
        for got, ref in zip(g[1:], w[1:]):
            assert np.array_equal(got, ref)
    else:
        ids = np.array([3, 0, 49, 7])
        den = np.float32(3.7)
        got = prepare_data_device(fid["data"], ids, den, fftshift=False).cpu().numpy()
        assert np.array_equal(got, fid["data"][ids] / den)
        every = prepare_data_device(torch.from_numpy(fid["data"]).cuda(), None, 1.0, fftshift=True)
        assert np.array_equal(every.cpu().numpy(), np.fft.fftshift(fid["data"], axes=(1, 2)))
    with pytest.raises(IndexError):
        prepare_data_device(fid["data"], np.array([50]), 1.0)
