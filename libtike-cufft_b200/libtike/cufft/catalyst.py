"""Data-ingestion conventions of the reference's catalyst driver, for real ptycho-tomography scans.

The reference keeps this code in a user script (tests/catalyst/test_rec_script.py): `PtychoDAO`
with its `h5_reader` (:11-102) and the normalisation steps of the driver (:182-210).  It is the
step in front of the hot path: what is on disk -> the arrays `CGPtychoSolver.run_batch` takes.
Same names and options here; two differences:

  * the file is optional: `PtychoDAO.from_arrays` takes the datasets / attributes of the file as
    arrays (h5py is not a dependency; `h5_reader` imports it lazily and raises ImportError if absent);
  * the data-side work -- keep the in-view frames, `fftshift` them, divide by max|probe|^2
    (:44-46, :98-100, :209) -- the reference marks "needs to be moved to compute kernels" (:43); here
    it is ONE pass over the frames on the device (`prepare_data_device`, ptx_prepare_data) instead of
    three host passes, bit-identical to the host result.
"""
import ctypes
import re

import numpy as np
import torch

from libtike.cufft.ptychofft import lib, check, current_stream

__all__ = ["PtychoDAO", "driver_prepare", "prepare_data_device"]


class PtychoDAO(object):
    def __init__(self, pid, data, positions, probes, rotation_angle=None):
        self.pid = pid
        self.data = data
        self.positions = positions
        self.probes = probes
        self.rotation_angle = rotation_angle
        self.ids = None  # frames kept by the in-view filter (indices into the file's /data)

    @classmethod
    def from_arrays(cls, data, positions_0, positions_1, initprobe, recprobe, attrs, pid=None,
                    use_original_positions=False, swap_position_axes=True,
                    reset_position_coordinates=True, use_original_probes=False,
                    swap_probe_axes=False, data_fftshift=True, view_dims=(2048, 2048),
                    map_position_detector_pixel=1., defer_data=False):
        """h5_reader (test_rec_script.py:21-102) on the file's contents.  `defer_data=True` leaves
        `data` untouched (raw frames, all of them) and records `ids`, so that selection, fftshift
        and normalisation can run as one device pass (`prepare_data_device`)."""
        probes = np.array(initprobe if use_original_probes else recprobe, dtype=np.complex64, order='C')
        if swap_probe_axes:
            probes = np.array(probes.swapaxes(1, 2), order='C')
        positions = np.array(positions_0 if use_original_positions else positions_1,
                             dtype=np.float32, order='C')
        pos2det_const = np.float64(((attrs.get('detector_pixel_size') * probes.shape[-1]) /
                                    (attrs.get('detector_distance') * 1e-10
                                     * attrs.get('incident_wavelength')))
                                   * map_position_detector_pixel)
        positions = np.float32(positions * pos2det_const)
        if swap_position_axes:
            positions[:, (0, 1)] = positions[:, (1, 0)]
        if not reset_position_coordinates:
            # the reference raises here too (an undefined `ValueException`, i.e. a NameError, :97)
            raise ValueError("Currently reset_position_coordinates has to be set to True.")
        positions[:, 0] = positions[:, 0] - min(positions[:, 0])
        positions[:, 1] = positions[:, 1] - min(positions[:, 1])
        ids = np.where((positions[:, 1] >= 0) * (positions[:, 1] < view_dims[1]) *
                       (positions[:, 0] >= 0) * (positions[:, 0] < view_dims[0]))[0]
        positions = np.array(positions[ids, :], dtype=np.float32, order='C')
        if defer_data:
            out = np.asarray(data)
        else:
            out = np.array(data, dtype=np.float32, order='C')
            if data_fftshift:
                out = np.fft.fftshift(out[:], axes=(1, 2))
            out = out[ids]
        dao = cls(pid, out, positions, probes, attrs.get('rotation_angle'))
        dao.ids = ids
        return dao

    @classmethod
    def h5_reader(cls, input_file, pid=None, **options):
        """test_rec_script.py:21-102 with the same keyword options; needs h5py."""
        try:
            import h5py
        except ImportError as e:  # pragma: no cover - h5py is absent from this image
            raise ImportError("PtychoDAO.h5_reader needs h5py; use PtychoDAO.from_arrays") from e
        if pid is None:
            pid = np.int32(re.findall(r'\d+', input_file)[-2])
        with h5py.File(input_file, 'r') as fid:
            return cls.from_arrays(fid['data'], fid['/positions_0'], fid['/positions_1'],
                                   fid['/initprobe'], fid['/recprobe'], dict(fid.attrs), pid=pid,
                                   **options)


def prepare_data_device(raw, ids=None, denominator=1.0, fftshift=True):
    """out[s] = fftshift(raw[ids[s]]) / denominator as ONE device pass (ptx_prepare_data).

    raw: [F, N, N] float32, host (copied up once) or device; ids: kept frame indices (None = all);
    denominator: float32 divisor, e.g. max|probe|^2 (test_rec_script.py:209).  Returns a float32
    CUDA tensor [len(ids), N, N], bit-identical to the reference's three host passes."""
    t = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32))
    t = t.cuda().contiguous()
    if t.dtype != torch.float32 or t.ndim != 3 or t.shape[1] != t.shape[2]:
        raise ValueError("raw frames must be [F, N, N] float32")
    nsel = t.shape[0] if ids is None else len(ids)
    d_ids = None
    if ids is not None:
        d_ids = torch.as_tensor(np.asarray(ids, dtype=np.int64)).cuda()
        if nsel and (int(d_ids.min()) < 0 or int(d_ids.max()) >= t.shape[0]):
            raise IndexError("frame index out of range")
    out = torch.empty((nsel, t.shape[1], t.shape[2]), dtype=torch.float32, device=t.device)
    if nsel:
        check(lib.ptx_prepare_data(ctypes.c_void_p(t.data_ptr()),
                                   ctypes.c_void_p(d_ids.data_ptr()) if d_ids is not None else None,
                                   nsel, t.shape[1], ctypes.c_float(np.float32(denominator)),
                                   1 if fftshift else 0, ctypes.c_void_p(out.data_ptr()),
                                   current_stream()))
    return out


def driver_prepare(dao, nmodes, view_dims, device_data=False):
    """The driver's steps between the reader and the solver (test_rec_script.py:182-210).

    Returns (data, psi, scan, prb) shaped for `CGPtychoSolver.run_batch` (one angle).  With
    `device_data` the DAO must come from `from_arrays(..., defer_data=True)` and `data` is returned
    as a CUDA tensor prepared by `prepare_data_device`."""
    prb = np.array(dao.probes, copy=True)
    prb.shape = (1,) + prb.shape
    scan = np.array(dao.positions, copy=True)
    scan.shape = (1,) + scan.shape
    ndet = dao.data.shape[-1]
    psi = np.zeros((1, view_dims[0] + ndet, view_dims[1] + ndet), dtype='complex64', order='C') \
        + 1 * np.exp(-1j * 0.25)
    prb = prb[:, :nmodes]
    den = np.amax(np.abs(prb)) ** 2
    if device_data:
        data = prepare_data_device(dao.data, dao.ids, den, fftshift=True)[None]
    else:
        data = np.array(dao.data, copy=True)
        data.shape = (1,) + data.shape
        data /= den
    prb = prb / np.amax(np.abs(prb))
    return data, psi, scan, prb
