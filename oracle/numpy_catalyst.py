"""Restatement of the reference's catalyst data-ingestion conventions -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/tests/catalyst/test_rec_script.py statement by statement, with the HDF5 file
replaced by the arrays / attributes it holds (h5py is not part of this image):
  PtychoDAO.h5_reader ........ test_rec_script.py:21-102  -> h5_reader_arrays
  driver normalisation ....... test_rec_script.py:182-210 -> driver_prepare
The reference has no test or golden vector for this code; the restatement is pinned by reading.
Only tests/ may import this module.
"""
import numpy as np


def h5_reader_arrays(fid, attrs, use_original_positions=False, swap_position_axes=True,
                     reset_position_coordinates=True, use_original_probes=False,
                     swap_probe_axes=False, data_fftshift=True, view_dims=(2048, 2048),
                     map_position_detector_pixel=1.):
    """fid: mapping with 'data', 'positions_0', 'positions_1', 'initprobe', 'recprobe';
    attrs: mapping with detector_pixel_size, detector_distance, incident_wavelength, rotation_angle.
    Returns (data, positions, probes, rotation_angle)."""
    data = np.array(fid['data'], dtype=np.float32, order='C')                         # :41
    if data_fftshift:
        data = np.fft.fftshift(data[:], axes=(1, 2))                                  # :44-46
    detector_pixel_size = attrs.get('detector_pixel_size')                            # :49-52
    detector_distance = attrs.get('detector_distance')
    incident_wavelength = attrs.get('incident_wavelength')
    rotation_angle = attrs.get('rotation_angle')
    if use_original_probes:                                                           # :61-66
        probes = np.array(fid['initprobe'], dtype=np.complex64, order='C')
    else:
        probes = np.array(fid['recprobe'], dtype=np.complex64, order='C')
    if swap_probe_axes:                                                               # :67-70
        probes = np.array(probes.swapaxes(1, 2), order='C')
    if use_original_positions:                                                        # :72-77
        positions = np.array(fid['positions_0'], dtype=np.float32, order='C')
    else:
        positions = np.array(fid['positions_1'], dtype=np.float32, order='C')
    pos2det_const = np.float64(((detector_pixel_size * probes.shape[-1]) /             # :78-80
                                (detector_distance * 1e-10 * incident_wavelength))
                               * map_position_detector_pixel)
    positions = np.float32(positions * pos2det_const)
    if swap_position_axes:                                                            # :83-85
        positions[:, (0, 1)] = positions[:, (1, 0)]
    if reset_position_coordinates:                                                    # :86-95
        positions[:, 0] = positions[:, 0] - min(positions[:, 0])
        positions[:, 1] = positions[:, 1] - min(positions[:, 1])
        ids = np.where((positions[:, 1] >= 0) * (positions[:, 1] < view_dims[1]) *
                       (positions[:, 0] >= 0) * (positions[:, 0] < view_dims[0]))[0]
        positions = np.array(positions[ids, :], dtype=np.float32, order='C')
    else:
        raise ValueError("Currently reset_position_coordinates has to be set to True.")
    if ids is not None:                                                               # :98-100
        data = data[ids]
    return data, positions, probes, rotation_angle


def driver_prepare(data, positions, probes, nmodes, view_dims):
    """test_rec_script.py:182-210: add the angle axis, build the initial object, keep `nmodes`
    probes, normalise data by max|probe|^2 and the probes by max|probe|."""
    prb = probes.copy()
    prb.shape = (1,) + prb.shape
    scan = positions.copy()
    scan.shape = (1,) + scan.shape
    data = data.copy()
    data.shape = (1,) + data.shape
    psi = np.zeros((1, view_dims[0] + data.shape[-1], view_dims[1] + data.shape[-1]),
                   dtype='complex64', order='C') + 1 * np.exp(-1j * 0.25)
    prb = prb[:, :nmodes]
    data /= np.amax(np.abs(prb)) ** 2
    prb /= np.amax(np.abs(prb))
    return data, psi, scan, prb
