"""Convert the reference's input fixtures to one compressed .npz (run HERE, not on the GPU box).

The reference ships its test *inputs* (not outputs) as TIFF/NPY under
/root/reference/tests/model/ (read by tests/test_adjoint.py:25-38,
tests/test.py:30-47, tests/test_modes.py:29-52 through dxchange, which is not
installed in this image).  /root/reference does not exist on the GPU box, so the
arrays are re-packed, bit-for-bit (float32 in, float32 out), into
tests/golden/model.npz which travels with the repo.

    python tests/golden/make_inputs.py
"""
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/tests/model/"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model.npz")


def read_tiff(name):
    im = Image.open(SRC + name)
    frames = []
    i = 0
    while True:
        try:
            im.seek(i)
        except EOFError:
            break
        frames.append(np.array(im, dtype=np.float32))
        i += 1
    return np.stack(frames) if len(frames) > 1 else frames[0]


if __name__ == "__main__":
    out = {
        name: read_tiff(name + ".tiff")
        for name in ("prbamp", "prbang", "probes_amp", "probes_ang", "initpsiamp", "initpsiang")
    }
    out["coords"] = np.load(SRC + "coords.npy").astype(np.float32)
    np.savez_compressed(DST, **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    print("wrote", DST, os.path.getsize(DST), "bytes")
