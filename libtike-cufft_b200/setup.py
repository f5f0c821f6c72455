"""Packaging of the B200-native libtike.cufft (replaces the reference's scikit-build setup.py:17-32).

The CUDA library is built by ../__graft_entry__.py:build() (plain nvcc, sm_100a) into
libtike/cufft/libptychofft_b200.so and shipped as package data; the plugin entry point of the
reference (setup.py:27-31) is kept so that `tike` finds the backend under the same name.
"""
from setuptools import setup, find_namespace_packages

setup(
    name="libtike-cufft",
    version="0.4.0+b200",
    packages=find_namespace_packages(include=["libtike.*"]),
    package_data={"libtike.cufft": ["libptychofft_b200.so"]},
    zip_safe=False,
    entry_points={
        "tike.PtychoBackend": [
            "cudafft = libtike.cufft.ptycho:PtychoCuFFT",
        ],
    },
)
