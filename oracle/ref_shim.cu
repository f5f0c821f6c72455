// TEST INFRASTRUCTURE (oracle) -- C-ABI shim around the UNMODIFIED reference class.
//
// Nothing from the reference is copied into this repository: the reference's
// own translation unit /root/reference/src/cuda/ptychofft.cu (which itself
// includes kernels.cu, ptychofft.cu:2) is compiled where it lies by
// oracle/Makefile and this file only adds extern "C" entry points so that the
// class `ptychofft` (src/include/ptychofft.cuh:6-44) can be driven through
// ctypes with raw device pointers, exactly as the reference's pybind11/SWIG
// wrappers do (src/cuda/pybind11/ptychofft.cxx:8-26).  Output: oracle/_ref/.
#include "ptychofft.cuh"  // -I/root/reference/src/include
#include <cuda_runtime.h>

extern "C" {

void* ref_create(size_t ptheta, size_t nz, size_t n, size_t nscan, size_t ndet, size_t nprb) {
  return new ptychofft(ptheta, nz, n, nscan, ndet, nprb);
}
void ref_fwd(void* h, size_t g, size_t f, size_t scan, size_t prb) {
  static_cast<ptychofft*>(h)->fwd(g, f, scan, prb);
}
void ref_adj(void* h, size_t f, size_t g, size_t scan, size_t prb, int flg) {
  static_cast<ptychofft*>(h)->adj(f, g, scan, prb, flg);
}
void ref_free(void* h) { static_cast<ptychofft*>(h)->free(); }
void ref_destroy(void* h) { delete static_cast<ptychofft*>(h); }
// the reference checks no CUDA status; the oracle harness does.
int ref_last_cuda_error() { return static_cast<int>(cudaGetLastError()); }
int ref_device_sync() { return static_cast<int>(cudaDeviceSynchronize()); }

}  // extern "C"
