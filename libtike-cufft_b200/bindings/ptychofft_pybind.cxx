// pybind11 extension module `ptychofft_pb`: the reference's compiled-module surface
// (/root/reference/src/cuda/pybind11/ptychofft.cxx:8-26 -- class `ptychofft`, keyword constructor
// (ptheta, nz, n, nscan, detector_shape, probe_shape), six read-only attributes, fwd / adj / free on
// raw device addresses) bound to the C ABI of include/ptychofft_b200.h instead of the C++ class of
// src/include/ptychofft.cuh.  INTEGRATION.md option B, compiled by __graft_entry__.build() with
// plain g++ and exercised by tests/test_gpu_bindings.py: a reference maintainer keeps pybind11 and
// swaps `target_link_libraries(... cudart cufft)` for `ptychofft_b200`.
//
// Like the reference's module, work goes to the legacy default stream (stream = NULL), which is
// what CuPy uses; unlike it, CUDA failures raise RuntimeError (the reference checks nothing).
#include <pybind11/pybind11.h>

#include <stdexcept>
#include <string>

#include "ptychofft_b200.h"

namespace py = pybind11;

namespace {

struct Plan {
  ptx_plan* h = nullptr;
  Plan(size_t ptheta, size_t nz, size_t n, size_t nscan, size_t ndet, size_t nprb) {
    ok(ptx_create(&h, ptheta, nz, n, nscan, ndet, nprb));
  }
  ~Plan() { ptx_destroy(h); }
  Plan(const Plan&) = delete;
  Plan& operator=(const Plan&) = delete;

  static void ok(int rc) {
    if (rc != PTX_OK)
      throw std::runtime_error("libptychofft_b200: error " + std::to_string(rc) + ": " + ptx_last_error());
  }
  size_t dim(int which) const { return ptx_dim(h, which); }
  void fwd(size_t g, size_t f, size_t scan, size_t prb) {
    ok(ptx_fwd(h, (void*)g, (const void*)f, (const void*)scan, (const void*)prb, 0, nullptr));
  }
  void adj(size_t f, size_t g, size_t scan, size_t prb, int flg) {
    ok(ptx_adj(h, (void*)f, (const void*)g, (const void*)scan, (void*)prb, 0, flg, nullptr));
  }
  void release() { ok(ptx_free(h)); }
};

}  // namespace

PYBIND11_MODULE(ptychofft_pb, m) {
  m.doc() = "pybind11 binding of libptychofft_b200 with the reference's `ptychofft` class surface";
  py::class_<Plan>(m, "ptychofft")
      .def(py::init<size_t, size_t, size_t, size_t, size_t, size_t>(), py::arg("ptheta"), py::arg("nz"),
           py::arg("n"), py::arg("nscan"), py::arg("detector_shape"), py::arg("probe_shape"))
      .def_property_readonly("ptheta", [](const Plan& p) { return p.dim(PTX_DIM_PTHETA); })
      .def_property_readonly("nz", [](const Plan& p) { return p.dim(PTX_DIM_NZ); })
      .def_property_readonly("n", [](const Plan& p) { return p.dim(PTX_DIM_N); })
      .def_property_readonly("nscan", [](const Plan& p) { return p.dim(PTX_DIM_NSCAN); })
      .def_property_readonly("ndet", [](const Plan& p) { return p.dim(PTX_DIM_NDET); })
      .def_property_readonly("nprb", [](const Plan& p) { return p.dim(PTX_DIM_NPRB); })
      .def("fwd", &Plan::fwd)
      .def("adj", &Plan::adj)
      .def("free", &Plan::release);
}
