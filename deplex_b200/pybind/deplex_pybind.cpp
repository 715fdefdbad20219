// deplex_pybind.cpp -- the Python extension module `deplex.pybind`, over the C++ drop-in layer (which is
// itself a thin owner of C-ABI handles): Python -> pybind11 -> deplex::PlaneExtractor -> dpx_* -> CUDA.
//
// Same module layout and signatures as the reference (cpp/pybind/deplex_pybind.cpp:20-24,
// plane_extraction/plane_extraction.cpp:28-37, utils/utils.cpp:29-36):
//   pybind.plane_extraction.Config(path)
//   pybind.plane_extraction.PlaneExtractor(image_height, image_width, config=Config()).process(pcd_array)
//   pybind.utils.DepthImage(image_path) .height .width .transform_to_pcd(intrinsics) .reset(image_path)
// The reference converts arguments with pybind11/eigen.h; Eigen is not a dependency here, so arrays are taken
// as py::array_t<float> with forcecast (any numeric (N,3) array is converted to float32, like the Eigen caster).
// C-contiguous and Fortran-contiguous inputs both go to the device as they are (layout flag), no host transpose.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <deplex/deplex.h>

namespace py = pybind11;

#include <map>
#include <mutex>

#include "deplex_b200.h"

namespace {

// Page-locked host memory for the arrays this module hands out (transform_to_pcd's points, process()'s labels): copies
// between pinned memory and the device run at full PCIe speed and without the driver's staging, which is worth ~200 us per
// 640x480 frame.  Blocks are recycled through a small pool (cudaHostAlloc itself is slow); without a CUDA device the
// arrays are ordinary numpy arrays.  Blocks still in the pool at interpreter exit are left to the OS.
class PinnedPool {
 public:
  void* get(size_t bytes, size_t* granted) {
    const size_t want = (bytes + kGranule - 1) / kGranule * kGranule;
    {
      std::lock_guard<std::mutex> lock(mu_);
      auto it = free_.find(want);
      if (it != free_.end()) {
        void* p = it->second;
        free_.erase(it);
        *granted = want;
        return p;
      }
    }
    void* p = nullptr;
    if (dpx_host_alloc(&p, want) != DPX_OK) return nullptr;
    *granted = want;
    return p;
  }
  void put(void* p, size_t granted) {
    std::lock_guard<std::mutex> lock(mu_);
    free_.emplace(granted, p);
  }

 private:
  static constexpr size_t kGranule = 1 << 16;
  std::mutex mu_;
  std::multimap<size_t, void*> free_;
};

PinnedPool& pool() {
  static PinnedPool* p = new PinnedPool();  // never destroyed: capsules may outlive static destruction
  return *p;
}

struct PinnedBlock {
  void* ptr;
  size_t granted;
};

template <class T>
py::array_t<T> pinned_array(std::vector<py::ssize_t> shape) {
  size_t n = 1;
  for (py::ssize_t d : shape) n *= static_cast<size_t>(d);
  size_t granted = 0;
  void* p = n ? pool().get(n * sizeof(T), &granted) : nullptr;
  if (!p) return py::array_t<T>(shape);
  auto* blk = new PinnedBlock{p, granted};
  py::capsule owner(blk, [](void* b) {
    auto* blk = static_cast<PinnedBlock*>(b);
    pool().put(blk->ptr, blk->granted);
    delete blk;
  });
  return py::array_t<T>(shape, static_cast<T*>(p), owner);
}

using deplex::PlaneExtractor;
using deplex::PointLayout;
using deplex::config::Config;
using deplex::utils::DepthImage;

py::array_t<int32_t> process(PlaneExtractor& self, py::array pcd_array) {
  // any numeric array -> float32, keeping C or Fortran order when the input already has one of them
  py::array_t<float> a;
  PointLayout layout = PointLayout::RowMajor;
  const bool f_order = (pcd_array.flags() & py::array::f_style) && !(pcd_array.flags() & py::array::c_style);
  if (f_order) {
    a = py::array_t<float, py::array::f_style | py::array::forcecast>::ensure(pcd_array);
    layout = PointLayout::ColMajor;
  } else {
    a = py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(pcd_array);
  }
  if (!a) throw py::type_error("process(): pcd_array must be convertible to a float32 array");
  // Eigen::MatrixX3f semantics: two dimensions with three columns; rows() is what the size check sees
  if (a.ndim() != 2 || a.shape(1) != 3) throw py::type_error("process(): incompatible function arguments: pcd_array must have shape (N, 3)");
  const int64_t n = static_cast<int64_t>(a.shape(0));
  py::array_t<int32_t> labels = pinned_array<int32_t>({static_cast<py::ssize_t>(n)});
  const float* src = a.data();
  int32_t* dst = labels.mutable_data();
  {
    py::gil_scoped_release release;
    self.process(src, n, layout, dst);
  }
  return labels;
}

py::array_t<float> transform_to_pcd(const DepthImage& self, py::array_t<float, py::array::c_style | py::array::forcecast> intrinsics) {
  if (intrinsics.ndim() != 2 || intrinsics.shape(0) != 3 || intrinsics.shape(1) != 3)
    throw py::type_error("transform_to_pcd(): intrinsics must have shape (3, 3)");
  deplex::utils::Intrinsics k;
  for (int i = 0; i < 9; ++i) k[i] = intrinsics.data()[i];
  std::vector<float> pts = self.toPointCloudRowMajor(k);
  const py::ssize_t n = static_cast<py::ssize_t>(self.getWidth()) * self.getHeight();
  // The reference returns an Eigen::MatrixX3f (cpp/pybind/utils/utils.cpp:34), which pybind11 hands to numpy as a
  // Fortran-ordered (N, 3) array: X[N] Y[N] Z[N] in memory.  Same here, so that arr.flags, ravel(order="K") and
  // the column-major path of process() behave as with the reference.
  py::array_t<float> flat = pinned_array<float>({static_cast<py::ssize_t>(3) * n});
  float* dst = flat.mutable_data();
  for (py::ssize_t i = 0; i < n; ++i) {
    dst[i] = pts[3 * i];
    dst[n + i] = pts[3 * i + 1];
    dst[2 * n + i] = pts[3 * i + 2];
  }
  py::array_t<float> out(std::vector<py::ssize_t>{n, 3},
                         std::vector<py::ssize_t>{static_cast<py::ssize_t>(sizeof(float)), static_cast<py::ssize_t>(sizeof(float)) * n},
                         dst, flat);
  return out;
}

}  // namespace

PYBIND11_MODULE(pybind, m) {
  m.doc() = "deplex plane extraction on NVIDIA B200 (drop-in for the reference's deplex.pybind)";

  py::module m_plane_extraction = m.def_submodule("plane_extraction", "Module with Plane Extraction algorithm");
  py::class_<Config>(m_plane_extraction, "Config")
      .def(py::init<std::string>(), py::arg("path"))
      // additions: keyword-less default construction and field access (the reference exposes neither)
      .def(py::init<>())
      .def_readwrite("patch_size", &Config::patch_size)
      .def_readwrite("histogram_bins_per_coord", &Config::histogram_bins_per_coord)
      .def_readwrite("min_cos_angle_merge", &Config::min_cos_angle_merge)
      .def_readwrite("max_merge_dist", &Config::max_merge_dist)
      .def_readwrite("min_region_growing_candidate_size", &Config::min_region_growing_candidate_size)
      .def_readwrite("min_region_growing_cells_activated", &Config::min_region_growing_cells_activated)
      .def_readwrite("min_region_planarity_score", &Config::min_region_planarity_score)
      .def_readwrite("depth_sigma_coeff", &Config::depth_sigma_coeff)
      .def_readwrite("depth_sigma_margin", &Config::depth_sigma_margin)
      .def_readwrite("min_pts_per_cell", &Config::min_pts_per_cell)
      .def_readwrite("depth_discontinuity_threshold", &Config::depth_discontinuity_threshold)
      .def_readwrite("max_number_depth_discontinuity", &Config::max_number_depth_discontinuity)
      .def_readwrite("ransac_refinement", &Config::ransac_refinement)
      .def_readwrite("ransac_max_iterations", &Config::ransac_max_iterations)
      .def_readwrite("ransac_threshold", &Config::ransac_threshold)
      .def_readwrite("ransac_inliers_ratio", &Config::ransac_inliers_ratio);

  py::class_<PlaneExtractor>(m_plane_extraction, "PlaneExtractor")
      .def(py::init<int, int, Config>(), py::arg("image_height"), py::arg("image_width"), py::arg("config") = Config())
      .def("process", &process, py::arg("pcd_array"))
      // addition: plane parameters of the last processed frame as (normal[3], d, n_points, merge_label) tuples
      .def("planes", [](PlaneExtractor& self) {
        py::list out;
        for (const deplex::PlaneParams& p : self.planes())
          out.append(py::make_tuple(py::make_tuple(p.normal[0], p.normal[1], p.normal[2]), p.d, p.n_points, p.merge_label));
        return out;
      });

  py::module m_utils = m.def_submodule("utils", "Plane Extraction utilities");
  py::class_<DepthImage>(m_utils, "DepthImage")
      .def(py::init<std::string>(), py::arg("image_path"))
      .def_property_readonly("height", &DepthImage::getHeight)
      .def_property_readonly("width", &DepthImage::getWidth)
      .def("transform_to_pcd", &transform_to_pcd, py::arg("intrinsics"))
      .def("reset", &DepthImage::reset, py::arg("image_path"));
}
