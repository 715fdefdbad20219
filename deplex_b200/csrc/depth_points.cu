// depth_points.cu -- DepthImage::toPointCloud (cpp/deplex/src/deplex/utils/depth_image.cpp:55-78) as a kernel:
// z = float(raw), x = ((col - cx) * z) / fx, y = ((row - cy) * z) / fy, each step rounded to fp32 (the reference
// evaluates the Eigen array expression left to right, SSE2 mulps/divps).  Used when the fused path of the
// cell-stats kernel cannot take the geometry (odd patch size, unaligned rows) or when the refinement stage needs
// the points again; 2 B/pixel in, 12 B/pixel out.
#include "depth_points.cuh"

namespace dpx {
namespace {

__global__ void __launch_bounds__(256) depth_to_points_kernel(const uint16_t* __restrict__ depth, long long total, int width,
                                                              long long n_points, Pinhole k, float* __restrict__ xyz) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long pix = i % n_points;
  const int row = static_cast<int>(pix / width), col = static_cast<int>(pix - static_cast<long long>(row) * width);
  const float z = static_cast<float>(depth[i]);
  xyz[3 * i + 0] = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(col), k.cx), z), k.fx);
  xyz[3 * i + 1] = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(row), k.cy), z), k.fy);
  xyz[3 * i + 2] = z;
}

}  // namespace

cudaError_t launch_depth_to_points(const uint16_t* depth, int n_frames, const Geometry& g, const Pinhole& k, float* xyz,
                                   cudaStream_t stream) {
  const long long total = static_cast<long long>(n_frames) * g.n_points;
  if (total == 0) return cudaSuccess;
  depth_to_points_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(depth, total, g.width, g.n_points, k, xyz);
  return cudaGetLastError();
}

}  // namespace dpx
