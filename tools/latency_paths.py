"""Per-call latency of process() on the shipped TUM frame through the different host entry points:
pybind (numpy, pageable), ctypes mirror (numpy, pageable), pinned host pointers, device-resident."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deplex_b200", "python"))
import deplex
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
g = os.path.join(ROOT, "tests", "golden")
depth = np.load(os.path.join(g, "tum_depth.npz"))["depth"]
K = np.loadtxt(os.path.join(g, "TUM_fr3_long_val.K"), dtype=np.float32)
k = dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
xyz = synth.depth_to_cloud(depth, k, "rowmajor")
def bench(fn, n=200):
    for _ in range(10): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    return f"min {ts.min():7.1f}  mean {ts.mean():7.1f}  max {ts.max():7.1f} us"
alg = deplex.PlaneExtractor(480, 640)
print("pybind  process(float32 C-order numpy)   ", bench(lambda: alg.process(xyz)))
x64 = xyz.astype(np.float64)
print("pybind  process(float64 numpy, converted)", bench(lambda: alg.process(x64), 50))
ex = PlaneExtractor(480, 640, Config())
print("ctypes  process(float32 numpy)           ", bench(lambda: ex.process(xyz)))
pin = torch.from_numpy(xyz).pin_memory(); lab = torch.empty(480 * 640, dtype=torch.int32).pin_memory()
print("C-ABI   pinned host pointers             ", bench(lambda: ex.process_batch_host_ptr(pin.data_ptr(), 1, LAYOUT_ROWMAJOR, lab.data_ptr())))
d16 = torch.from_numpy(depth.view(np.int16)).pin_memory()
print("C-ABI   pinned raw depth                 ", bench(lambda: ex.process_depth_batch_host_ptr(d16.data_ptr(), 1, k, lab.data_ptr())))
import cv2, tempfile
png = os.path.join(tempfile.mkdtemp(), "tum.png"); cv2.imwrite(png, depth)
img = deplex.utils.DepthImage(png)
pts = img.transform_to_pcd(K)   # pinned-backed when a GPU is present
print("pybind  process(transform_to_pcd output) ", bench(lambda: alg.process(pts)))
assert np.array_equal(alg.process(pts), alg.process(xyz))
print("pybind  transform_to_pcd                 ", bench(lambda: img.transform_to_pcd(K), 50))
