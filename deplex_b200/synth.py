"""Synthetic organized clouds: random piecewise-planar scenes with depth noise and holes.

Shared by the parity tests, smoke() and bench.py, so that the CUDA path and the CPU oracle always see
the same arrays.  This is workload generation, not part of the hot path: plain numpy on the host.

Scene model:
  * the image is split by a random BSP of 2-7 lines into regions, each carrying its own plane
    n.X = d with the normal within 60 degrees of the optical axis and 0.8-6 m depth;
  * raw depth = metres * 5000 (TUM convention, so the shipped thresholds stay meaningful), plus
    Gaussian noise sigma = 0.1 * 1.425e-6 * raw^2 (half the Kinect model 1.425e-6 * z_mm^2 expressed
    in raw units of 0.2 mm), rounded to uint16;
  * holes: 3-6 random elliptical blobs (about 5-15 % of the pixels) and sparse salt (0.05 % of the
    pixels) set to 0.  With the reference's default minPtsPerCell=3 a single zero pixel invalidates
    its whole cell (cell_segment.cpp:23), so denser salt would leave no planar cell at all;
  * back-projection exactly as DepthImage::toPointCloud (depth_image.cpp:64-73), fp32, left to right.
Frame i of a sequence uses seed SEED0 + i with numpy's counter-based Philox generator.
"""
import numpy as np

SEED0 = 0xD3B1E5
TUM_K = dict(fx=535.4, fy=539.2, cx=320.1, cy=247.6)  # data/configs/TUM_fr3_long_val.K


def intrinsics_for(height, width):
    """TUM intrinsics scaled with the image width (640 -> x1, 1280 -> x2, 1920 -> x3)."""
    s = width / 640.0
    return dict(fx=TUM_K["fx"] * s, fy=TUM_K["fy"] * s, cx=TUM_K["cx"] * s, cy=TUM_K["cy"] * s)


def make_depth(height, width, frame_index, k=None):
    """uint16 (H,W) raw depth of synthetic frame `frame_index`."""
    k = k or intrinsics_for(height, width)
    rng = np.random.Generator(np.random.Philox(key=SEED0 + int(frame_index)))
    v, u = np.mgrid[0:height, 0:width].astype(np.float32)
    rx = (u - np.float32(k["cx"])) / np.float32(k["fx"])
    ry = (v - np.float32(k["cy"])) / np.float32(k["fy"])

    # BSP: each line splits one existing region in two
    region = np.zeros((height, width), dtype=np.int32)
    n_lines = int(rng.integers(2, 8))
    for i in range(n_lines):
        target = int(rng.integers(0, i + 1))
        px, py = rng.uniform(0.2, 0.8) * width, rng.uniform(0.2, 0.8) * height
        ang = rng.uniform(0, np.pi)
        side = (u - px) * np.float32(np.cos(ang)) + (v - py) * np.float32(np.sin(ang)) > 0
        region[(region == target) & side] = i + 1
    n_regions = n_lines + 1

    depth = np.zeros((height, width), dtype=np.float64)
    for r in range(n_regions):
        mask = region == r
        if not mask.any():
            continue
        # normal within 60 degrees of the optical axis, facing the camera
        tilt = rng.uniform(0, np.pi / 3)
        az = rng.uniform(0, 2 * np.pi)
        n = np.array([np.sin(tilt) * np.cos(az), np.sin(tilt) * np.sin(az), -np.cos(tilt)])
        z0 = rng.uniform(0.8, 6.0)  # metres along the ray through the region centroid
        cu, cv = u[mask].mean(), v[mask].mean()
        rc = np.array([(cu - k["cx"]) / k["fx"], (cv - k["cy"]) / k["fy"], 1.0])
        d = float(n @ (rc * z0))
        denom = n[0] * rx[mask].astype(np.float64) + n[1] * ry[mask].astype(np.float64) + n[2]
        z = d / denom
        z[(z < 0.3) | (z > 12.0) | ~np.isfinite(z)] = 0.0
        depth[mask] = z
    raw = depth * 5000.0
    sigma = 0.1 * 1.425e-6 * raw * raw
    raw = raw + rng.standard_normal(raw.shape) * sigma
    raw[depth == 0.0] = 0.0
    raw = np.clip(np.rint(raw), 0, 65535)

    # holes: elliptical blobs + sparse salt
    for _ in range(int(rng.integers(3, 7))):
        ex, ey = rng.uniform(0, width), rng.uniform(0, height)
        ax, ay = rng.uniform(0.04, 0.12) * width, rng.uniform(0.04, 0.12) * height
        raw[((u - ex) / ax) ** 2 + ((v - ey) / ay) ** 2 < 1.0] = 0
    raw[rng.random(raw.shape) < 5e-4] = 0
    return raw.astype(np.uint16)


def depth_to_cloud(depth_u16, k, layout="rowmajor"):
    """DepthImage::toPointCloud: z = float(raw); x = (col - cx) * z / fx; y = (row - cy) * z / fy (fp32)."""
    h, w = depth_u16.shape
    z = depth_u16.astype(np.float32)
    cols = np.arange(w, dtype=np.float32)[None, :]
    rows = np.arange(h, dtype=np.float32)[:, None]
    x = (cols - np.float32(k["cx"])) * z / np.float32(k["fx"])
    y = (rows - np.float32(k["cy"])) * z / np.float32(k["fy"])
    if layout == "rowmajor":
        return np.stack([x, y, z], axis=-1).reshape(h * w, 3)
    return np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], axis=0)  # (3, N): column-major N x 3


def make_cloud(height, width, frame_index, layout="rowmajor", k=None):
    k = k or intrinsics_for(height, width)
    return depth_to_cloud(make_depth(height, width, frame_index, k), k, layout)


def make_batch(height, width, first_frame, n_frames, layout="rowmajor", out=None):
    """(F,N,3) row-major or (F,3,N) column-major float32 batch of frames first_frame .. first_frame+F-1."""
    n = height * width
    shape = (n_frames, n, 3) if layout == "rowmajor" else (n_frames, 3, n)
    if out is None:
        out = np.empty(shape, dtype=np.float32)
    k = intrinsics_for(height, width)
    for i in range(n_frames):
        out[i] = make_cloud(height, width, first_frame + i, layout, k)
    return out
