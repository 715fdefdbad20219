"""Host-pointer throughput of the batched entry points under different label transports / widening thread counts /
chunk sizes (env is read at dpx_create, so one process can sweep them).  python tools/e2e_probe.py [frames] [passes]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
h, w = 480, 640
k = synth.intrinsics_for(h, w)
depth = np.stack([synth.make_depth(h, w, i, k) for i in range(16)])
clouds = np.stack([synth.depth_to_cloud(d, k, "rowmajor") for d in depth])
reps = frames // 16
pin_depth = torch.from_numpy(np.concatenate([depth] * reps).view(np.int16)).pin_memory()
pin_xyz = torch.from_numpy(np.concatenate([clouds] * reps)).pin_memory()
pin_out = torch.empty((frames, h * w), dtype=torch.int32).pin_memory()
ref = None


def run(kind, env):
    global ref
    for key in ("DPX_LABEL_TRANSPORT", "DPX_HOST_THREADS", "DPX_HOST_CHUNK"):
        os.environ.pop(key, None)
    os.environ.update(env)
    ex = PlaneExtractor(h, w, Config(), max_batch=frames)
    call = (lambda: ex.process_depth_batch_host_ptr(pin_depth.data_ptr(), frames, k, pin_out.data_ptr())) if kind == "depth" else \
           (lambda: ex.process_batch_host_ptr(pin_xyz.data_ptr(), frames, LAYOUT_ROWMAJOR, pin_out.data_ptr()))
    call(); call()
    t0 = time.perf_counter()
    for _ in range(passes):
        call()
    dt = time.perf_counter() - t0
    if ref is None:
        ref = pin_out.clone()
    ok = bool(torch.equal(ref, pin_out))
    print(f"{kind:5s} {str(env):70s} {passes * frames / dt:9.0f} frames/s  labels {'ok' if ok else 'DIFFER'}", flush=True)
    ex.close()


pin_out16 = torch.empty((frames, h * w), dtype=torch.int16).pin_memory()


def run_u16(env):
    """raw depth in, uint16 labels out (dpx_process_depth_batch_host_u16)"""
    for key in ("DPX_LABEL_TRANSPORT", "DPX_HOST_THREADS", "DPX_HOST_CHUNK"):
        os.environ.pop(key, None)
    os.environ.update(env)
    ex = PlaneExtractor(h, w, Config(), max_batch=frames)
    call = lambda: ex.process_depth_batch_host_ptr(pin_depth.data_ptr(), frames, k, pin_out16.data_ptr(), labels_u16=True)
    call(); call()
    t0 = time.perf_counter()
    for _ in range(passes):
        call()
    dt = time.perf_counter() - t0
    print(f"depth, u16 labels out {str(env):50s} {passes * frames / dt:9.0f} frames/s", flush=True)
    ex.close()


for c in (8, 16, 24, 34, 48, 64):
    run_u16({"DPX_HOST_CHUNK": str(c)})
run("depth", {"DPX_LABEL_TRANSPORT": "i32"})
for t in (2, 4, 8, 12, 16):
    run("depth", {"DPX_LABEL_TRANSPORT": "u16", "DPX_HOST_THREADS": str(t)})
for c in (16, 32, 128):
    run("depth", {"DPX_LABEL_TRANSPORT": "u16", "DPX_HOST_THREADS": "8", "DPX_HOST_CHUNK": str(c)})
run("depth", {"DPX_LABEL_TRANSPORT": "i32", "DPX_HOST_CHUNK": "16"})
run("xyz", {"DPX_LABEL_TRANSPORT": "i32"})
run("xyz", {"DPX_LABEL_TRANSPORT": "u16", "DPX_HOST_THREADS": "8"})
