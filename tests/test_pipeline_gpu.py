"""GPU tests of the orchestration layers of the C-ABI (include/deplex_b200.h, csrc/pipeline.cu) and of the host path's
label transport: dpx_pipeline (batches in flight on one GPU), dpx_sequence (a host frame range sharded over the visible
GPUs from one process), uint16 labels over PCIe, unaligned label buffers, and the final gather on CUDA tensors."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, to_oracle_cfg

pytestmark = pytest.mark.gpu


def _batch(h, w, first, n):
    from deplex_b200 import synth
    k = synth.intrinsics_for(h, w)
    depth = np.stack([synth.make_depth(h, w, first + i, k) for i in range(n)])
    clouds = np.stack([synth.depth_to_cloud(d, k, "rowmajor") for d in depth])
    return depth, clouds, k


@pytest.mark.parametrize("transport", ["i32", "u16"])
@pytest.mark.parametrize("n_frames", [2, 17, 70])
def test_host_path_label_transport_matches_oracle(oracle_mod, transport, n_frames):
    """dpx_process_batch_host / dpx_process_depth_batch_host with int32 and uint16 label transport, across the chunk
    schedules (too short to taper, tapered, several full chunks): identical labels, identical to the oracle."""
    from deplex_b200 import Config, PlaneExtractor, LAYOUT_ROWMAJOR
    h, w = 480, 640
    depth, clouds, k = _batch(h, w, 5100, min(n_frames, 6))
    reps = (n_frames + len(depth) - 1) // len(depth)
    depth, clouds = np.concatenate([depth] * reps)[:n_frames], np.concatenate([clouds] * reps)[:n_frames]
    cfg = Config()
    ex = PlaneExtractor(h, w, cfg, max_batch=64)
    ex.set_label_transport(transport)
    got_pts = ex.process_batch_host(clouds, LAYOUT_ROWMAJOR)
    got_depth = ex.process_depth_batch_host(depth, k)
    ref = oracle_mod.process_batch(h, w, to_oracle_cfg(oracle_mod, cfg), clouds[: min(n_frames, 6)], 1, os.cpu_count() or 1)
    ref = np.concatenate([ref] * reps)[:n_frames]
    assert np.array_equal(got_pts, ref)
    assert np.array_equal(got_depth, ref)
    # a second call reuses the staging buffers and the widening threads
    assert np.array_equal(ex.process_depth_batch_host(depth, k), ref)
    # the uint16-output entry points: the same values, narrow
    l16 = ex.process_depth_batch_host_u16(depth, k)
    assert l16.dtype == np.uint16 and np.array_equal(l16, ref)
    assert np.array_equal(ex.process_batch_host_u16(clouds, LAYOUT_ROWMAJOR), ref)
    assert np.array_equal(ex.process_depth_batch_host_u16(depth[:1], k), ref[:1])


def test_env_chunk_override_and_small_max_batch(oracle_mod, monkeypatch):
    from deplex_b200 import Config, PlaneExtractor, LAYOUT_ROWMAJOR
    h, w = 480, 640
    depth, clouds, k = _batch(h, w, 5200, 5)
    ref = oracle_mod.process_batch(h, w, oracle_mod.OracleConfig(), clouds, 1, 4)
    monkeypatch.setenv("DPX_HOST_CHUNK", "2")
    monkeypatch.setenv("DPX_LABEL_TRANSPORT", "u16")
    monkeypatch.setenv("DPX_HOST_THREADS", "3")
    ex = PlaneExtractor(h, w, Config(), max_batch=3)
    assert np.array_equal(ex.process_batch_host(clouds, LAYOUT_ROWMAJOR), ref)
    assert np.array_equal(ex.process_depth_batch_host(depth, k), ref)


def test_unaligned_label_pointer(oracle_mod):
    """d_labels that is only 4-byte aligned must take the scalar stores of both painters (the fused one in
    the region-growing kernel and the labeling kernel), not fault on st.global.v4."""
    import torch
    from deplex_b200 import Config, PlaneExtractor, LAYOUT_ROWMAJOR
    h, w, F = 480, 640, 20  # 20 frames: 18 painted by the region-growing CTAs, 2 by the labeling kernel
    depth, clouds, k = _batch(h, w, 5300, 4)
    clouds = np.concatenate([clouds] * 5)
    ref = np.concatenate([oracle_mod.process_batch(h, w, oracle_mod.OracleConfig(), clouds[:4], 1, 4)] * 5)
    ex = PlaneExtractor(h, w, Config(), max_batch=F)
    d_xyz = torch.from_numpy(clouds).cuda()
    for off in (1, 2, 3):
        buf = torch.full((F * h * w + 4,), -7, dtype=torch.int32, device="cuda")
        out = buf[off:off + F * h * w].view(F, h * w)
        ex.process_batch_device(d_xyz, LAYOUT_ROWMAJOR, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref), off
        assert int(buf[off - 1].item()) == -7 and int(buf[off + F * h * w].item()) == -7  # nothing written outside
    d_depth = torch.from_numpy(np.concatenate([depth] * 5).view(np.int16)).cuda()
    buf = torch.empty(F * h * w + 1, dtype=torch.int32, device="cuda")
    out = ex.process_depth_batch_device(d_depth, k, buf[1:].view(F, h * w))
    assert np.array_equal(out.cpu().numpy(), ref)


def test_pipeline_capi_device_pointers_and_errors(oracle_mod):
    """dpx_pipeline_* straight through ctypes: raw device pointers, an explicit producer stream, join on another stream,
    and the error paths (batch too large, bad lane count, the reference's constructor error)."""
    import torch
    from deplex_b200 import _capi, Config, LAYOUT_ROWMAJOR
    lib = _capi.load()
    h, w, F = 480, 640, 6
    _, clouds, _ = _batch(h, w, 5400, F)
    ref = oracle_mod.process_batch(h, w, oracle_mod.OracleConfig(), clouds, 1, 4)
    cfg = Config()
    p = C.c_void_p()
    assert lib.dpx_pipeline_create(h, w, C.byref(cfg._c), -1, F, 2, C.byref(p)) == _capi.DPX_OK
    assert lib.dpx_pipeline_lanes(p) == 2 and lib.dpx_pipeline_lane(p, 1) and not lib.dpx_pipeline_lane(p, 2)
    producer, consumer = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(producer):
        d_xyz = torch.from_numpy(clouds).cuda(non_blocking=True)
    outs = [torch.empty((F, h * w), dtype=torch.int32, device="cuda") for _ in range(5)]
    for o in outs:
        assert lib.dpx_pipeline_submit_device(p, d_xyz.data_ptr(), F, LAYOUT_ROWMAJOR, o.data_ptr(), producer.cuda_stream) == 0
    assert lib.dpx_pipeline_join(p, consumer.cuda_stream) == 0
    with torch.cuda.stream(consumer):
        got = [o.cpu() for o in outs]
    consumer.synchronize()
    for g in got:
        assert np.array_equal(g.numpy(), ref)
    assert lib.dpx_pipeline_kernel_launches(p) >= 3 * 5
    assert lib.dpx_pipeline_submit_device(p, d_xyz.data_ptr(), F + 1, LAYOUT_ROWMAJOR, outs[0].data_ptr(), None) == _capi.DPX_ERR_ARGUMENT
    assert b"exceeds max_batch" in lib.dpx_pipeline_last_error(p)
    assert lib.dpx_pipeline_synchronize(p) == 0
    lib.dpx_pipeline_destroy(p)
    q = C.c_void_p()
    assert lib.dpx_pipeline_create(h, w, C.byref(cfg._c), -1, F, 0, C.byref(q)) == _capi.DPX_ERR_ARGUMENT
    bad = Config(patch_size=0)
    assert lib.dpx_pipeline_create(h, w, C.byref(bad._c), -1, F, 2, C.byref(q)) == _capi.DPX_ERR_RUNTIME
    assert b"patchSize(0)" in lib.dpx_pipeline_last_error(None)


def test_sequence_extractor_matches_frame_by_frame(oracle_mod):
    """dpx_sequence_*: a host frame range sharded over every visible GPU from one process (one worker thread per GPU)
    gives the labels of process() frame by frame, for clouds and raw depth, including ranges that do not divide."""
    import torch
    from deplex_b200 import Config, SequenceExtractor, LAYOUT_ROWMAJOR
    h, w, n = 480, 640, 11
    depth, clouds, k = _batch(h, w, 5500, n)
    cfg = Config()
    ref = oracle_mod.process_batch(h, w, to_oracle_cfg(oracle_mod, cfg), clouds, 1, os.cpu_count() or 1)
    seq = SequenceExtractor(h, w, cfg, max_batch=4)
    assert seq.n_devices == torch.cuda.device_count()
    ranges = [seq.frame_range(n, g) for g in range(seq.n_devices)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert np.array_equal(seq.process_host(clouds, LAYOUT_ROWMAJOR), ref)
    assert np.array_equal(seq.process_depth_host(depth, k), ref)
    assert seq.process_host(clouds[:0], LAYOUT_ROWMAJOR).shape == (0, h * w)
    seq.close()
    # an explicit device list: the same device twice behaves like two GPUs (two handles, two worker threads)
    seq = SequenceExtractor(h, w, cfg, devices=[0, 0], max_batch=4)
    assert seq.n_devices == 2 and seq.frame_range(n, 0) == (0, 6) and seq.frame_range(n, 1) == (6, 11)
    assert np.array_equal(seq.process_host(clouds, LAYOUT_ROWMAJOR), ref)
    seq.close()


def test_sequence_with_refinement_on_every_gpu(oracle_mod):
    """The refinement kernel's launch variants (16-CTA clusters need a per-device function attribute) through the sequence
    API: one worker thread per GPU in one process, chunks of 4, 2 and 1 frames.  With one GPU visible the same device is
    listed twice (two handles, two threads)."""
    import torch
    from deplex_b200 import Config, SequenceExtractor, LAYOUT_ROWMAJOR
    h, w, n = 480, 640, 7
    _, clouds, _ = _batch(h, w, 6600, n)
    cfg = Config(ransac_refinement=1, ransac_threshold=3.0, ransac_inliers_ratio=0.5, ransac_max_iterations=200)
    ref = oracle_mod.process_batch(h, w, to_oracle_cfg(oracle_mod, cfg), clouds, 1, os.cpu_count() or 1)
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0]
    seq = SequenceExtractor(h, w, cfg, devices=devices, max_batch=4)
    assert np.array_equal(seq.process_host(clouds, LAYOUT_ROWMAJOR), ref)
    seq.close()


def test_sequence_device_resident_capi(oracle_mod):
    """dpx_sequence_process_device: every device runs its resident frames through a lanes-deep pipeline."""
    import torch
    from deplex_b200 import _capi, Config, LAYOUT_ROWMAJOR
    lib = _capi.load()
    h, w, n = 480, 640, 10
    _, clouds, _ = _batch(h, w, 5600, n)
    ref = oracle_mod.process_batch(h, w, oracle_mod.OracleConfig(), clouds, 1, os.cpu_count() or 1)
    cfg = Config()
    s = C.c_void_p()
    devs = (C.c_int32 * 2)(0, 0)
    assert lib.dpx_sequence_create(h, w, C.byref(cfg._c), devs, 2, 4, C.byref(s)) == 0
    d_in = [torch.from_numpy(clouds).cuda(), torch.from_numpy(clouds[::-1].copy()).cuda()]
    d_out = [torch.empty((n, h * w), dtype=torch.int32, device="cuda") for _ in range(2)]
    ins = (C.c_void_p * 2)(*[t.data_ptr() for t in d_in])
    outs = (C.c_void_p * 2)(*[t.data_ptr() for t in d_out])
    ms = (C.c_float * 2)()
    torch.cuda.synchronize()
    assert lib.dpx_sequence_process_device(s, ins, n, LAYOUT_ROWMAJOR, outs, 2, ms) == 0, lib.dpx_sequence_last_error(s)
    assert ms[0] > 0 and ms[1] > 0
    assert np.array_equal(d_out[0].cpu().numpy(), ref)
    assert np.array_equal(d_out[1].cpu().numpy(), ref[::-1])
    lib.dpx_sequence_destroy(s)


def test_device_memory_helpers():
    from deplex_b200 import _capi
    lib = _capi.load()
    assert lib.dpx_device_count() >= 1
    p = C.c_void_p()
    assert lib.dpx_device_alloc(0, C.byref(p), 1 << 20) == 0 and p.value
    src = np.arange(1 << 18, dtype=np.int32)
    dst = np.zeros_like(src)
    assert lib.dpx_memcpy_to_device(0, p, src.ctypes.data, src.nbytes) == 0
    assert lib.dpx_memcpy_to_host(0, dst.ctypes.data, p, src.nbytes) == 0
    assert np.array_equal(src, dst)
    lib.dpx_device_free(0, p)


def test_gather_labels_on_cuda_tensors_single_rank():
    """sharding.gather_labels / process_range_device on CUDA tensors with the real extractor (world size 1 here; the
    NCCL exchange itself runs in bench.py's sharded_sequence leg under torchrun and in test_gather_labels_nccl_two_ranks
    when two GPUs are visible)."""
    import torch
    from deplex_b200 import Config, PipelinedExtractor, sharding, LAYOUT_ROWMAJOR
    h, w, U, B, T = 480, 640, 3, 4, 23
    _, clouds, _ = _batch(h, w, 5700, U)
    pool = torch.from_numpy(np.concatenate([clouds] * 3)).cuda()  # 9 >= B + U frames
    pipe = PipelinedExtractor(h, w, Config(), max_batch=B, lanes=2)
    base = pipe.lanes[0].process_batch_device(pool[:U].contiguous(), LAYOUT_ROWMAJOR).clone()
    out = torch.empty((T, h * w), dtype=torch.int32, device="cuda")
    b, e = sharding.frame_range(T, 0, 1)
    sharding.process_range_device(pipe, pool, U, b, e, LAYOUT_ROWMAJOR, B, out, torch.cuda.current_stream())
    pipe.join()
    full = sharding.gather_labels(out, T, out=out)
    assert full.data_ptr() == out.data_ptr()
    idx = torch.arange(T, device="cuda") % U
    assert bool(torch.equal(full, base[idx]))
    pipe.close()


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from deplex_b200 import Config, PipelinedExtractor, sharding, synth, LAYOUT_ROWMAJOR
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
h, w, U, B, T = 480, 640, 3, 4, 23
k = synth.intrinsics_for(h, w)
clouds = np.stack([synth.depth_to_cloud(synth.make_depth(h, w, 5700 + i, k), k, "rowmajor") for i in range(U)])
pool = torch.from_numpy(np.concatenate([clouds] * 3)).cuda()
pipe = PipelinedExtractor(h, w, Config(), max_batch=B, device=rank, lanes=2)
base = pipe.lanes[0].process_batch_device(pool[:U].contiguous(), LAYOUT_ROWMAJOR).clone()
b, e = sharding.frame_range(T, rank, world)
result = torch.empty((T, h * w), dtype=torch.int32, device="cuda") if rank == 0 else None
local = result[b:e] if rank == 0 else torch.empty((e - b, h * w), dtype=torch.int32, device="cuda")
sharding.process_range_device(pipe, pool, U, b, e, LAYOUT_ROWMAJOR, B, local, torch.cuda.current_stream())
pipe.join()
full = sharding.gather_labels(local, T, dst=0, out=result)
torch.cuda.synchronize()
if rank == 0:
    idx = torch.arange(T, device="cuda") % U
    assert bool(torch.equal(full, base[idx])), "gathered labels differ"
    print("NCCL_GATHER_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_gather_labels_nccl_two_ranks(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (bench.py --gpus N runs the same exchange in its sharded_sequence leg)")
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_GATHER_OK" in r.stdout, r.stderr[-3000:]


def test_cpp_sequence_example_all_gpus(tmp_path):
    """cpp/examples/process_sequence --clouds: deplex::SequenceExtractor over every visible GPU with no Python in the
    loop; host mode and device-resident mode must agree, and the checksum must equal the ctypes path's."""
    from deplex_b200 import Config, PlaneExtractor, LAYOUT_ROWMAJOR
    h, w, U = 480, 640, 4
    _, clouds, _ = _batch(h, w, 5800, U)
    path = tmp_path / "clouds.bin"
    clouds.tofile(path)
    exe = os.path.join(ROOT, "deplex_b200", "cpp", "build", "process_sequence")
    r = subprocess.run([exe, "--clouds", str(path), str(h), str(w), "-", "8", "2", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "device == host labels: yes" in r.stdout
    labels = PlaneExtractor(h, w, Config(), max_batch=U).process_batch_host(clouds, LAYOUT_ROWMAJOR)
    want = 0
    for f in range(U):
        i = np.arange(0, h * w, 997)
        want += int((labels[f, i].astype(np.uint64) * (1 + (i & 1023)).astype(np.uint64)).sum())
    assert f"labels checksum {want}," in r.stdout, r.stdout
