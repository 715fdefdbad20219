"""Does running two extractors on two streams overlap the HBM-bound cell-stats kernel of one batch with the latency-bound
region growing of the other?   python tools/overlap_probe.py [frames] [steps]
Each configuration runs in a child process (DPX_STREAM_WARPS is read at dpx_create)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(frames, steps, lanes, split):
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    if os.environ.get("DPX_LIB"):
        import deplex_b200._capi as _c
        _c.LIB_PATH = os.environ["DPX_LIB"]
    from deplex_b200 import Config, PlaneExtractor, synth, LAYOUT_ROWMAJOR
    h, w = 480, 640
    dev = torch.device("cuda", 0)
    uniq = 64
    base = synth.make_batch(h, w, 0, uniq, "rowmajor")
    host = np.concatenate([base] * ((frames + uniq - 1) // uniq))[:frames]
    per = frames // split  # frames per call
    d_xyz = torch.from_numpy(host).to(dev)
    d_lab = torch.empty((frames, h * w), dtype=torch.int32, device=dev)
    exs = [PlaneExtractor(h, w, Config(), max_batch=per, device=0) for _ in range(lanes)]
    streams = [torch.cuda.Stream(dev) for _ in range(lanes)]

    def step(i):
        # one step = the whole batch, as `split` calls dealt round-robin over the lanes
        for c in range(split):
            lane = (i * split + c) % lanes
            exs[lane].process_batch_device(d_xyz[c * per:(c + 1) * per], LAYOUT_ROWMAJOR, d_lab[c * per:(c + 1) * per], streams[lane])

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    e0.record(main)
    for s in streams:
        s.wait_stream(main)
    for i in range(steps):
        step(i)
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ref = exs[0].process_batch_device(d_xyz[:per], LAYOUT_ROWMAJOR).cpu()
    ok = bool(torch.equal(ref, d_lab[:per].cpu()))
    print(json.dumps({"lib": os.path.basename(os.environ.get("DPX_LIB", "default")), "warps": os.environ.get("DPX_STREAM_WARPS", "16"), "lanes": lanes, "split": split,
                      "ms_per_step": round(ms, 4), "frames_per_s": round(frames / ms * 1e3), "labels_equal": ok}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(*map(int, sys.argv[2:6]))
        sys.exit(0)
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    plan = (("16", ((1, 1), (2, 1), (3, 1))),)
    if os.environ.get("OVERLAP_FULL"):
        plan += (("12", ((1, 1), (2, 1))), ("8", ((1, 1), (2, 1), (2, 2), (2, 4), (3, 1))))
    for warps, combos in plan:
        for lanes, split in combos:
            env = dict(os.environ, DPX_STREAM_WARPS=warps)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(frames), str(steps), str(lanes), str(split)],
                               env=env, capture_output=True, text=True, timeout=300)
            print(r.stdout.strip() or r.stderr[-400:], flush=True)
