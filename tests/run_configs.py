"""tests/run_configs.py — runs BASELINE.json's five configurations (C1..C5) on 1..8 GPUs.

    python tests/run_configs.py --config c1,c3            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_configs.py --config c3,c4,c5            # one rank per GPU

Per configuration: device-resident frames/s (CUDA events, max over ranks), Mpixels/s, Mtriangles/s (input
triangles), per-stage time, fraction of the HBM roofline for B_alg = 12V + 28A + 8I + 4WH, a parity check
of sample frames against the oracle (checker only), and — rank 0, bounded sample — the CPU reference.
C2/C5 are frame-parallel across ranks; C3/C4 are screen-band partitioned with an NCCL all-gather per frame.
One JSON line per configuration (also appended to gpurun_out/configs.jsonl).
(Lives under tests/: it uses the oracle as the checker of its sample frames.)"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from swift3drenderer_b200 import assets, multigpu, renderer as R, scene as S  # noqa: E402

PEAK_GBS = 6549.8
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK_GBS = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def small_path(n: int) -> np.ndarray:
    """Slow drift for the synthetic fields: creep forward, pan a little (keeps the field in view)."""
    inp = np.zeros(n, S.INPUT_DTYPE)
    for f in range(n):
        inp[f]["up"] = 1.0
        inp[f]["mouse"] = (0.5 * f, 0.2 * f)
    return inp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1,c2,c4,c5")
    ap.add_argument("--frames", type=int, default=0, help="override the frame count of every config")
    ap.add_argument("--c3-solids", type=int, default=1_000_000)
    ap.add_argument("--c4-solids", type=int, default=50_000)
    ap.add_argument("--c5-views", type=int, default=4096)
    ap.add_argument("--partition", default="fused", choices=["fused", "rows", "bands"],
                    help="multi-GPU split of a frame: interleaved tile rows stored straight into every rank's frame over "
                         "NVLink peer memory (fused, general-path scenes), interleaved tile rows + all-gather (rows), or "
                         "contiguous bands + all-gather (bands)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def shared_scene_file(name: str, build):
        """rank 0 writes the data.bin once to /dev/shm, every rank loads it."""
        path = f"/dev/shm/s3r_{name}.data.bin"
        if rank == 0 and not os.path.exists(path):
            S.write_data_bin(path, build())
        barrier()
        return path

    r = R.Renderer(local)
    results = []

    def run(name, path, inputs, W, H, mode, parity_frames, cpu_frames, note=""):
        r.load_scene_file(path)
        counts = {}
        import ctypes
        v, i, a, t = (ctypes.c_uint64() for _ in range(4))
        R.load_library().s3r_scene_counts(r._h, ctypes.byref(v), ctypes.byref(i), ctypes.byref(a), ctypes.byref(t))
        counts = {"V": v.value, "I": i.value, "A": a.value, "T": i.value // 3}
        mats = R.camera_path(inputs)
        n = len(mats)
        b_alg = 12 * counts["V"] + 28 * counts["A"] + 8 * counts["I"] + 4 * W * H
        line = {"config": name, "gpus": world, "mode": mode if mode == "frames" else args.partition, "W": W, "H": H, "frames": n, **counts, "note": note}

        if mode == "frames":  # frame-parallel: this rank renders its share of the poses, several per launch
            mine = list(multigpu.frame_shard(n, rank, world))
            vpl = 64 if W * H <= 1024 * 1024 else 8   # poses per launch set (bench.py renders C2 the same way)
            out = torch.empty((2, vpl, H, W), dtype=torch.int32, device=dev)

            def step():
                k = 0
                for f0 in range(0, len(mine), vpl):
                    idx = mine[f0:f0 + vpl]
                    r.render_device(mats[idx], W, H, out[k % out.shape[0]].data_ptr(), stream=stream.cuda_stream)
                    k += 1
            frames_done = n
        else:  # screen bands + all-gather of the assembled frame on every rank
            y0, y1 = multigpu.band_edges(H, world)[rank]
            full = [None]
            if args.partition == "fused" and world > 1:  # no pixel collective: rows go straight into every rank's frame
                pf = multigpu.PeerFrames(r, H, W, rank, world, dev, ring=2)
                last = [0]

                def step():
                    for f in range(n):
                        last[0] = pf.render(mats[f], stream=stream.cuda_stream)
                        pf.fence()

                def read_own_slot():   # the assembled frame of the last pose, from this rank's own ring slot
                    torch.cuda.synchronize(dev)
                    return pf.read(last[0])
                full[0] = read_own_slot
            elif args.partition in ("rows", "fused") and world > 1:  # interleaved tile rows: compacted buffer, gather, de-interleave
                asm = multigpu.InterleavedAssembler(H, W, rank, world, dev, R.tile_height())

                def step():
                    for f in range(n):
                        r.render_device_rows(mats[f], W, H, world, rank, asm.mine.data_ptr(), stream=stream.cuda_stream)
                        full[0] = asm.gather()
            elif multigpu.equal_bands(H, world):  # render into the frame's own rows, gather in place
                frames = torch.zeros((4, H, W), dtype=torch.int32, device=dev)

                def step():
                    for f in range(n):
                        fr = frames[f % 4]
                        r.render_device(mats[f], W, H, fr[y0:y1].data_ptr(), y0=y0, y1=y1, stream=stream.cuda_stream)
                        full[0] = multigpu.gather_bands_inplace(fr, rank, world)
            else:
                band = torch.empty((y1 - y0, W), dtype=torch.int32, device=dev)

                def step():
                    for f in range(n):
                        r.render_device(mats[f], W, H, band.data_ptr(), y0=y0, y1=y1, stream=stream.cuda_stream)
                        full[0] = multigpu.assemble_bands(band, H, rank, world)
            frames_done = n

        for _ in range(2):  # warm-up incl. capacity regrowth (decided together: step() may contain a collective)
            step()
            while max_over_ranks(1.0 if r.finish() else 0.0) > 0:
                step()
        r.set_option("timing", 1)
        r.timing(reset=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if n <= 64 else 1
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
        barrier()
        assert not r.finish()
        if mode != "frames" and args.partition == "fused" and world > 1:
            line["assembly"] = "shading kernel stores rows into every rank's frame (CUDA IPC peer memory) + 1-element all-reduce per frame"
        ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        tm = r.timing(reset=True)
        r.set_option("timing", 0)
        fps = frames_done / (ms / 1e3)
        st = r.stats(0)
        per_frame_s = ms / 1e3 / (frames_done / (world if mode == "frames" else 1))
        line.update({
            "fps": fps, "ms_per_frame": 1e3 / fps, "mpixels_per_s": fps * W * H / 1e6,
            "mtriangles_per_s": fps * counts["T"] / 1e6,
            "geometry_ms_per_submission": tm["geometry_ms"] / max(tm["chunks"], 1),
            "raster_ms_per_submission": tm["raster_ms"] / max(tm["chunks"], 1),
            "b_alg_bytes": b_alg, "hbm_roofline_frac_per_gpu": b_alg / per_frame_s / 1e9 / PEAK_GBS,
            "stats_last_view": st, "straddle_fraction": st["clipped"] / max(counts["T"], 1),
            "survivor_fraction": st["setups"] / max(counts["T"], 1),
        })

        if not args.no_parity and rank == 0:
            from oracle import port  # checker only
            osc = port.OracleScene(path=path)
            bad = 0
            for f in parity_frames:
                t0 = time.time()
                want = osc.render(mats[f], W, H)["pixels"]
                got = r.render(mats[f], W, H)[0]
                diff = int((got != want).sum())
                bad += diff
                line.setdefault("parity", []).append({"frame": int(f), "differing_pixels": diff, "oracle_s": round(time.time() - t0, 2)})
            line["parity_ok"] = bad == 0
            if mode == "bands" and world > 1:
                pass
        if mode == "bands":
            # every rank: the assembled frame of the last pose equals a single-GPU whole-frame render
            whole = r.render(mats[n - 1], W, H)[0]
            assembled = full[0]() if callable(full[0]) else full[0].cpu().numpy().view(np.uint32)
            same = bool(np.array_equal(assembled, whole))
            line["banded_equals_whole"] = bool(max_over_ranks(0.0 if same else 1.0) == 0.0)
        if not args.no_cpu and rank == 0 and cpu_frames:
            from oracle import port
            osc = port.OracleScene(path=path)
            t0 = time.perf_counter()
            for f in cpu_frames:
                osc.render(mats[f], W, H)
            dt = time.perf_counter() - t0
            line["cpu_port_1core_fps"] = len(cpu_frames) / dt
            line["cpu_sample"] = f"{len(cpu_frames)} frames, oracle C port, 1 core"
        if rank == 0:
            print(json.dumps(line), flush=True)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as fh:
                fh.write(json.dumps(line) + "\n")
        results.append(line)
        barrier()

    todo = args.config.split(",")
    shipped = assets.ensure_shipped_data_bin()
    if "c1" in todo:
        n = args.frames or 300
        run("C1 data.bin scene 1280x720, 300-frame path", shipped, S.input_script("c1_path", n), 1280, 720, "frames",
            [0, 59, 119, 200, n - 1], list(range(0, n, max(1, n // 12))))
    if "c2" in todo:
        n = args.frames or 600
        run("C2 data.bin scene 3840x2160, 600-frame fly-through", shipped, S.input_script("flythrough", n), 3840, 2160,
            "frames", [0, 100, 215, 330, 470], list(range(0, n, max(1, n // 8))))
    if "c3" in todo:
        n = args.frames or 8
        path = shared_scene_file(f"c3_{args.c3_solids}", lambda: S.c3_scene(args.c3_solids))
        run(f"C3 {args.c3_solids} textured icosahedrons, 4K, screen bands", path, small_path(n), 3840, 2160, "bands",
            [0], [0], note="camera at the origin, field centred at z = -4000")
    if "c4" in todo:
        n = args.frames or 16
        path = shared_scene_file(f"c4_{args.c4_solids}", lambda: S.clip_stress_scene(args.c4_solids))
        run(f"C4 clipping stress, {args.c4_solids} icosahedrons, 4K, screen bands", path, S.input_script("strafe", n), 3840, 2160,
            "bands", [0, n - 1], [0])
    if "c5" in todo:
        n = args.frames or args.c5_views
        run(f"C5 {n} camera poses at 512x512, frame-parallel", shipped, S.input_script("spin", n), 512, 512, "frames",
            [0, 700, n - 1], list(range(0, n, max(1, n // 64))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
