"""Fused frame assembly across processes: one process per GPU, NCCL for the ordering fence, CUDA IPC peer
memory for the pixels (swift3drenderer_b200.multigpu.PeerFrames).  Needs >= 2 GPUs on one node; skipped
otherwise (the single-GPU mechanics are covered by
tests/test_gpu_parity.py::test_fused_assembly_writes_rows_to_every_destination)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from swift3drenderer_b200 import multigpu, renderer as R, scene as S
    sc = S.icosahedron_field(20000, seed=9, extent=40, r_range=(0.3, 1.0))
    r = R.Renderer(rank)
    r.load_scene(sc)
    mats = R.camera_path(S.input_script("spin", 6))
    W, H = 800, 450
    pf = multigpu.PeerFrames(r, H, W, rank, world, dev, ring=2)
    ok = True
    for f in range(6):       # capacity regrowth happens here (a regrowth on any rank repeats the frame on all of them)
        for attempt in range(4):
            slot = pf.render(mats[f])
            flag = torch.tensor([1 if r.finish() else 0], device=dev)
            dist.all_reduce(flag)
            pf.fence()
            if flag.item() == 0:
                break
            pf.count -= 1
    for f in range(6):       # checked pass: nothing but the fence orders this rank's reads after its peers' stores
        slot = pf.render(mats[f])
        pf.fence()
        torch.cuda.current_stream(dev).synchronize()   # my fence has completed, hence every rank's render before its fence
        got = pf.read(slot)
        assert not r.finish()
        want = r.render(mats[f], W, H)[0]
        ok = ok and bool(np.array_equal(got, want))
        dist.barrier()
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.asarray([ok]))
    pf.close()
    r.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_frames_assemble_the_whole_frame_on_every_rank(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs on one node")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for k in range(world):
        assert np.load(tmp_path / f"ok_{k}.npy")[0], f"rank {k}: fused frame differs from the single-GPU frame"
