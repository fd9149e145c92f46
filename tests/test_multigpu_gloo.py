"""The N > 1 host path (band partition + all-gather assembly, frame sharding) on CPU: world_size 2 and 3
over gloo.  The band renderer is the oracle here (the CUDA band render itself is checked against the
whole frame in tests/test_gpu_parity.py::test_bands_reassemble_to_the_whole_frame)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, width, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from swift3drenderer_b200 import multigpu, scene as S
    from oracle import port as oracle_port
    sc = S.shipped_scene(1)
    osc = oracle_port.OracleScene(sc)
    mats = oracle_port.camera_path(S.input_script("flythrough", 600)[:120])
    whole = osc.render(mats[110], width, height)["pixels"]

    def render_band(y0, y1):  # a rank only ever produces its own rows
        return torch.from_numpy(whole[y0:y1].astype(np.int32))

    frame = multigpu.render_banded(render_band, height, rank, world)
    ok = np.array_equal(frame.numpy().view(np.uint32), whole)
    if multigpu.equal_bands(height, world):  # in-place fast path: each rank fills only its own rows
        y0, y1 = multigpu.band_edges(height, world)[rank]
        mine = torch.zeros((height, width), dtype=torch.int32)
        mine[y0:y1] = render_band(y0, y1)
        ok = ok and np.array_equal(multigpu.gather_bands_inplace(mine, rank, world).numpy().view(np.uint32), whole)
    # interleaved tile rows: every rank fills its compacted buffer, one all-gather + de-interleave
    th = 32
    asm = multigpu.InterleavedAssembler(height, width, rank, world, torch.device("cpu"), th)
    from swift3drenderer_b200.renderer import rows_layout
    _, frame_rows, buf_rows = rows_layout(height, world, rank, th)
    asm.mine[torch.from_numpy(buf_rows)] = torch.from_numpy(whole[frame_rows].astype(np.int32))
    ok = ok and np.array_equal(asm.gather().numpy().view(np.uint32), whole)
    shard = list(multigpu.frame_shard(600, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    covered = sorted(sum(gathered, []))
    ok = ok and covered == list(range(600))
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 90), (3, 91)])
def test_banded_assembly_matches_whole_frame(world, height, tmp_path, oracle_port):
    mp.spawn(_worker, args=(world, _free_port(), height, 160, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.ok").read() == "1", f"rank {r}"


def test_band_edges_cover_rows_exactly():
    from swift3drenderer_b200 import multigpu
    for h, n in ((2160, 8), (2160, 7), (91, 3), (5, 5), (1080, 1)):
        e = multigpu.band_edges(h, n)
        assert e[0][0] == 0 and e[-1][1] == h and all(a[1] == b[0] for a, b in zip(e, e[1:]))
        assert all(b > a for a, b in e) and max(b - a for a, b in e) - min(b - a for a, b in e) <= 1
