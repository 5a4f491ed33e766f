"""CPU, world_size 2, gloo: the host-side sharding / final-gather logic of the multi-GPU path.
The per-track compute is stood in by the CPU oracle (tests may use it); on GPUs it is Tracker.step."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shard = importlib.import_module("parallel-video-object-tracker_b200.shard")


def test_partition_is_exact_and_balanced():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.shard_tracks(n, world, r) for r in range(world)]
            allids = sorted(i for p in parts for i in p)
            assert allids == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            for r, p in enumerate(parts):
                assert len(p) == shard.local_count(n, world, r)
                for slot, i in enumerate(p):
                    assert shard.owner_of(i, world) == (r, slot)
    with pytest.raises(ValueError):
        shard.shard_tracks(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_tracks, n_steps, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from tools import synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O.lib().orc_set_threads(1)

    def make_local(ids):
        st = []
        for i in ids:                                    # stream i has its own seed: independent tracks
            c = synth.make_clip(seed=50 + i, W=160, H=120, tw=16, th=16, n_frames=n_steps + 1, R=20)
            g0 = O.to_gray_f32(c["frames"][0])
            x, y, w, h = c["roi"]
            st.append({"frames": c["frames"], "templ": np.ascontiguousarray(g0[y:y + h, x:x + w]), "xy": [x, y]})
        return st

    def step_fn(st, k):
        out = np.zeros((len(st), 5), np.float64)
        for j, s in enumerate(st):
            rec, _, _ = O.track_step(O.to_gray_f32(s["frames"][k + 1]), s["templ"], s["xy"][0], s["xy"][1], rx=20, ry=20)
            s["xy"] = [rec.x, rec.y]
            out[j] = (rec.x, rec.y, rec.conf, rec.moved, rec.updated)
        return out

    full = shard.run_sharded(n_tracks, n_steps, world, rank, make_local, step_fn, dist)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_matches_single_process():
    import torch.multiprocessing as mp

    n_tracks, n_steps, world = 5, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tracks, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run of the same tracks
    q1 = ctx.Queue()
    p1 = ctx.Process(target=_worker, args=(0, 1, _free_port(), n_tracks, n_steps, q1))
    p1.start()
    _, single = q1.get(timeout=240)
    p1.join(timeout=60)
    assert single.shape == (n_steps, n_tracks, 5)
    assert np.array_equal(res[0], res[1]), "ranks disagree after the gather"
    assert np.array_equal(res[0], single), "sharded run differs from the single-process run"
    assert (single[:, :, 3] == 1).all()                    # every synthetic track is followed
