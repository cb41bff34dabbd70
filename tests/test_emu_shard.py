"""The N > 1 cell-shard data path on the CPU: two gloo ranks, each with its own (emulated) engine holding all inputs,
evaluate their share of every disjoint group (sweep.shard_cells) with lexp_plan_eval_device_tiles, exchange the per-cell
unary tiles with one all-gather per group -- the collective bench.py issues over NCCL -- and reassemble the cost image,
which must equal the unsharded evaluation bit for bit."""
import os
import socket
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    from emu import emu_lib
    import localexpstereo_b200 as L
    from localexpstereo_b200.sweep import shard_cells, tile_offsets
    from oracle import lexp_oracle as O
    from lexp_testlib import make_scene
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    H, W, D, windR = 120, 168, 16, 20
    imL, imR, volL, volR = make_scene(H, W, D, seed=5)
    ok = True
    with emu_lib.emulated():
        prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
        E = L.CostVolumeEnergy(imL, imR, volL, volR, prm, D - 1)
        lay = L.LayerManager(W, H, windR).addLayer(12)
        rng = O.CvRNG(123)  # same seed on every rank: identical planes, as after a broadcast
        for g in lay.disjointRegionSets[:6]:
            planes = np.stack([O.create_random_label(rng, *lay.unitRegions[r][:2], 0.0, D - 1.0) for r in g])
            # unsharded truth (every rank can compute it: all inputs are replicated)
            full = np.full((H, W), -7.0, np.float32)
            E.ComputeUnaryPotentialBatch([lay.filterRegions[r] for r in g], [lay.sharedRegions[r] for r in g], full, planes)
            # this rank's share -> contiguous tiles
            mine = shard_cells(np.arange(len(g)), rank, world)
            rects = [lay.sharedRegions[g[i]] for i in mine]
            offs, total = tile_offsets(rects)
            sizes = [None] * world
            dist.all_gather_object(sizes, total)
            mx = max(sizes)
            tiles = np.full(mx, -1.0, np.float32)
            if len(mine):
                plan = E.make_plan([lay.filterRegions[g[i]] for i in mine], rects)
                plan.eval_device_tiles(planes[mine], tiles.ctypes.data, True, 0)
                E.sync()
                plan.close()
            gathered = [torch.empty(mx) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(tiles))
            img = np.full((H, W), -7.0, np.float32)
            for rk in range(world):
                idx = shard_cells(np.arange(len(g)), rk, world)
                rects_rk = [lay.sharedRegions[g[i]] for i in idx]
                offs_rk, _ = tile_offsets(rects_rk)
                buf = gathered[rk].numpy()
                for (x, y, w, h), o in zip(rects_rk, offs_rk):
                    img[y:y + h, x:x + w] = buf[o:o + w * h].reshape(h, w)
            ok &= bool(np.array_equal(img, full))
        E.close()
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_cell_shard_with_the_emulated_engine_world2_gloo():
    import torch.multiprocessing as mp
    from emu import emu_lib
    emu_lib.load()  # build the emulator library once, before the workers race for it
    os.environ["LEXP_EMU_NO_REBUILD"] = "1"  # inherited by the spawned workers: they load what the parent built
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=400) for _ in procs]
    finally:
        os.environ.pop("LEXP_EMU_NO_REBUILD", None)
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
