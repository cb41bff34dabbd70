"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): the multi-GPU cell shard of the
PatchMatch phase over CUDA IPC peer memory, checked on rank 0 against the single-GPU sweep of the same problem -- every rank's copy of
currentCost_ / currentLabeling_ must be bit-identical to it (same cells, same random streams; the shard only changes who computes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import localexpstereo_b200 as L
    from localexpstereo_b200 import synth
    from localexpstereo_b200.sweep import PMSweep
    from lexp_testlib import make_scene
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    H, W, D, windR = 300, 420, 24, 20
    imL, _, volL, _ = make_scene(H, W, D)
    prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
    props = [[(L.PROP_EXPANSION, 1), (L.PROP_RANDOM, 4)], [(L.PROP_EXPANSION, 2), (L.PROP_RANDOM, 1)], [(L.PROP_EXPANSION, 2), (L.PROP_RANDOM, 1)]]
    units = [10, 31, 93]
    E = L.CostVolumeEnergy(imL, None, volL, None, prm, D - 1, device=local)
    S = PMSweep(E, unit_sizes=units, proposers=props, rank=rank, world=world)
    labels = synth.synthetic_planes(S.lm.layers[0].unitRegions, 1, D, 5)[0]
    S.begin()
    handles = [None] * world
    dist.all_gather_object(handles, E.pm_ipc_export(0))
    S.connect(handles)
    dist.barrier()
    S.init(labels[S.init_index])
    n_it = 3
    for it in range(n_it):
        S.iteration(it, 31)
    E.sync()
    dist.barrier()          # every rank's kernels (and their stores into this rank's copy) are done
    cost, lab = S.get()
    ok = True
    ref = None
    if rank == 0:
        E1 = L.CostVolumeEnergy(imL, None, volL, None, prm, D - 1, device=local)
        S1 = PMSweep(E1, unit_sizes=units, proposers=props)
        S1.begin(); S1.init(labels)
        for it in range(n_it):
            S1.iteration(it, 31)
        ref = S1.get()
        S1.close(); E1.close()
    box = [ref]
    dist.broadcast_object_list(box, 0)
    ref = box[0]
    ok = bool(np.array_equal(cost, ref[0]) and np.array_equal(lab, ref[1]) and np.isfinite(cost).all())
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if rank == 0:
        idle = sum(1 for (_, _, g, owners) in S.schedule if len(owners) < world)
        print(f"MULTI_GPU_CHECK world={world} ok={all(oks)} per_rank={oks} groups_with_idle_ranks={idle}", flush=True)
    dist.barrier()
    os._exit(0 if all(oks) else 1)


if __name__ == "__main__":
    main()
