"""The call schedule of one local-expansion sweep over one view, as the reference issues it
(FastGCStereo::localExpansionMovesForLayer_CPU, FastGCStereo.h:22-72): for each layer, for each
of its <= 16 disjoint groups (sequential), for each proposal step k (sequential), ONE batched
evaluation of all cells of the group (the `omp parallel for` axis, :30-31).

Multi-GPU: the cells of every group are dealt round-robin to ranks (cell-shard path,
SURVEY.md section 8e); each rank owns a plan per (layer, group) for its own cells only."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .energy import CostVolumeEnergy, LayerManager, Plan


def v3_layer_units(width):  # main.cpp:395-397
    return [int(width * 0.01), int(width * 0.03), int(width * 0.09)]


V3_STEPS = [9, 3, 3]  # Exp(1)+Ransac(1)+Random(7) | Exp(2)+Ransac(1) | same (main.cpp:391-397)


def shard_cells(cells, rank, world):
    """Cell-shard rule: the cells of one disjoint group are dealt round-robin to ranks."""
    return np.asarray(cells, dtype=np.int64)[rank::world]


def tile_offsets(target_rects):
    """Float offsets of the per-cell unary tiles inside a rank's contiguous tile buffer (the all-gather payload,
    same rule as lexp_plan_eval_device_tiles): tile i starts at the sum of the areas of tiles 0..i-1."""
    areas = np.array([r[2] * r[3] for r in target_rects], dtype=np.int64)
    return np.concatenate([[0], np.cumsum(areas)[:-1]]) if len(areas) else np.zeros(0, np.int64), int(areas.sum())


@dataclass
class GroupPlan:
    layer: int
    group: int
    cells: np.ndarray        # indices (into the layer's cell list) evaluated by THIS rank
    plan: Plan
    n_steps: int


class UnarySweep:
    def __init__(self, energy: CostVolumeEnergy, unit_sizes=None, steps=None, rank=0, world=1):
        self.energy = energy
        self.unit_sizes = unit_sizes or v3_layer_units(energy.width)
        self.steps = steps or V3_STEPS
        self.lm = LayerManager(energy.width, energy.height, energy.params.windR)
        self.groups: List[GroupPlan] = []
        self.total_filter_px = 0     # sum over the WHOLE sweep (all ranks): evals
        self.total_target_px = 0
        self.local_filter_px = 0
        self.local_target_px = 0
        self.local_alg_bytes = 0
        for li, u in enumerate(self.unit_sizes):
            lay = self.lm.addLayer(u)
            K = self.steps[li]
            for gi, cells in enumerate(lay.disjointRegionSets):
                cells = np.asarray(cells, dtype=np.int64)
                self.total_filter_px += K * sum(lay.filterRegions[r][2] * lay.filterRegions[r][3] for r in cells)
                self.total_target_px += K * sum(lay.sharedRegions[r][2] * lay.sharedRegions[r][3] for r in cells)
                mine = shard_cells(cells, rank, world)
                if len(mine) == 0:
                    continue
                plan = energy.make_plan([lay.filterRegions[r] for r in mine], [lay.sharedRegions[r] for r in mine])
                self.groups.append(GroupPlan(li, gi, mine, plan, K))
                self.local_filter_px += K * plan.filter_px
                self.local_target_px += K * plan.target_px
                self.local_alg_bytes += K * plan.algorithmic_bytes
        self.launches_per_sweep = sum(g.n_steps for g in self.groups)

    def layer(self, li):
        return self.lm.layers[li]

    def close(self):
        for g in self.groups:
            g.plan.close()
        self.groups = []
