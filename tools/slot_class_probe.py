"""GPU experiment (plan statistics only): how many shared-memory wavefronts would a copy thread's row reads cost if staged rows
were rotated by their SLOT (slot % 8 = bank-group class) instead of by the reader's lane -- for the plan's current slot
numbering and for a greedy per-tile colouring of the distinct rows into 8 classes (<= 58 rows each)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hplflownet_b200 import plans
from hplflownet_b200.batching import concat_lattices

dev = torch.device("cuda")
nbr = concat_lattices([bench.cloud_tables(s) for s in range(8)])["blur_neighbors"][0].to(dev)
plan = plans.build(nbr)
local = (plan.view(3).long() & 0xffff).cpu().numpy()[:, :15]          # (tiles, taps, 128)
nu = plan.view(1).cpu().numpy()
umax = 464
rng = np.random.default_rng(0)
tiles = rng.choice(plan.n_tiles, size=24, replace=False)


def degree(cls_of_slot, lt):
    """mean over (tap, quarter-warp group) of the max multiplicity of a class among the 8 rows (zero-row reads excluded)."""
    tot, cnt = 0, 0
    for f in range(lt.shape[0]):
        for g in range(16):
            s = lt[f, 8 * g:8 * g + 8]
            s = s[s < umax]
            if len(s) == 0:
                continue
            tot += np.bincount(cls_of_slot[s], minlength=8).max()
            cnt += 1
    return tot / max(cnt, 1)


cur, greedy, worst_class = [], [], []
for t in tiles:
    lt = local[t]
    n = int(nu[t])
    cur.append(degree(np.arange(umax) % 8, lt))
    # greedy: visit groups, give each still uncoloured row the least used class of its group (ties: globally least loaded)
    cls = -np.ones(umax, dtype=np.int64)
    load = np.zeros(8, dtype=np.int64)
    groups = [lt[f, 8 * g:8 * g + 8] for f in range(15) for g in range(16)]
    for s in groups:
        s = np.unique(s[s < umax])
        used = np.bincount(cls[s][cls[s] >= 0], minlength=8)
        for r in s:
            if cls[r] >= 0:
                continue
            order = np.lexsort((load, used))                       # least used in the group, then least loaded overall
            c = next(c for c in order if load[c] < 58)
            cls[r] = c
            used[c] += 1
            load[c] += 1
    cls[cls < 0] = 0
    greedy.append(degree(cls, lt))
    worst_class.append(load.max())
print("tiles probed: %d   distinct rows per tile: mean %.0f max %d" % (len(tiles), nu[tiles].mean(), nu[tiles].max()))
print("wavefront multiplier of the row reads, slot-rotated staging: current slot numbering %.2f   greedy classes %.2f   (lane rotation: 1.00)"
      % (np.mean(cur), np.mean(greedy)))
print("largest class after greedy colouring: %d rows (cap 58)" % max(worst_class))
