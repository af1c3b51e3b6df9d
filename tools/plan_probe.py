"""Experiment: how many unique neighbour rows does a 128-vertex tile gather, for different vertex orders?
(CPU, numpy; uses the oracle lattice -- a tool, not product code.)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lattice as OL
from hplflownet_b200.synthetic import frustum_pair, box_cloud


def coords_from_table(nbr, offs):
    """propagate integer coordinates through the table (BFS per component)."""
    F, H = nbr.shape
    coord = np.zeros((H, 4), np.int64)
    root = np.full(H, -1, np.int64)
    for s in range(H):
        if root[s] >= 0:
            continue
        root[s] = s
        front = np.array([s])
        while front.size:
            nxt = []
            for f in range(F):
                u = nbr[f, front]
                ok = u >= 0
                uu, vv = u[ok], front[ok]
                new = root[uu] < 0
                uu, vv = uu[new], vv[new]
                uu, first = np.unique(uu, return_index=True)
                vv = vv[first]
                root[uu] = s
                coord[uu] = coord[vv] + offs[f]
                nxt.append(uu)
            front = np.concatenate(nxt) if nxt else np.array([], np.int64)
    return coord, root


def morton3(c):
    c = c - c.min(0)
    key = np.zeros(len(c), np.int64)
    for b in range(12):
        for d in range(3):
            key |= ((c[:, d] >> b) & 1) << (3 * b + d)
    return key


def tile_stats(nbr, order, tm=128, taps=None):
    F, H = nbr.shape
    out = []
    for t0 in range(0, H, tm):
        rows = order[t0:t0 + tm]
        sub = nbr[:, rows] if taps is None else nbr[taps][:, rows]
        u = np.unique(sub[sub >= 0])
        out.append(len(u))
    return np.array(out)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "frustum"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    n = 8192
    if which == "frustum":
        pc1, pc2 = frustum_pair(n, 0)
    else:
        pc1 = box_cloud(n, 0); pc2 = box_cloud(n, 1)
    d = OL.generate(pc1, pc2, [[scale, 1, -1, -1]])[0]
    nbr = d["pc1_blur_neighbors"]
    H = nbr.shape[1]
    offs = OL.neighbor_offsets(1)
    print("H", H, "valid neighbour fraction", (nbr >= 0).mean())
    coord, root = coords_from_table(nbr, offs)
    print("components", len(np.unique(root)))
    for name, order in (("api", np.arange(H)),
                        ("morton(c0,c1,c2)", np.argsort(morton3(coord[:, :3]), kind="stable")),
                        ("lex", np.lexsort((coord[:, 2], coord[:, 1], coord[:, 0])))):
        for tm in (128,):
            s = tile_stats(nbr, order, tm)
            print("%-18s tm=%d  unique rows per tile: mean %.0f  p50 %.0f  p90 %.0f  max %d   (refs %d)" %
                  (name, tm, s.mean(), np.median(s), np.percentile(s, 90), s.max(), 15 * tm))


main()
