"""Several clouds in one launch: concatenate per-cloud lattice tables into one "super cloud".

The reference supports batch_size = 1 only (README.md:57; models/bilateralNN.py:137-140: splat
indices carry no batch offset).  Every lattice here is independent, so a batch is simply the
concatenation of the clouds' points and vertices with vertex ids shifted by the running vertex
count (-1 stays -1).  All kernels are oblivious to cloud boundaries.
"""
import torch


def concat_lattices(items, prefix="pc1"):
    """items: list of per-cloud dicts with ``{prefix}_barycentric`` (4, N), ``_lattice_offset`` (4, N),
    ``_blur_neighbors`` (F, H), ``_hash_cnt``.  Returns a dict of batched (1, ., sum) tensors plus
    ``point_counts`` / ``vertex_counts`` lists."""
    bary, off, nbr, npts, nvert = [], [], [], [], []
    base = 0
    for d in items:
        b, o, nb = d[prefix + "_barycentric"], d[prefix + "_lattice_offset"], d[prefix + "_blur_neighbors"]
        h = int(d[prefix + "_hash_cnt"])
        bary.append(b)
        off.append(o + base)
        nbr.append(torch.where(nb >= 0, nb + base, nb))
        npts.append(b.size(-1))
        nvert.append(h)
        base += h
    return {
        "barycentric": torch.cat(bary, -1)[None].contiguous(),
        "lattice_offset": torch.cat(off, -1)[None].contiguous(),
        "blur_neighbors": torch.cat(nbr, -1)[None].contiguous(),
        "point_counts": npts, "vertex_counts": nvert,
    }
