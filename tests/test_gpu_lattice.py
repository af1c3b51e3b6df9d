"""GPU lattice builder vs the reference: BIT-EXACT on every entry of generated_data
(fixtures dumped from the unmodified reference; the oracle at BASELINE size)."""
import numpy as np
import pytest
import torch

from hplflownet_b200.synthetic import box_cloud, frustum_pair
from hplflownet_b200.transforms import GenerateDataUnsymmetric, collate_batch1
from oracle import lattice as OL
from tests._util import bits_equal, golden, golden_files

pytestmark = pytest.mark.gpu


class _Args:
    dim = 3

    def __init__(self, sfm):
        self.scales_filter_map = sfm


FULL = [[3., 1, -1, -1], [2., 1, -1, -1], [1., 1, 1, 1], [.5, 1, 1, 1], [.25, 1, 1, 1], [.125, 1, 1, 1],
        [.0625, 1, 1, 1]]


def _compare(got, want_fn, n_scales):
    for k in range(n_scales):
        d = got[k]
        for key, v in d.items():
            ref = want_fn(k, key)
            if isinstance(v, int):
                assert v == int(ref), (k, key, v, int(ref))
            else:
                assert bits_equal(v.cpu().numpy(), np.asarray(ref)), "scale %d %s differs" % (k, key)


@pytest.mark.parametrize("name", golden_files("lattice_"))
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_matches_reference_fixture(name, idx_dtype):
    g = golden(name)
    sfm = [[float(r[0]), int(r[1]), int(r[2]), int(r[3])] for r in g["scales_filter_map"]]
    gen = GenerateDataUnsymmetric(_Args(sfm), index_dtype=idx_dtype)
    pc1, pc2, sf, got = gen([g["pc1"], g["pc2"], np.zeros_like(g["pc1"])])
    assert pc1.shape == (3, g["pc1"].shape[0]) and pc1.is_cuda
    for d in got:
        assert d["pc1_lattice_offset"].dtype == idx_dtype
    _compare(got, lambda k, key: g["s%d_%s" % (k, key)], len(sfm))


@pytest.mark.parametrize("n,seed", [(8192, 0), (8192, 7), (2048, 1)])
def test_full_hierarchy_matches_oracle_at_baseline_size(n, seed):
    pc1, pc2 = frustum_pair(n, seed)
    want = OL.generate(pc1, pc2, FULL)
    gen = GenerateDataUnsymmetric(_Args(FULL))
    got = gen([pc1, pc2, pc2 - pc1])[3]
    _compare(got, lambda k, key: want[k][key], len(FULL))


def test_ragged_box_and_repeat_determinism():
    a, b = box_cloud(5000, 1, 12.0), box_cloud(3111, 2, 12.0)
    sfm = [[2.0, 1, 1, 1], [1.0, 1, 1, 1]]
    want = OL.generate(a, b, sfm)
    gen = GenerateDataUnsymmetric(_Args(sfm))
    got1 = gen([a, b, np.zeros_like(a)])[3]
    got2 = gen([a, b, np.zeros_like(a)])[3]
    _compare(got1, lambda k, key: want[k][key], 2)
    for d1, d2 in zip(got1, got2):          # parallel insert, deterministic result
        for key in d1:
            if not isinstance(d1[key], int):
                assert torch.equal(d1[key], d2[key])


def test_duplicate_points_collide_in_the_hash():
    # every point repeated 8x: massive key collisions, ids must still be first-occurrence order
    base = frustum_pair(600, 4)[0]
    pc = np.repeat(base, 8, axis=0)
    sfm = [[1.0, 1, 1, 1]]
    want = OL.generate(pc, pc[::-1].copy(), sfm)
    got = GenerateDataUnsymmetric(_Args(sfm))([pc, pc[::-1].copy(), np.zeros_like(pc)])[3]
    _compare(got, lambda k, key: want[k][key], 1)


def test_collate_adds_batch_axis():
    pc1, pc2 = frustum_pair(256, 2)
    got = GenerateDataUnsymmetric(_Args([[1.0, 1, -1, -1]]))([pc1, pc2, pc1])[3]
    col = collate_batch1(got)
    assert col[0]["pc1_barycentric"].shape == (1, 4, 256)
    assert col[0]["pc1_hash_cnt"].item() == got[0]["pc1_hash_cnt"]
    assert col[0]["pc1_corr_indices"].shape == (1, 1)
