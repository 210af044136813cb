"""GPU: the support-graph builder (csrc/radius.cu) against a brute-force distance matrix — the reference's
transforms/support_graph.py:56-59 radius query (self loops included, rows (j, i) grouped by j, at most 512 per j)."""
import types

import pytest
import torch

import fieldconv_b200 as fcb

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _brute(pos, r):
    d = torch.cdist(pos.double(), pos.double())
    j, i = torch.nonzero(d <= r, as_tuple=True)
    return j, i, d


def _as_set(e, n):
    return set((e[:, 0] * n + e[:, 1]).tolist())


@pytest.mark.parametrize("n,r,shape", [(3000, 0.12, "sphere"), (2000, 0.3, "box"), (500, 5.0, "box"), (1, 0.1, "box")])
def test_radius_graph_matches_brute_force(n, r, shape):
    g = torch.Generator().manual_seed(n)
    p = torch.randn(n, 3, generator=g)
    if shape == "sphere":
        p = p / p.norm(dim=1, keepdim=True)
    pos = p.to(DEV)
    e = fcb.radius_graph(pos, r, max_num_neighbors=100000)
    j, i, d = _brute(pos, r)
    # pairs at distance r +- rounding may fall either side in float32: compare outside a thin shell
    shell = (d[j, i] - r).abs() < 1e-5 * max(r, 1.0)
    want = set((j * n + i)[~shell].tolist())
    maybe = set((j * n + i)[shell].tolist())
    got = _as_set(e, n)
    assert want <= got and got <= (want | maybe)
    assert torch.equal(e[:, 0], e[:, 0].sort().values)                 # grouped by the query point (column 0)
    assert _as_set(torch.stack((torch.arange(n), torch.arange(n)), 1).to(DEV), n) <= got      # self loops
    assert len(got) == e.shape[0]                                       # no duplicates


def test_radius_graph_caps_the_neighbour_count_and_support_graph_dropin():
    g = torch.Generator().manual_seed(5)
    pos = torch.rand(4000, 3, generator=g).to(DEV)
    e = fcb.radius_graph(pos, 0.5, max_num_neighbors=16)
    counts = torch.bincount(e[:, 0], minlength=4000)
    assert int(counts.max()) == 16 and int(counts.min()) >= 1
    d = (pos[e[:, 0]] - pos[e[:, 1]]).norm(dim=1)
    assert float(d.max()) <= 0.5 + 1e-6
    # drop-in transform: same attribute names as the reference (data.sample_idx, data.supp_edges)
    data = types.SimpleNamespace(pos=pos)
    out = fcb.SupportGraph(0.1, sample_n=500)(data)
    assert out.sample_idx.shape == (500,) and torch.equal(out.sample_idx, out.sample_idx.sort().values)
    assert out.supp_edges.dtype == torch.int64 and int(out.supp_edges.max()) < 500
    sub = pos[out.sample_idx]
    j, i, _ = _brute(sub, 0.1)
    assert abs(out.supp_edges.shape[0] - j.numel()) <= 2
    data2 = types.SimpleNamespace(pos=pos)
    out2 = fcb.SupportGraph(0.05)(data2)
    assert torch.equal(out2.sample_idx, torch.arange(4000, device=DEV))
