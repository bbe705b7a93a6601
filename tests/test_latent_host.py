"""CPU: the LatentCategoricalEmbedding mirror (input.py:777-952 of the reference) and get_kernel("mixed_latent")
(kernels.py:73-84) -- host logic only, no device calls."""
import torch

import bo_pr_b200 as B
from bo_pr_b200.kernels import kernel_terms
from oracle import gp_oracle as G


def _emb():
    specs = [B.LatentCategoricalSpec(idx=2, num_categories=4, latent_dim=2),
             B.LatentCategoricalSpec(idx=3, num_categories=12, latent_dim=2),
             B.LatentCategoricalSpec(idx=-1, num_categories=3, latent_dim=1)]
    return B.LatentCategoricalEmbedding(specs, dim=5)


def test_transform_layout_pinning_and_round_trip():
    emb = _emb()
    assert emb.categ_idcs.tolist() == [2, 3, 4] and emb._emb_dim == 5
    g = torch.Generator().manual_seed(0)
    X = torch.rand(9, 5, generator=g, dtype=torch.float64)
    X[:, 2] = torch.randint(0, 4, (9,), generator=g)
    X[:, 3] = torch.randint(0, 12, (9,), generator=g)
    X[:, 4] = torch.randint(0, 3, (9,), generator=g)
    Z = emb.transform(X)
    assert Z.shape == (9, 7)
    assert torch.equal(Z[:, :2], X[:, :2])
    tables = [emb.get_emb_table(i).detach() for i in (2, 3, 4)]
    for t in tables:  # row 0 is the origin; a 2-d table has its second row on the second axis
        assert bool((t[0] == 0).all())
        if t.shape[1] > 1:
            assert float(t[1, 0]) == 0.0
    # the oracle's restatement agrees, and nearest-embedding decoding inverts the transform
    assert torch.equal(Z.detach(), G.latent_embedding_transform(tables, [2, 3, 4], 5)(X))
    assert torch.equal(emb.untransform(Z).detach(), X)
    b = emb.transform_bounds(torch.stack([torch.zeros(5), torch.tensor([1.0, 1.0, 3.0, 11.0, 2.0])]).double())
    assert b.shape == (2, 7) and b[1].tolist() == [1.0] * 7 and b[0].tolist() == [0.0] * 7
    # batched inputs keep their leading shape
    assert emb.transform(X.reshape(3, 3, 1, 5)).shape == (3, 3, 1, 7)


def test_get_kernel_mixed_latent_structure():
    ctf = {2: 2, 4: 2, 6: 1}  # first latent column -> latent dim, as experiment_utils.py:291-305 builds it
    k = B.get_kernel("mixed_latent", 7, [], ctf, None, None)
    terms = kernel_terms(k, 7)
    assert len(terms) == 1 and not terms[0][1]
    factors = terms[0][2]
    assert [f[1] for f in factors] == [[0, 1], [2, 3], [4, 5], [6]]
    assert all(f[0] == "matern52" and f[3] == 1 for f in factors)
    spec = G.reference_kernel_spec("mixed_latent", 7, [], ctf, cont_lengthscale=[0.5, 0.7],
                                   latent_lengthscale=[1.0, 2.0, 3.0])
    assert [f.dims for f in spec[0].factors] == [f[1] for f in factors]
    assert spec[0].factors[2].lengthscale == [2.0, 2.0]
