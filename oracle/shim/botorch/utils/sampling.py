"""Stand-in for botorch.utils.sampling: scrambled-Sobol draws generated on CPU."""
import torch
from torch.quasirandom import SobolEngine


def draw_sobol_samples(bounds, n, q, batch_shape=None, seed=None):
    batch_shape = batch_shape or torch.Size()
    batch_size = int(torch.prod(torch.tensor(batch_shape)))
    d = bounds.shape[-1]
    lower = bounds[0]
    rng = bounds[1] - bounds[0]
    sobol_engine = SobolEngine(q * d, scramble=True, seed=seed)
    samples_raw = sobol_engine.draw(batch_size * n, dtype=lower.dtype)
    samples_raw = samples_raw.view(*batch_shape, n, q, d).to(device=lower.device)
    if batch_shape != torch.Size():
        samples_raw = samples_raw.permute(-3, *range(len(batch_shape)), -2, -1)
    return lower + rng * samples_raw


def draw_sobol_normal_samples(d, n, device=None, dtype=None, seed=None):
    """Quasi-random standard-normal rows (input.py:873 uses them as the INITIAL latent embedding tables; the golden cases
    overwrite the tables with seeded values, so only shape / dtype matter here): scrambled Sobol through the inverse
    normal CDF."""
    u = SobolEngine(d, scramble=True, seed=seed).draw(n, dtype=torch.float64)
    v = 0.5 + (1.0 - 1e-10) * (u - 0.5)
    out = (2.0 ** 0.5) * torch.erfinv(2.0 * v - 1.0)
    return out.to(device=device, dtype=dtype or torch.get_default_dtype())


def sample_simplex(d, n=1, qmc=False, seed=None, device=None, dtype=None):  # imported by input.py:16, unused on the PR path
    raise NotImplementedError
