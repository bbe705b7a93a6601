"""Minimal stand-in for the GPyTorch names reference input.py:19-21 imports (LatentCategoricalEmbedding only; off the PR path)."""
