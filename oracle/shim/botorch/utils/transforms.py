"""Stand-in for botorch.utils.transforms.normalize_indices (imported by input.py:18)."""


def normalize_indices(indices, d):
    if indices is None:
        return None
    out = []
    for i in indices:
        if i < 0:
            i = i + d
        if i < 0 or i > d - 1:
            raise ValueError(f"Index {i} out of bounds for tensor or length {d}.")
        out.append(i)
    return out
