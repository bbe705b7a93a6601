"""Stand-in for botorch.models.transforms.input: InputTransform / Normalize / ChainedInputTransform.

Semantics restated from the BoTorch API the reference relies on:
* InputTransform.forward dispatches to .transform in train/eval mode according to transform_on_* flags.
* Normalize(d, bounds, indices, reverse): _transform = (x - min) / range, _untransform = min + x * range,
  applied to `indices` columns only; reverse swaps the two.
* ChainedInputTransform: ordered ModuleDict applying members in insertion order.
"""
from collections import OrderedDict

import torch
from torch.nn import Module, ModuleDict


class InputTransform:
    transform_on_train: bool = True
    transform_on_eval: bool = True
    transform_on_fantasize: bool = True

    def forward(self, X):
        if self.training:
            if self.transform_on_train:
                return self.transform(X)
        elif self.transform_on_eval:
            return self.transform(X)
        return X

    def transform(self, X):  # pragma: no cover - abstract
        raise NotImplementedError

    def untransform(self, X):  # pragma: no cover - abstract
        raise NotImplementedError

    def equals(self, other):
        return type(self) is type(other)


class Normalize(InputTransform, Module):
    def __init__(self, d, indices=None, bounds=None, batch_shape=torch.Size(), transform_on_train=True,
                 transform_on_eval=True, transform_on_fantasize=True, reverse=False, min_range=1e-8,
                 learn_bounds=None):
        super().__init__()
        if indices is not None and len(indices) > 0:
            self.register_buffer("indices", torch.as_tensor(indices, dtype=torch.long))
        mins = bounds[..., 0:1, :]
        ranges = bounds[..., 1:2, :] - mins
        self.register_buffer("mins", mins.clone())
        self.register_buffer("ranges", ranges.clone())
        self.transform_on_train = transform_on_train
        self.transform_on_eval = transform_on_eval
        self.transform_on_fantasize = transform_on_fantasize
        self.reverse = reverse

    def _apply_cols(self, X, fn):
        if hasattr(self, "indices"):
            X_new = X.clone()
            X_new[..., self.indices] = fn(X[..., self.indices], self.mins[..., self.indices], self.ranges[..., self.indices])
            return X_new
        return fn(X, self.mins, self.ranges)

    def _norm(self, X):
        return self._apply_cols(X, lambda x, m, r: (x - m) / r)

    def _unnorm(self, X):
        return self._apply_cols(X, lambda x, m, r: m + x * r)

    def transform(self, X):
        return self._unnorm(X) if self.reverse else self._norm(X)

    def untransform(self, X):
        return self._norm(X) if self.reverse else self._unnorm(X)


class ChainedInputTransform(InputTransform, ModuleDict):
    def __init__(self, **transforms):
        super().__init__(OrderedDict(transforms))
        self.transform_on_train = False
        self.transform_on_eval = False
        self.transform_on_fantasize = False
        for tf in transforms.values():
            self.transform_on_train |= tf.transform_on_train
            self.transform_on_eval |= tf.transform_on_eval
            self.transform_on_fantasize |= tf.transform_on_fantasize

    def transform(self, X):
        for tf in self.values():
            X = tf.forward(X)
        return X

    def untransform(self, X):
        for tf in reversed(self.values()):
            X = tf.untransform(X)
        return X
