"""Drop-in for the model builder of ``tone_bias_optuna`` (reference src/tone_bias_optuna.py) -- evaluation only.

``define_isic_model(classes, trial)`` (reference :123-173) builds an ``nn.Sequential`` of a 7x7 block, 1-6 3x3
blocks and 2-5 hidden Linear layers whose widths (16..256) come from an Optuna trial; ``TrialDummy`` (:47-76)
replays a fixed dict of hyperparameters and ``create_best_model`` (:116-120) the shipped "best" trial
(192 / 172 / 22 / 86 conv channels, 227 / 80 / 86 linear units).  Here the same builder returns a Sequential with
the same children and ``state_dict()`` keys (``0.weight``, ``3.weight`` ...) whose ``forward`` runs the sm_100a
kernels: arbitrary widths go through the same conv / linear kernels on channel buffers zero-padded to multiples of
64 (``tone_bias_model.CnnPlan``).  The Optuna search itself (objective / study, :234-343) is training and out of
scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import SiaError
from .tone_bias_model import CnnPlan, _B200Eval

CLASSES = 2

__all__ = ["TrialDummy", "create_best_hyperparameters", "create_best_model", "define_isic_model", "CLASSES"]


class TrialDummy:
    """Dict of hyperparameters behind the Optuna ``trial.suggest_*`` interface (reference :47-76).  The bounds
    check is the reference's: a value BELOW the minimum raises ValueError; the upper bound is not enforced
    (the reference compares ``min_value > max_value`` there, :64 / :71)."""

    def __init__(self, hyperparameters):
        self.hyperparameters = hyperparameters

    def put(self, key, value):
        self.hyperparameters[key] = value

    def get(self, key):
        return self.hyperparameters[key]

    def _checked(self, key, min_value, max_value):
        value = self.get(key)
        if value < min_value or min_value > max_value:
            raise ValueError(f"Expected value between in [{min_value},{max_value}] but got {value}")
        return value

    def suggest_int(self, key, min_value, max_value):
        return int(self._checked(key, min_value, max_value))

    def suggest_float(self, key, min_value, max_value):
        return float(self._checked(key, min_value, max_value))

    def __str__(self):
        return str(self.hyperparameters)


def create_best_hyperparameters():
    """The trial the reference ships as best (reference :96-114; "TRIALS=100, SAMPLESIZE=96*2, EPOCHS=10")."""
    return TrialDummy({
        "n_conv_layers": 3,
        "n_units_l0": 192,
        "n_units_conv_l0": 172,
        "n_units_conv_l1": 22,
        "n_units_conv_l2": 86,
        "n_linear_layers": 3,
        "n_units_linear_l0": 227,
        "dropout_l0": 0.4750108276372097,
        "n_units_linear_l1": 80,
        "dropout_l1": 0.33605861431570366,
        "n_units_linear_l2": 86,
        "dropout_l2": 0.26780264501531464,
        "optimizer": "Adam",
        "lr": 0.03627331743927454,
    })


class _B200Sequential(_B200Eval, nn.Sequential):
    """nn.Sequential container (same children, same state_dict keys) with the B200 evaluation forward."""

    def __init__(self, *layers):
        nn.Sequential.__init__(self, *layers)
        self.class_names = None

    @classmethod
    def from_modules(cls, modules):
        """Same children as an existing ``nn.Sequential`` (e.g. the reference's own, un-pickled) -- parameters shared."""
        return cls(*list(modules))


def define_isic_model(classes, trial):
    """Same layer recipe as the reference (:123-173), same ``trial.suggest_*`` call order and bounds."""
    if classes != 2:
        raise SiaError("the fused tail handles exactly two classes (benign / malignant)")
    n_conv_layers = trial.suggest_int("n_conv_layers", 1, 6)
    layers = []
    image_size = 224
    in_features = 3
    out_features = trial.suggest_int("n_units_l0", 16, 256)
    layers += [nn.Conv2d(in_features, out_features, kernel_size=7, stride=1, padding="same"), nn.ReLU(),
               nn.MaxPool2d(kernel_size=(2, 2))]
    image_size //= 2
    in_features = out_features
    for i in range(n_conv_layers):
        out_features = trial.suggest_int(f"n_units_conv_l{i}", 16, 256)
        layers += [nn.Conv2d(in_features, out_features, kernel_size=3, stride=1, padding="same"), nn.ReLU(),
                   nn.MaxPool2d(kernel_size=(2, 2))]
        image_size //= 2
        in_features = out_features
    n_linear_layers = trial.suggest_int("n_linear_layers", 2, 5)
    layers.append(nn.Flatten())
    in_features = out_features * (image_size * image_size)
    print(f"DEBUGGING {in_features} = {out_features} * ({image_size}*{image_size})")     # the reference prints this
    for i in range(n_linear_layers):
        out_features = trial.suggest_int(f"n_units_linear_l{i}", 16, 256)
        layers += [nn.Linear(in_features, out_features), nn.ReLU()]
        p = trial.suggest_float(f"dropout_l{i}", 0.2, 0.5)
        layers.append(nn.Dropout(p))
        in_features = out_features
    layers += [nn.Linear(in_features, classes), nn.LogSoftmax(dim=1)]
    return _B200Sequential(*layers)


def create_best_model():
    return define_isic_model(CLASSES, create_best_hyperparameters())
