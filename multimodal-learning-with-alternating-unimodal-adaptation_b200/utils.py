"""Seeding and initialisation recipe (reference utils/utils.py:98-114)."""
import random

import numpy as np
import torch
import torch.nn as nn

from .gs_plugin import GSPlugin  # noqa: F401  (the reference exports GSPlugin from utils.utils)


def setup_seed(seed):
    """utils.py:98-103."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True


def weight_init(m):
    """utils.py:106-114 — applied with model.apply() to AVClassifier only (main.py:719)."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_normal_(m.weight)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
