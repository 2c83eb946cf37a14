"""mla_b200 — B200-native drop-in for MLA's alternating unimodal training step and test-time
fusion (reference: Cecile-hi/Multimodal-Learning-with-Alternating-Unimodal-Adaptation).

The directory name contains hyphens, so import it through the `mla_b200` alias package at the
repo root (`import mla_b200`). Names mirror the reference's modules:
    GSPlugin, setup_seed, weight_init            (utils/utils.py)
    calculate_entropy, calculate_gating_weights[3], train_epoch, valid, get_arguments (main.py)
    AVClassifier, M3AEClassifier, Modal3Classifier (models/basic_model.py, models/m3ae.py, models/cav_mae.py)
    ConcatFusion, ConcatFusion3                  (models/fusion_modules.py)
    resnet18                                     (models/backbone.py)
    FrameBatchProducer                           (dataset/dataset.py: the visual transform pipeline, per batch on the GPU)
"""
from . import _lib, ops  # noqa: F401
from .gs_plugin import GSPlugin  # noqa: F401
from .fusion import calculate_entropy, calculate_gating_weights, calculate_gating_weights3  # noqa: F401
from .fusion_modules import ConcatFusion, ConcatFusion3, head_turn  # noqa: F401
from .backbone import resnet18  # noqa: F401
from .basic_model import AVClassifier  # noqa: F401
from .m3ae import M3AEClassifier  # noqa: F401
from .cav_mae import Modal3Classifier  # noqa: F401
from .utils import setup_seed, weight_init  # noqa: F401
from .engine import ModuleHolder, train_epoch, valid  # noqa: F401
from .main import get_arguments  # noqa: F401
from .dataset import FrameBatchProducer, SpecBatchProducer  # noqa: F401
