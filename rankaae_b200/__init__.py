"""rankaae_b200 — the RankAAE adversarial-autoencoder train step as fused sm_100a CUDA kernels behind the
reference's Python surface (Trainer / AE_CLS_DICT / Parameters / fix_config.yaml)."""
from .parameter import AE_CLS_DICT, OPTIM_DICT, Parameters  # noqa: F401

__version__ = "0.1.0"
