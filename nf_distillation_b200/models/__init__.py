"""Same exports as the reference's models/__init__.py (VGGPerceptualLoss is out of scope: SURVEY.md §2 #9)."""
from .kd_flows import create_glow_model, inherit_permutation_matrix
from .flows import FlowStep
from .layers import gaussian_sample, SqueezeLayer

__all__ = ["create_glow_model", "inherit_permutation_matrix", "FlowStep", "gaussian_sample", "SqueezeLayer"]
