"""Dynamic models."""

from .lipm_model import LIPMModel, discrete_model, zmp_preview_matrices, velocity_preview_matrices

__all__ = ["LIPMModel", "discrete_model", "zmp_preview_matrices", "velocity_preview_matrices"]
