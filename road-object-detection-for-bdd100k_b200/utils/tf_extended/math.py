"""utils/tf_extended/math.py:25-38."""
import torch

__all__ = ["safe_divide"]


def safe_divide(numerator, denominator, name=None):
    """0 where `denominator` <= 0, else numerator / denominator."""
    return torch.where(denominator > 0, numerator / denominator, torch.zeros_like(numerator))
