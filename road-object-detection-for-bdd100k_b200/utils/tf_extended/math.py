"""utils/tf_extended/math.py:25-67."""
import torch

__all__ = ["safe_divide", "cummax"]


def safe_divide(numerator, denominator, name=None):
    """0 where `denominator` <= 0, else numerator / denominator."""
    return torch.where(denominator > 0, numerator / denominator, torch.zeros_like(numerator))


def cummax(x, reverse=False, name=None):
    """Cumulative maximum of a 1-D tensor (utils/tf_extended/math.py:41-67)."""
    if x.numel() == 0:
        return x.clone()
    if reverse:
        return torch.flip(torch.cummax(torch.flip(x, [0]), 0).values, [0])
    return torch.cummax(x, 0).values
