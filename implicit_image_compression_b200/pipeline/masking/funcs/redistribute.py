"""Layer statistics that steer how regrowth is split between layers (reference:
pipeline/masking/funcs/redistribute.py)."""
import torch


def momentum_redistribution(masking, name, weight, mask):
    """redistribute.py:19-40: mean |Adam momentum| over the active weights."""
    momentum = masking.get_momentum_for_weight(weight)
    return torch.abs(momentum[mask.bool()]).mean().item()


def grad_redistribution(masking, name, weight, mask):
    """redistribute.py:43-63."""
    return torch.abs(weight.grad[mask.bool()]).mean().item()


def nonzero_redistribution(masking, name, weight, mask):
    """redistribute.py:66-87."""
    return (weight != 0.0).sum().item()


registry = {"grad": grad_redistribution, "momentum": momentum_redistribution,
            "nonzero": nonzero_redistribution, "none": nonzero_redistribution}
