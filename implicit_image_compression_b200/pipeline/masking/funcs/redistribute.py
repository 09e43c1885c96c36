"""Layer statistics that steer how regrowth is split between layers (reference:
pipeline/masking/funcs/redistribute.py).  Each rule returns a 0-d tensor on the weight's device; the caller
(`Masking.gather_statistics`) reads all layers' values back in ONE transfer.  The reference gathers the active
elements with a boolean index (a device-to-host round trip per layer for the element count) and calls `.item()`;
here the mean over the active elements is a masked sum divided by the active count — the same value up to fp32
summation order."""
import torch


def _masked_mean_abs(values, mask):
    active = mask != 0
    return (torch.abs(values) * active).sum() / active.sum()  # 0 / 0 = nan for a fully pruned layer, as mean([])


def momentum_redistribution(masking, name, weight, mask):
    """redistribute.py:19-40: mean |Adam momentum| over the active weights."""
    return _masked_mean_abs(masking.get_momentum_for_weight(weight), mask)


def grad_redistribution(masking, name, weight, mask):
    """redistribute.py:43-63."""
    return _masked_mean_abs(weight.grad, mask)


def nonzero_redistribution(masking, name, weight, mask):
    """redistribute.py:66-87."""
    return (weight != 0.0).sum()


registry = {"grad": grad_redistribution, "momentum": momentum_redistribution,
            "nonzero": nonzero_redistribution, "none": nonzero_redistribution}
