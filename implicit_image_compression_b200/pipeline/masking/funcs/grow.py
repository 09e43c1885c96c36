"""Growth rules (reference: pipeline/masking/funcs/grow.py): absolute-gradient (RigL), momentum (SNFS),
random, none.  Each returns the new boolean mask of one layer."""
import torch


def _inactive(mask_bool, like):
    return (mask_bool == 0).to(like.dtype)


def momentum_growth(masking, name, total_regrowth, weight):
    """grow.py:25-55: enable the masked-out positions with the largest |Adam momentum|."""
    new_mask = masking.mask_dict[name].data.bool()
    momentum = masking.get_momentum_for_weight(weight)
    momentum = momentum * _inactive(new_mask, momentum)
    _, order = torch.sort(torch.abs(momentum).flatten(), descending=True)
    new_mask.data.view(-1)[order[: int(total_regrowth)]] = 1.0
    return new_mask


def abs_grad_growth(masking, name, total_regrowth, weight):
    """grow.py:58-97: enable the masked-out positions with the largest |grad|; grown weights start at 0."""
    new_mask = masking.mask_dict[name].data.bool()
    # (the reference returns early when the layer has no inactive position; then nothing was pruned either, so
    # total_regrowth is 0 and the selection below is empty - no need to synchronise for the count)
    grad = weight.grad * _inactive(new_mask, weight.grad)
    _, order = torch.sort(torch.abs(grad).flatten(), descending=True)
    pick = order[: int(total_regrowth)]
    new_mask.data.view(-1)[pick] = 1.0
    weight.data.view(-1)[pick] = 0.0
    return new_mask


def random_growth(masking, name, total_regrowth, weight):
    """grow.py:100-136."""
    new_mask = masking.mask_dict[name].data.bool()
    n = (new_mask == 0).sum().item()
    if n == 0:
        return new_mask
    prob = total_regrowth / n
    new_weights = torch.zeros_like(new_mask).bool()
    new_weights[new_mask == 0] = torch.rand_like(new_weights[new_mask == 0].float()) < prob
    new_mask = new_mask.bool() | new_weights.bool()
    weight.data[new_weights == 1] = 0.0
    weight.data[new_mask == 0] = 0.0
    return new_mask


def no_growth(masking, name, total_regrowth, weight):
    """grow.py:139-161."""
    return masking.mask_dict[name].data.bool()


registry = {"absolute-gradient": abs_grad_growth, "momentum": momentum_growth, "none": no_growth,
            "random": random_growth}
