"""Initial sparsity patterns (reference: pipeline/masking/funcs/init_scheme.py): `random`,
`erdos-renyi(-kernel)` and `resume`.  Masks are drawn with torch.rand on the CPU generator in
named_parameters() order, exactly like the reference, so a shared seed gives identical masks.
Lottery-ticket and the conv-only struct-* schemes are out of scope."""
from functools import partial

import numpy as np
import torch


def erdos_renyi_densities(shapes, density, is_kernel=True):
    """init_scheme.py:40-142.  shapes: {name: shape} of the masked tensors, in order."""
    dense_layers = set()
    while True:
        divisor, rhs, raw = 0, 0, {}
        for name, shape in shapes.items():
            n_param = np.prod(shape)
            n_zeros = int(n_param * (1 - density))
            n_ones = int(n_param * density)
            if name in dense_layers:
                rhs -= n_zeros
            else:
                rhs += n_ones
                if is_kernel:
                    raw[name] = (np.sum(shape) / np.prod(shape)) ** 1.0
                else:
                    n_in, n_out = shape[:2]
                    raw[name] = (n_in + n_out) / (n_in * n_out)
                divisor += raw[name] * n_param
        epsilon = rhs / divisor
        max_prob = np.max(list(raw.values()))
        if max_prob * epsilon > 1:
            for name, p in raw.items():
                if p == max_prob:
                    dense_layers.add(name)
        else:
            break
    return {name: (1.0 if name in dense_layers else epsilon * raw[name]) for name in shapes}


def erdos_renyi_init(masking, is_kernel=True, **kwargs):
    shapes = {n: tuple(m.shape) for n, m in masking.mask_dict.items()}
    probs = erdos_renyi_densities(shapes, masking.density, is_kernel)
    for name, weight in masking.module.named_parameters():
        if name not in masking.mask_dict:
            continue
        masking.mask_dict[name] = (torch.rand(weight.shape) < probs[name]).float().data
        masking.baseline_nonzero += (masking.mask_dict[name] != 0).sum().int().item()
        masking.total_params += weight.numel()


def random_init(masking, **kwargs):
    """init_scheme.py:188-212: the first enumerated parameter (layer-0 weight) is left dense."""
    for e, (name, weight) in enumerate(masking.module.named_parameters()):
        if e == 0:
            masking.remove_weight(name)
            continue
        if name not in masking.mask_dict:
            continue
        masking.mask_dict[name] = (torch.rand(weight.shape) < masking.density).float().data
        masking.baseline_nonzero += masking.mask_dict[name].sum().int().item()
        masking.total_params += weight.numel()


def resume_init(masking, **kwargs):
    """init_scheme.py:215-234: mask = (weight != 0)."""
    for name, weight in masking.module.named_parameters():
        if name not in masking.mask_dict:
            continue
        masking.mask_dict[name] = (weight != 0.0).float().data
        masking.baseline_nonzero += masking.mask_dict[name].sum().int().item()
        masking.total_params += weight.numel()


registry = {"erdos-renyi": partial(erdos_renyi_init, is_kernel=False),
            "erdos-renyi-kernel": partial(erdos_renyi_init, is_kernel=True),
            "random": random_init, "resume": resume_init}
