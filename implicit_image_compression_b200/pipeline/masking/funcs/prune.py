"""Prune rules (reference: pipeline/masking/funcs/prune.py).  Only the rules the shipped configs use on a
SIREN: per-layer `magnitude` (RigL / SNFS) and `global-magnitude` (Pruning).  The conv-only struct-* rules
never apply to 2-D weights and are out of scope."""
import math

import numpy as np
import torch


def magnitude_prune(masking, mask, weight, name):
    """prune.py:24-51: zero the k = zeros + ceil(rate * nonzeros) smallest |w| of the layer."""
    num_remove = math.ceil(masking.name2prune_rate[name] * masking.stats.nonzeros_dict[name])
    if num_remove == 0.0:
        return mask
    k = masking.stats.zeros_dict[name] + num_remove
    _, order = torch.sort(torch.abs(weight.data.view(-1)))
    mask.data.view(-1)[order[:k]] = 0.0
    return mask


def global_magnitude_prune(masking):
    """prune.py:54-104: search one global |w| threshold multiplicatively until the number of removed
    weights matches ceil(prune_rate * baseline_nonzero) within `tolerance`, or ten stalled tries.  The
    threshold persists on the Masking object between calls, so the trajectory (and hence the final mask)
    is part of the behaviour and is replicated step for step."""
    tokill = math.ceil(masking.prune_rate * masking.baseline_nonzero)
    if tokill <= 0:
        return 0
    masked = [(n, w) for n, w in masking.module.named_parameters() if n in masking.mask_dict]
    # The reference probes `(|w| > threshold).sum()` per layer per iteration (hundreds of probes, each a
    # kernel launch + host sync).  The same integer comes from ONE device sort: with the magnitudes sorted,
    # #(|w| > t) = n - upper_bound(t).  torch compares an fp32 tensor with a Python float in fp32, so the
    # threshold is rounded to fp32 for the probe exactly as `torch.abs(w) > threshold` rounds it.
    mags = torch.sort(torch.cat([torch.abs(w.data).reshape(-1) for _, w in masked]))[0].cpu().numpy()
    mags = mags[~np.isnan(mags)]  # NaN > t is False
    nonzero_total = sum(masking.stats.nonzeros_dict[n] for n, _ in masked)
    total_removed = prev_removed = tries = 0
    increment = masking.increment
    while abs(total_removed - tokill) > tokill * masking.tolerance:
        remain = mags.size - int(np.searchsorted(mags, np.float32(masking.prune_threshold), side="right"))
        total_removed = nonzero_total - remain
        if prev_removed == total_removed:
            tries += 1
            if tries == 10:
                break
        else:
            tries = 0
        prev_removed = total_removed
        if total_removed > tokill * (1.0 + masking.tolerance):
            masking.prune_threshold *= 1.0 - increment
            increment *= 0.99
        elif total_removed < tokill * (1.0 - masking.tolerance):
            masking.prune_threshold *= 1.0 + increment
            increment *= 0.99
    for n, w in masked:
        masking.mask_dict[n][:] = torch.abs(w.data) > masking.prune_threshold
    return int(total_removed)


registry = {"global-magnitude": global_magnitude_prune, "magnitude": magnitude_prune}
