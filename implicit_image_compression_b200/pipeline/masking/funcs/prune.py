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


def _host_threshold_search(mags, nonzero_total, tokill, tolerance, threshold, increment):
    """The reference's search loop (prune.py:74-95) over magnitudes sorted once: #(|w| > t) = n - upper_bound(t),
    with t rounded to fp32 as `torch.abs(w) > threshold` rounds it.  Host reference of the device kernel
    (tests) and the path for CPU tensors.  Returns (threshold, total_removed)."""
    mags = mags[~np.isnan(mags)]  # NaN > t is False
    total_removed = prev_removed = tries = 0
    while abs(total_removed - tokill) > tokill * tolerance:
        remain = mags.size - int(np.searchsorted(mags, np.float32(threshold), side="right"))
        total_removed = nonzero_total - remain
        if prev_removed == total_removed:
            tries += 1
            if tries == 10:
                break
        else:
            tries = 0
        prev_removed = total_removed
        if total_removed > tokill * (1.0 + tolerance):
            threshold *= 1.0 - increment
            increment *= 0.99
        elif total_removed < tokill * (1.0 - tolerance):
            threshold *= 1.0 + increment
            increment *= 0.99
    return threshold, int(total_removed)


def global_magnitude_prune(masking):
    """prune.py:54-104: search one global |w| threshold multiplicatively until the number of removed
    weights matches ceil(prune_rate * baseline_nonzero) within `tolerance`, or ten stalled tries.  The
    threshold persists on the Masking object between calls, so the trajectory (and hence the final mask)
    is part of the behaviour and is replicated step for step.

    The reference probes `(|w| > threshold).sum()` per layer per iteration (hundreds of probes, each a kernel
    launch + host sync).  Here the magnitudes are sorted ONCE on the device and the whole search runs in one
    warp of `sirenb200_prune_threshold_search` (same IEEE-double trajectory); what comes back to the host is
    16 bytes: the new threshold (a Python float on the Masking object, as in the reference) and the count."""
    tokill = math.ceil(masking.prune_rate * masking.baseline_nonzero)
    if tokill <= 0:
        return 0
    masked = [(n, w) for n, w in masking.module.named_parameters() if n in masking.mask_dict]
    mags = torch.sort(torch.cat([torch.abs(w.data).reshape(-1) for _, w in masked]))[0]
    nonzero_total = sum(masking.stats.nonzeros_dict[n] for n, _ in masked)
    if mags.is_cuda:
        from .... import _lib
        state = torch.tensor([masking.prune_threshold, masking.increment, 0.0], dtype=torch.float64,
                             device=mags.device)
        with torch.cuda.device(mags.device):
            _lib.check(_lib.load().sirenb200_prune_threshold_search(
                mags.data_ptr(), mags.numel(), int(nonzero_total), int(tokill), float(masking.tolerance),
                state.data_ptr(), state.data_ptr() + 16, torch.cuda.current_stream().cuda_stream))
        threshold, _, removed = state.tolist()
        masking.prune_threshold, total_removed = threshold, int(removed)
    else:
        masking.prune_threshold, total_removed = _host_threshold_search(
            mags.numpy(), nonzero_total, tokill, masking.tolerance, masking.prune_threshold, masking.increment)
    for n, w in masked:
        masking.mask_dict[n][:] = torch.abs(w.data) > masking.prune_threshold
    return int(total_removed)


registry = {"global-magnitude": global_magnitude_prune, "magnitude": magnitude_prune}
