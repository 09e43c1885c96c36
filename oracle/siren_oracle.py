"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32, explicit formulae, no autograd, no nn.Module)
of the SIREN per-image fit hot path of varun19299/implicit-image-compression.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
file; the product package never does (it fails loudly if its CUDA library is missing).

Pinning: the reference ships NO tests or golden vectors for this path.  This restatement is
pinned against outputs of the reference itself, imported unmodified from /root/reference in the build
container by tools/make_golden.py; the resulting vectors are committed under tests/golden/ and
tests/test_oracle_golden.py checks every function here against them (bit-exact for masks, k-means codes and
int8 codes; <= 2e-6 relative for floating point, the difference being fp32 summation order only).

Each function cites the reference file:line it restates.
"""
import math

import numpy as np
import torch

F32 = torch.float32


# ----------------------------------------------------------------------------------------------------
# data
# ----------------------------------------------------------------------------------------------------
def get_grid(height, width):
    """implicit_image/data.py:78-88 — unit-square pixel coordinates, [H, W, 2], 'ij' indexing."""
    ch = torch.linspace(0, 1, height)
    cw = torch.linspace(0, 1, width)
    gh = ch[:, None].expand(height, width)
    gw = cw[None, :].expand(height, width)
    return torch.stack([gh, gw], dim=-1).contiguous()


def synth_image(height, width, idx=0, bits=16):
    """SURVEY.md §8(d) synthetic target: 12 sinusoids + ramp + 4 rectangles per channel, min-max
    normalised, quantised to `bits` and divided by 2^bits-1 (data.py:54 normalisation)."""
    g = torch.Generator().manual_seed(1000 + idx)
    u = torch.linspace(0, 1, height, dtype=torch.float64)[:, None]
    v = torch.linspace(0, 1, width, dtype=torch.float64)[None, :]
    img = torch.zeros(height, width, 3, dtype=torch.float64)
    for c in range(3):
        f = torch.rand(12, generator=g, dtype=torch.float64) * 23.5 + 0.5
        gq = torch.rand(12, generator=g, dtype=torch.float64) * 23.5 + 0.5
        ph = torch.rand(12, generator=g, dtype=torch.float64) * 2 * math.pi
        amp = 1.0 / (1.0 + torch.sqrt(f * f + gq * gq))
        ch = torch.zeros(height, width, dtype=torch.float64)
        for k in range(12):
            ch += amp[k] * torch.sin(2 * math.pi * (f[k] * u + gq[k] * v) + ph[k])
        ramp = torch.rand(2, generator=g, dtype=torch.float64) - 0.5
        ch += 0.3 * (ramp[0] * u + ramp[1] * v)
        for _ in range(4):
            r = torch.rand(5, generator=g, dtype=torch.float64)
            h0, w0 = int(r[0] * height * 0.8), int(r[1] * width * 0.8)
            h1, w1 = h0 + max(1, int(r[2] * height * 0.2)), w0 + max(1, int(r[3] * width * 0.2))
            ch[h0:h1, w0:w1] = (r[4] - 0.5) * 0.5
        img[:, :, c] = ch
    lo, hi = img.min(), img.max()
    img = (img - lo) / (hi - lo)
    q = torch.round(img * (2 ** bits - 1))
    return (q / (2 ** bits - 1)).to(F32)


# ----------------------------------------------------------------------------------------------------
# model: init, forward, backward
# ----------------------------------------------------------------------------------------------------
def siren_layer_dims(depth, hidden_size, input_size=2, output_size=3, small_dense_density=1.0):
    """implicit_image/models/siren.py:88-118 — hidden = int(hidden*sqrt(density)); depth layers."""
    hidden = int(hidden_size * np.sqrt(small_dense_density))
    dims = [(input_size, hidden)] + [(hidden, hidden)] * (depth - 2) + [(hidden, output_size)]
    return dims


def siren_init(seed, depth, hidden_size, first_omega_0=50.0, hidden_omega_0=30.0, input_size=2,
               output_size=3, small_dense_density=1.0):
    """implicit_image/models/siren.py:35-54,72-121 — draws parameters with the SAME torch RNG stream as the
    reference: nn.Linear's default init (kaiming_uniform(a=sqrt(5)) weight, then U(-1/sqrt(in), 1/sqrt(in))
    bias), then weight.uniform_(-b, b) with b = 1/in (first) or sqrt(6/in)/omega (others).
    Returns [w0, b0, w1, b1, ...] (fp32, [out, in] row-major)."""
    torch.manual_seed(seed)
    params = []
    for li, (fin, fout) in enumerate(siren_layer_dims(depth, hidden_size, input_size, output_size,
                                                      small_dense_density)):
        w = torch.empty(fout, fin, dtype=F32)
        b = torch.empty(fout, dtype=F32)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound_b = 1 / math.sqrt(fin)
        torch.nn.init.uniform_(b, -bound_b, bound_b)
        omega = first_omega_0 if li == 0 else hidden_omega_0
        bound = 1 / fin if li == 0 else np.sqrt(6 / fin) / omega
        w.uniform_(-bound, bound)
        params += [w, b]
    return params


def _omegas(depth, first_omega_0, hidden_omega_0):
    return [first_omega_0] + [hidden_omega_0] * (depth - 1)


def siren_forward(params, grid, first_omega_0=50.0, hidden_omega_0=30.0, outermost_linear=True,
                  return_intermediates=False):
    """implicit_image/models/siren.py:56-68 (SineLayer.forward) and :123-134 (Siren.forward):
    x = (grid-0.5)*2; per layer z = x W^T + b, a = sin(omega z) (last layer linear when
    outermost_linear); out = y/2 + 0.5."""
    h, w, _ = grid.shape
    depth = len(params) // 2
    om = _omegas(depth, first_omega_0, hidden_omega_0)
    x = (grid.reshape(h * w, -1) - 0.5) * 2
    acts, zs = [x], []
    for li in range(depth):
        z = torch.addmm(params[2 * li + 1], x, params[2 * li].t())
        zs.append(z)
        last = li == depth - 1
        x = z if (last and outermost_linear) else torch.sin(z * om[li])
        acts.append(x)
    pred = (x / 2 + 0.5).reshape(h, w, -1)
    if return_intermediates:
        return pred, acts, zs
    return pred


def siren_loss_and_grads(params, grid, img, first_omega_0=50.0, hidden_omega_0=30.0,
                         outermost_linear=True, loss_scale=1.0):
    """train_helper.py:148-161 (F.mse_loss + backward) restated with the explicit chain rule:
    L = mean((pred-img)^2); dL/dy = (pred-img)/(N*C) (mse' = 2/(N*C), pred = y/2+0.5);
    per layer dz = da * omega*cos(omega z); dW = dz^T x; db = sum dz; dx = dz W.
    Returns (loss, [dW0, db0, ...]) with grads multiplied by loss_scale (GradScaler.scale)."""
    depth = len(params) // 2
    om = _omegas(depth, first_omega_0, hidden_omega_0)
    pred, acts, zs = siren_forward(params, grid, first_omega_0, hidden_omega_0, outermost_linear, True)
    n = img.numel()
    diff = (pred - img).reshape(-1, pred.shape[-1])
    loss = (diff * diff).mean()
    g = diff * (loss_scale / n)  # dL/dy (the 2 of mse' cancels the 1/2 of y/2+0.5)
    grads = [None] * len(params)
    for li in reversed(range(depth)):
        last = li == depth - 1
        if not (last and outermost_linear):
            g = g * (om[li] * torch.cos(zs[li] * om[li]))
        grads[2 * li] = g.t() @ acts[li]
        grads[2 * li + 1] = g.sum(0)
        if li > 0:
            g = g @ params[2 * li]
    return loss, grads


def eval_metrics(pred, img):
    """train_helper.py:48-57 — (mse, PSNR, PSNR_8bit); the 8-bit images use .int() (truncation)."""
    mse = ((pred - img) ** 2).mean()
    psnr = 10 * torch.log10(1 / mse)
    d8 = (img * 255).int() - (pred * 255).int()
    mse8 = (d8 ** 2).float().mean()
    psnr8 = 10 * torch.log10(255 ** 2 / mse8)
    return mse.item(), psnr.item(), psnr8.item()


# ----------------------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------------------
def steplr(base_lr, step_index, period=2000, gamma=0.5):
    """train_helper.py:80-84 — StepLR: lr used by optimizer step number `step_index` (0-based)."""
    return base_lr * gamma ** (step_index // period)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (train_helper.py:72-78, defaults betas=(0.9,0.999), eps=1e-8, wd=0), one tensor:
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
    `step` is the 1-based step count AFTER increment.  Returns new (p, m, v)."""
    m = m + (g - m) * (1 - beta1)  # torch uses lerp_
    v = v * beta2 + (g * g) * (1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# ----------------------------------------------------------------------------------------------------
# masking
# ----------------------------------------------------------------------------------------------------
def apply_mask(weight, mask):
    """pipeline/masking/core.py:272-279 — w <- (w * mask).to(dtype)."""
    return (weight * mask).to(weight.dtype)


def erk_densities(shapes, density):
    """pipeline/masking/funcs/init_scheme.py:40-142 (get_erdos_renyi_dist, is_kernel=True):
    shapes = {name: shape}; returns {name: prob}."""
    dense = set()
    while True:
        divisor, rhs, raw = 0.0, 0.0, {}
        for name, shape in shapes.items():
            n_param = np.prod(shape)
            n_zeros = int(n_param * (1 - density))
            n_ones = int(n_param * density)
            if name in dense:
                rhs -= n_zeros
            else:
                rhs += n_ones
                raw[name] = (np.sum(shape) / np.prod(shape)) ** 1.0
                divisor += raw[name] * n_param
        eps = rhs / divisor
        max_prob = np.max(list(raw.values()))
        if max_prob * eps > 1:
            for name, rp in raw.items():
                if rp == max_prob:
                    dense.add(name)
        else:
            break
    return {name: (1.0 if name in dense else eps * raw[name]) for name in shapes}


def magnitude_prune(weight, mask, prune_rate, nonzeros, zeros):
    """pipeline/masking/funcs/prune.py:24-51 — zero the k = zeros + ceil(rate*nonzeros) smallest |w|."""
    num_remove = math.ceil(prune_rate * nonzeros)
    if num_remove == 0.0:
        return mask
    k = zeros + num_remove
    _, idx = torch.sort(torch.abs(weight.reshape(-1)))
    mask = mask.clone()
    mask.view(-1)[idx[:k]] = 0.0
    return mask


def global_magnitude_threshold(weights, nonzeros, prune_rate, baseline_nonzero, threshold,
                               increment=0.2, tolerance=1e-6):
    """pipeline/masking/funcs/prune.py:54-104 — the multiplicative threshold search, including its
    persistent `prune_threshold` state and the 10-stalled-tries exit.  Returns (threshold, removed)."""
    tokill = math.ceil(prune_rate * baseline_nonzero)
    if tokill <= 0:
        return threshold, 0
    total_removed, prev_removed, tries = 0, 0, 0
    while abs(total_removed - tokill) > tokill * tolerance:
        total_removed = 0
        for w, nz in zip(weights, nonzeros):
            remain = (torch.abs(w) > threshold).sum().item()
            total_removed += nz - remain
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


def abs_grad_growth(mask, grad, weight, total_regrowth):
    """pipeline/masking/funcs/grow.py:58-97 — grow where |grad| is largest among masked-out weights;
    new weights start at 0.  Returns (new_mask_bool, new_weight)."""
    new_mask = mask.bool().clone()
    if (new_mask == 0).sum().item() == 0:
        return new_mask, weight
    g = grad * (new_mask == 0).float()
    _, idx = torch.sort(torch.abs(g).flatten(), descending=True)
    k = int(total_regrowth)
    new_mask.view(-1)[idx[:k]] = True
    weight = weight.clone()
    weight.view(-1)[idx[:k]] = 0.0
    return new_mask, weight


def cosine_prune_rate(prune_rate, t_max, step):
    """pipeline/masking/funcs/decay.py:25-69 — CosineAnnealingLR closed form at epoch `step`."""
    return 0.5 * prune_rate * (1 + math.cos(math.pi * step / t_max))


def magnitude_prune_decay_rate(step, current_sparsity, final_sparsity, t_max, t_start, interval,
                               initial_sparsity=0.0):
    """pipeline/masking/funcs/decay.py:113-158 — Zhu & Gupta cubic schedule, finite difference."""

    def cum(s):
        if s < t_start:
            return initial_sparsity
        if s < t_max:
            mul = (1 - (s - t_start) / (t_max - t_start)) ** 3
            return final_sparsity + (initial_sparsity - final_sparsity) * mul
        return final_sparsity

    if current_sparsity == -1:
        current_sparsity = cum(step - interval)
    return max(cum(step) - current_sparsity, 0)


# ----------------------------------------------------------------------------------------------------
# quantisation
# ----------------------------------------------------------------------------------------------------
def _sqdist(a, b):
    """pipeline/quant/kmeans_helper.py:10-21 — pairwise squared distance for 1-D features."""
    return (a.reshape(-1, 1) - b.reshape(1, -1)) ** 2.0


def kmeans_quantize(weight, bits, iter_limit=5, tolerance=1e-4):
    """pipeline/quant/kmeans.py:110-150 (find_centroids) + kmeans_helper.py:59-116 (kmeans_fit/predict,
    torch_scatter.scatter_mean restated as sum/count with empty clusters -> 0):
    cluster the NON-ZERO weights into 2^bits-1 centroids from a linspace(min,max) guess (<=5 Lloyd
    iterations, stop when (sum_i |c_i - c'_i|)^2 < 1e-4), prepend centroid 0, unique, sort by |c|, then
    assign ALL weights by first-min squared distance.  Returns (centroids[k], labels int64, new_weight)."""
    shape = weight.shape
    w = weight.reshape(-1)
    nz = w[w != 0]
    k = 2 ** bits - 1
    centers = torch.linspace(nz.min().item(), nz.max().item(), k, dtype=weight.dtype)
    for _ in range(iter_limit):
        labels = torch.argmin(_sqdist(nz, centers), dim=1)
        n_out = int(labels.max().item()) + 1
        sums = torch.zeros(n_out, dtype=weight.dtype).index_add_(0, labels, nz)
        cnts = torch.zeros(n_out, dtype=weight.dtype).index_add_(0, labels, torch.ones_like(nz))
        new_centers = sums / cnts.clamp(min=1)
        shift = torch.sqrt((centers[:n_out] - new_centers) ** 2).sum() if n_out == centers.numel() \
            else torch.sqrt((centers - new_centers) ** 2).sum()  # shape error in the reference too
        centers = new_centers
        if shift ** 2 < tolerance:
            break
    cent = torch.cat([torch.zeros(1, dtype=weight.dtype), centers])
    cent = torch.unique(cent)
    _, order = torch.sort(cent.abs())
    cent = cent[order]
    labels = torch.argmin(_sqdist(w, cent), dim=1).reshape(shape)
    return cent, labels, cent[labels]


def fake_quant_per_channel_weight(weight, row_min=None, row_max=None, neg_div=128.0, pos_div=127.0):
    """Per-channel weight fake-quant used by get_default_qat_qconfig('fbgemm') (pipeline/quant/context.py:
    35-47): per-output-channel symmetric qint8, quant range [-128, 127], zero point 0,
      scale = max(-min(lo,0)/neg_div, max(hi,0)/pos_div, eps), q = clamp(rne(w * (1/scale)), -128, 127).
    The installed torch (2.11) runs the FUSED observer kernel (fbgemm ChooseQuantizationParams with
    preserve_sparsity): (neg_div, pos_div) = (128, 127).  torch 1.7 (the reference's pin) used the Python
    observer formula max(|lo|, hi) / 127.5, i.e. (127.5, 127.5).
    Returns (codes int8, scale[out], dequantised weight)."""
    lo = weight.min(dim=1).values if row_min is None else row_min
    hi = weight.max(dim=1).values if row_max is None else row_max
    zero = torch.zeros_like(lo)
    scale = torch.maximum(-torch.minimum(lo, zero) / neg_div, torch.maximum(hi, zero) / pos_div)
    scale = scale.clamp(min=torch.finfo(torch.float32).eps)
    q = torch.clamp(torch.round(weight * (1.0 / scale)[:, None]), -128, 127)  # torch.round = RNE
    return q.to(torch.int8), scale, q * scale[:, None]


# ----------------------------------------------------------------------------------------------------
# QAT activation fake-quant
# ----------------------------------------------------------------------------------------------------
def fused_obs_fake_quant(x, state, training=True, averaging_const=0.01, qmin=0, qmax=127):
    """What torch.quantization.prepare_qat puts on every nn.Linear output of the reference model
    (pipeline/quant/context.py:35-47 -> torch FusedMovingAvgObsFakeQuantize; third-party arithmetic: PyTorch,
    aten/src/ATen/native/quantized/cpu/fused_obs_fake_quant.cpp + fbgemm's ChooseQuantizationParams), restated:
    state = [running min, running max, scale, zero point] (python floats, running min/max start at +-inf).
    Returns (fake_quant(x), mask, new_state).  Pinned against the installed torch by tests/test_oracle_golden.py."""
    f32 = np.float32
    rmin, rmax = state[0], state[1]
    if training:
        cmin, cmax = f32(x.min().item()), f32(x.max().item())
        if math.isinf(rmin) or math.isinf(rmax):
            rmin, rmax = cmin, cmax
        else:
            rmin = f32(f32(rmin) + f32(averaging_const) * f32(cmin - f32(rmin)))
            rmax = f32(f32(rmax) + f32(averaging_const) * f32(cmax - f32(rmax)))
    mn, mx = min(float(rmin), 0.0), max(float(rmax), 0.0)
    scale = (mx - mn) / (qmax - qmin)
    if f32(scale) == 0.0 or math.isinf(1.0 / f32(scale)):
        scale = 0.1
    small = 6.1e-5
    if scale < small:
        org = f32(scale)
        scale = small
        if mn == 0.0:
            mx = small * (qmax - qmin)
        elif mx == 0.0:
            mn = -small * (qmax - qmin)
        else:
            amp = f32(small / org)
            mn *= amp
            mx *= amp
    zmin, zmax = qmin - mn / scale, qmax - mx / scale
    emin, emax = abs(qmin) - abs(mn / scale), abs(qmax) - abs(mx / scale)
    z0 = zmin if emin < emax else zmax
    zp = qmin if z0 < qmin else (qmax if z0 > qmax else int(np.rint(z0)))
    s32 = f32(scale)
    inv = f32(1.0) / s32
    q = np.rint(x.detach().numpy().astype(f32) * inv) + f32(zp)
    mask = (q >= qmin) & (q <= qmax)
    out = (np.clip(q, qmin, qmax) - f32(zp)).astype(f32) * s32
    return torch.from_numpy(out), torch.from_numpy(mask), [float(rmin), float(rmax), float(s32), float(zp)]


# ----------------------------------------------------------------------------------------------------
# FourierNet
# ----------------------------------------------------------------------------------------------------
def fourier_forward(B, params, grid, keep=False):
    """implicit_image/models/fourier.py:20-25,58-69 — x = grid.view(N, 2) (RAW [0,1] coordinates);
    enc = [sin(2 pi x B) | cos(2 pi x B)]; (Linear, ReLU) x (depth - 2); Linear; Sigmoid.
    params = [w0, b0, w1, b1, ...] of the depth - 1 nn.Linear layers."""
    h, w, _ = grid.shape
    x = grid.reshape(-1, 2).to(F32)
    xp = (2 * np.pi * x) @ B
    a = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
    nlin = len(params) // 2
    zs, acts = [], [a]
    for l in range(nlin):
        z = torch.addmm(params[2 * l + 1], a, params[2 * l].t())
        zs.append(z)
        if l == nlin - 1:
            pred = torch.sigmoid(z)
            break
        a = torch.relu(z)
        acts.append(a)
    out = pred.reshape(h, w, -1)
    return (out, zs, acts) if keep else out


def fourier_loss_and_grads(B, params, grid, img):
    """F.mse_loss + explicit backward of the chain above (utils/train_helper.py:151-161 on a FourierNet)."""
    pred, zs, acts = fourier_forward(B, params, grid, keep=True)
    d = pred.reshape(-1, pred.shape[-1]) - img.reshape(-1, img.shape[-1])
    loss = (d * d).mean()
    p = pred.reshape(d.shape)
    g = (2.0 / d.numel()) * d * p * (1 - p)
    nlin = len(params) // 2
    grads = [None] * (2 * nlin)
    for l in range(nlin - 1, -1, -1):
        grads[2 * l] = g.t() @ acts[l]
        grads[2 * l + 1] = g.sum(0)
        if l > 0:
            g = (g @ params[2 * l]) * (zs[l - 1] > 0).to(F32)
    return loss, grads
