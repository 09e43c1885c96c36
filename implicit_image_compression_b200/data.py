"""Inputs of the fit: pixel grid (reference: implicit_image/data.py:78-88) and the seeded synthetic 16-bit
target images that replace the reference's dataset (the reference ships no images).
`load_img` (cv2 / kornia file reading, data.py:44-75) is out of scope; only its /(2^bits-1) normalisation
(data.py:54) is kept, in `synth_image`."""
import math

import torch


def get_grid(height, width, device=torch.device("cpu")):
    """[H, W, 2] coordinates in the unit square, (h, w) feature order, 'ij' indexing.  Built on the CPU with
    torch.linspace exactly like the reference so that the values are bit-identical, then moved."""
    ch = torch.linspace(0, 1, height)
    cw = torch.linspace(0, 1, width)
    grid = torch.stack([ch[:, None].expand(height, width), cw[None, :].expand(height, width)], dim=-1)
    return grid.contiguous().to(device)


def synth_image(height, width, idx=0, bits=16, device=torch.device("cpu")):
    """Deterministic smooth-plus-edges RGB target: per channel 12 sinusoids (0.5..24 cycles, 1/f
    amplitudes), a linear ramp and 4 constant rectangles; min-max normalised, quantised to `bits` and
    divided by 2^bits - 1 -> [H, W, 3] fp32 in [0, 1]."""
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
    return (q / (2 ** bits - 1)).to(torch.float32).to(device)
