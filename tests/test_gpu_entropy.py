"""GPU tests (-m gpu) of the device-side packing for the entropy coder against the exact bytes the
unmodified reference wrote (pipeline/entropy_coding/__init__.py with the 'plain' stream; tests/golden/entropy.npz)."""
import json

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _quantised_model_from_golden(g):
    """The reference's converted k-means model, rebuilt from its recorded state dict."""
    from implicit_image_compression_b200.models import Siren
    torch.manual_seed(0)
    model = Siren(depth=4, hidden_size=32, first_omega_0=50, hidden_omega_0=30)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    for name, module in model.named_modules():
        if isinstance(module, nn.Linear) and f"{name}.centroids" in sd:
            module.centroids = nn.Parameter(sd[f"{name}.centroids"].clone(), requires_grad=False)
            module.labeled_weight = nn.Parameter(sd[f"{name}.labeled_weight"].clone(), requires_grad=False)
    model.load_state_dict(sd)
    return model.cuda(), sd


def test_packed_stream_is_byte_identical_to_the_reference(golden, tmp_path):
    from implicit_image_compression_b200.pipeline import entropy_coding as ec
    g = golden("entropy.npz")
    model, sd = _quantised_model_from_golden(g)
    stream, meta = ec.pack_state_dict(model)
    assert stream.is_cuda and stream.dtype == torch.uint8
    assert np.array_equal(stream.cpu().numpy(), g["bytes"])
    want_meta = json.loads(str(g["meta_json"]))
    assert {int(k): v for k, v in want_meta.items()} == {k: {"shape": v["shape"], "dtype": v["dtype"], "name": v["name"]}
                                                         for k, v in meta.items()}
    # files through the reference-shaped interface
    size = ec.compress_state_dict(model, tmp_path, stream_name="plain")
    assert size == g["bytes"].size
    assert (tmp_path / "compressed_weights.data").read_bytes() == g["bytes"].tobytes()
    dec = ec.decompress_state_dict(tmp_path, stream_name="plain")
    for k in g.files:
        if k.startswith("dec/"):
            assert torch.equal(dec[k[4:]], torch.from_numpy(g[k])), k
    # lzma round trip (host coder around the same device stream)
    ec.compress_state_dict(model, tmp_path / "lz", stream_name="lzma")
    dec2 = ec.decompress_state_dict(tmp_path / "lz", stream_name="lzma")
    assert all(torch.equal(dec[k], dec2[k]) for k in dec)


def test_device_unpack_decodes_straight_into_a_model(golden):
    """decompress -> image without a host round trip: codes + fp16 code book -> fp32 weights on the device."""
    from implicit_image_compression_b200.data import get_grid
    from implicit_image_compression_b200.models import Siren
    from implicit_image_compression_b200.pipeline import entropy_coding as ec
    g = golden("entropy.npz")
    model, sd = _quantised_model_from_golden(g)
    stream, meta = ec.pack_state_dict(model)
    torch.manual_seed(5)
    fresh = Siren(depth=4, hidden_size=32, first_omega_0=50, hidden_omega_0=30).cuda()
    ec.unpack_into_model(stream, meta, fresh)
    for k in g.files:
        if k.startswith("dec/"):
            got = dict(fresh.named_parameters())[k[4:]].detach().cpu()
            assert torch.equal(got, torch.from_numpy(g[k])), k
    grid = get_grid(12, 12, "cuda")
    with torch.no_grad():
        a = fresh(grid)
        model.half().float()  # what the reference evaluates after compress.py:247
        b = model(grid)
    assert (a - b).abs().max().item() <= 1e-6
