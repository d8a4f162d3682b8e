"""GPU parity tests of the KL-autoencoder decode (run with `-m gpu` on a B200): the CUDA path through
the C ABI (ddb_ae_*) against oracle/ae_oracle.py (fp32, pinned to the reference by tests/golden/ae_decode_tiny.npz).

Tolerances (NHWC bf16 activations between layers, bf16 tensor-core operands, fp32 accumulation and GroupNorm
statistics; stated after measurement on a B200, measured values in brackets):
  * every tapped layer, free-running (errors accumulate over ~30 bf16 layers) : rel-L2 <= 2.5e-2 [1.7e-3 .. 1.24e-2]
  * final image                                                               : rel-L2 <= 2e-2   [9.4e-3 .. 9.9e-3]
                                                                                max-abs <= 3e-2 * max|ref| [1.1e-2]
  * chunked == unchunked, sample alone == sample in a batch                   : bit-exact (the GroupNorm partials are
                                                                                reduced in a fixed order)
"""
import numpy as np
import pytest
import torch

from oracle import ae_oracle as A
from tests.helpers import CONFIGS, rel_l2

pytestmark = pytest.mark.gpu

LAYER_REL_L2 = 2.5e-2
IMG_REL_L2 = 2e-2
IMG_MAX_ABS = 3e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_num_threads(16)
    return torch.device("cuda:0")


def _dd(spec):
    return dict(double_z=True, z_channels=spec.z_channels, resolution=spec.resolution, in_channels=3,
                out_ch=spec.out_ch, ch=spec.ch, ch_mult=spec.ch_mult, num_res_blocks=spec.num_res_blocks,
                attn_resolutions=[], dropout=0.0)


def _make(spec, seed, max_batch):
    from duodiff_b200.autoencoder import FrozenAutoencoderKL
    sd = A.random_state_dict(spec, seed)
    ae = FrozenAutoencoderKL(_dd(spec), spec.embed_dim, state_dict=sd, scale_factor=spec.scale_factor,
                             max_batch=max_batch)
    return ae, sd


def _latents(spec, B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, spec.z_channels, spec.z_res, spec.z_res, generator=g) * spec.scale_factor


def _check_image(out, ref):
    assert rel_l2(out, ref) <= IMG_REL_L2, rel_l2(out, ref)
    assert float((out - ref).abs().max()) <= IMG_MAX_ABS * float(ref.abs().max())


@pytest.mark.parametrize("spec,B", [
    (A.AESpec(ch=64, ch_mult=[1, 2], num_res_blocks=1, resolution=32), 2),       # 16x16 latents: 8-row pixel tiles
    (A.AESpec(ch=64, ch_mult=[1, 2, 2], num_res_blocks=1, resolution=128), 3),   # two upsamples, 128-wide rows
])
def test_decode_layer_by_layer(dev, spec, B):
    """Every layer the launch list exposes (conv_in, both convolutions and GroupNorms of every ResnetBlock incl. the
    nin_shortcut K-extension, the AttnBlock, the fused upsample convolutions, norm_out) against the oracle's taps."""
    ae, sd = _make(spec, 3, B)
    z = _latents(spec, B, 4)
    tap = {}
    ref = A.decode(sd, spec, z, tap)
    zc = z.to(dev)
    seen = set()
    for i, (name, c, h, w, _f32) in enumerate(ae.ops()):
        if c == 0 or name not in tap:
            continue
        _img, dump = ae.decode_debug(zc, i)
        t = tap[name]
        assert dump.shape[2:] == t.shape[2:], name
        e = rel_l2(dump.cpu()[:, :t.shape[1]], t)
        assert e <= LAYER_REL_L2, (name, e)
        if dump.shape[1] > t.shape[1]:  # zero padding of the 64-channel latent
            assert float(dump[:, t.shape[1]:].abs().max()) == 0.0
        seen.add(name.rsplit(".", 1)[-1])
    assert {"conv_in", "norm1", "conv1", "norm2", "conv2+res", "conv2+nin", "pv", "proj_out+res", "upsample",
            "norm_out"} <= seen
    _check_image(ae.decode(zc).cpu(), ref)


def test_reference_default_config_matches_oracle_and_is_batch_invariant(dev):
    """get_autoencoder's ddconfig (ch 128, ch_mult [1,2,4,4], 32x32x4 latents -> 3x256x256): B = 3 decoded in chunks of
    2 against the oracle; chunking and batch position must not change a single bit."""
    spec = A.AESpec()
    ae, sd = _make(spec, 7, 2)
    z = _latents(spec, 3, 8)
    zc = z.to(dev)
    out = ae.decode(zc)
    assert out.shape == (3, 3, 256, 256) and bool(torch.isfinite(out).all())
    ref = A.decode(sd, spec, z[:2])
    _check_image(out[:2].cpu(), ref)
    alone = ae.decode(zc[2:3])
    assert torch.equal(alone[0], out[2])
    swapped = ae.decode(zc[[1, 0]])
    assert torch.equal(swapped[0], out[1]) and torch.equal(swapped[1], out[0])
    prof = ae.profile_decode(zc)
    assert abs(sum(v["flops"] for v in prof.values()) / 3 / 1e9 - 622.2) < 1.0  # algorithmic GFLOP / image
    assert all(v["ms"] >= 0 for v in prof.values())


def test_unsupported_geometry_fails_loudly(dev):
    """The golden fixture's 32-channel toy decoder is below the kernel's 64-channel K chunk: rejected, no fallback."""
    from duodiff_b200._lib import DuoDiffError
    spec = A.AESpec(ch=32, ch_mult=[1, 2, 2], num_res_blocks=1, resolution=32)
    with pytest.raises(DuoDiffError, match="autoencoder:"):
        _make(spec, 0, 1)
    ok = A.AESpec(ch=64, ch_mult=[1, 2], num_res_blocks=1, resolution=32)
    ae, _sd = _make(ok, 0, 1)
    with pytest.raises(DuoDiffError, match="expects"):
        ae.decode(torch.zeros(1, 4, 8, 8, device=dev))
    sd = A.random_state_dict(ok, 0)
    del sd["decoder.mid.attn_1.k.weight"]
    from duodiff_b200.autoencoder import FrozenAutoencoderKL
    with pytest.raises(DuoDiffError, match="missing"):
        FrozenAutoencoderKL(_dd(ok), 4, state_dict=sd)


def test_get_samples_decodes_latents(dev):
    """sampler.get_samples(autoencoder=...) (sampler.py:141-153): the final and the saved intermediate latents go
    through decode before `(x + 1) / 2` and the NHWC permutation; eesampler.get_samples likewise (eesampler.py:84-88)."""
    import duodiff_b200 as ddb
    from duodiff_b200 import eesampler, sampler
    spec = A.AESpec(ch=64, ch_mult=[1, 2], num_res_blocks=1, resolution=64)  # 32x32x4 latents like ImageNet-256
    ae, _sd = _make(spec, 11, 2)
    torch.manual_seed(5)
    net = ddb.UViT(**CONFIGS["imagenet256_3"]).eval().to(dev)
    y = torch.tensor([3, 977], device=dev)
    kw = dict(model=net, batch_size=2, postprocessing=sampler.predict_noise_postprocessing, seed=1, num_channels=4,
              sample_height=32, sample_width=32, use_ddim=True, ddim_steps=20, ddim_eta=0.0, y=y)
    lat, lat_mid = sampler.get_samples(timesteps_save=[475], **kw)
    img, img_mid = sampler.get_samples(timesteps_save=[475], autoencoder=ae, **kw)
    assert lat.shape == (2, 32, 32, 4) and img.shape == (2, 64, 64, 3) and len(img_mid) == len(lat_mid) == 1

    def expect(l):
        z = torch.from_numpy(l).to(dev).permute(0, 3, 1, 2).contiguous() * 2 - 1
        return ((ae.decode(z) + 1) / 2).permute(0, 2, 3, 1).cpu().numpy()

    np.testing.assert_allclose(img, expect(lat), atol=2e-2)  # (x+1)/2 and back is not bit-exact in fp32
    np.testing.assert_allclose(img_mid[0], expect(lat_mid[0]), atol=2e-2)

    ee = ddb.EarlyExitUViT(ddb.UViT(**CONFIGS["imagenet256_3"]), "mlp_probe_per_layer").eval().to(dev)
    s, err_log, idx_log = eesampler.get_samples(model=ee, batch_size=2, seed=0, num_channels=4, sample_height=32,
                                                sample_width=32, threshold=0.0, depth=3, y=y, autoencoder=ae)
    assert s.shape == (2, 64, 64, 3) and err_log.shape == (1000, 3) and idx_log.shape == (1000, 2)
    assert np.isfinite(s).all()
