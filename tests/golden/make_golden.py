"""Generates the golden fixtures that pin ``oracle/`` to the reference implementation.

Run ONLY in the build container (needs the read-only reference checkout):

    python tests/golden/make_golden.py [/root/reference]

It imports the *unmodified* reference modules (``models.uvit``, ``models.early_exit``, ``sampler``, ``eesampler``)
with a stub ``matplotlib`` (not installed; only the PNG dump uses it), builds tiny random-init models with the
reference's own constructors, runs the reference forward / samplers on the CPU and stores weights + inputs + outputs
as compressed ``.npz`` files next to this script.  Nothing here is read at test time except the ``.npz`` files.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent

stub, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
stub.pyplot = pp
sys.modules.setdefault("matplotlib", stub)
sys.modules.setdefault("matplotlib.pyplot", pp)
sys.path.insert(0, str(REF))

import eesampler as ref_ee  # noqa: E402
import sampler as ref_sampler  # noqa: E402
from models.early_exit import EarlyExitUViT  # noqa: E402
from models.uvit import UViT  # noqa: E402

TINY = dict(img_size=8, patch_size=2, in_chans=3, embed_dim=32, depth=3, num_heads=2, mlp_ratio=4, qkv_bias=False,
            mlp_time_embed=False, num_classes=-1, normalize_timesteps=True)
TINY_FULL = dict(TINY, depth=5)
TINY_CLS = dict(TINY, num_classes=10, normalize_timesteps=False, in_chans=4, depth=5)


def heat(model, seed):
    """Random-init (std 0.02, zero bias, unit LN) leaves LayerNorm affine and biases untested: perturb them."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            elif "pos_embed" in name:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
            else:
                p.mul_(4.0)


def sd_np(model, prefix="w::"):
    return {prefix + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def params_np(params):
    return {f"p::{k}": np.asarray(v) for k, v in params.items()}


def uvit_forward_fixture(name, params, seed, with_y):
    torch.manual_seed(seed)
    m = UViT(**params).eval()
    heat(m, seed + 1)
    B = 3
    x = torch.randn(B, params["in_chans"], params["img_size"], params["img_size"])
    t = torch.tensor([999.0, 500.0, 3.0])
    y = torch.tensor([0, 9, 4]) if with_y else None
    hidden = []
    hooks = [blk.register_forward_hook(lambda _m, _i, o: hidden.append(o.detach().numpy().copy()))
             for blk in list(m.in_blocks) + [m.mid_block] + list(m.out_blocks)]
    with torch.no_grad():
        out = m(x, t, y)
    for h in hooks:
        h.remove()
    fx = dict(x=x.numpy(), t=t.numpy(), out=out.numpy(), **params_np(params), **sd_np(m))
    if with_y:
        fx["y"] = y.numpy()
    for i, h in enumerate(hidden):
        fx[f"hidden_{i}"] = h
    np.savez_compressed(OUT / f"{name}.npz", **fx)
    print(name, "out", tuple(out.shape), float(out.abs().max()))


def ee_forward_fixture():
    torch.manual_seed(11)
    m = EarlyExitUViT(UViT(**TINY_FULL), "mlp_probe_per_layer").eval()
    heat(m, 12)
    with torch.no_grad():  # spread the probes around the threshold so exits happen at different layers
        for i in range(TINY_FULL["depth"]):
            m.matrix[f"{i}"].classifier[0].weight.mul_(6.0)
            m.matrix[f"{i}"].classifier[0].bias.fill_(0.3 - 0.7 * i)
    B = 4
    x = torch.randn(B, 3, 8, 8)
    t = torch.full((B,), 321.0)
    with torch.no_grad():
        eps, cls, outs = m(x, t, None)
    fx = dict(x=x.numpy(), t=t.numpy(), eps=eps.numpy(), cls=torch.stack(cls).numpy(), outs=torch.stack(outs).numpy(),
              **params_np(TINY_FULL), **sd_np(m))
    np.savez_compressed(OUT / "ee_forward_tiny.npz", **fx)
    print("ee_forward", torch.stack(cls).numpy().round(3))
    return m


def duodiff_sampler_fixture():
    """Unmodified sampler.get_samples: DuoDiff hand-off at t_switch=300, predict_noise, CPU RNG stream."""
    torch.manual_seed(21)
    early = UViT(**TINY).eval()
    late = UViT(**TINY_FULL).eval()
    heat(early, 22)
    heat(late, 23)
    calls = {"early": [], "late": []}
    early.register_forward_hook(lambda _m, i, _o: calls["early"].append(int(i[1][0])))
    late.register_forward_hook(lambda _m, i, _o: calls["late"].append(int(i[1][0])))
    fx = dict(**params_np(TINY), **{f"q::{k}": np.asarray(v) for k, v in TINY_FULL.items()}, **sd_np(early, "we::"),
              **sd_np(late, "wl::"))
    for rule, fn in (("predict_noise", ref_sampler.predict_noise_postprocessing),
                     ("predict_original", ref_sampler.predict_original_postprocessing),
                     ("predict_previous", ref_sampler.predict_previous_postprocessing)):
        calls["early"].clear(), calls["late"].clear()
        samples, inter = ref_sampler.get_samples(
            model=early, batch_size=2, postprocessing=fn, seed=5, num_channels=3, sample_height=8, sample_width=8,
            use_ddim=False, ddim_steps=50, ddim_eta=0.0, timesteps_save=[1, 300, 990], y=None, autoencoder=None,
            late_model=late, t_switch=300)
        fx[f"samples_{rule}"] = samples
        for i, s in enumerate(inter):
            fx[f"inter_{rule}_{i}"] = s
        if rule == "predict_noise":
            fx["early_t_min"], fx["early_calls"] = min(calls["early"]), len(calls["early"])
            fx["late_t_max"], fx["late_calls"] = max(calls["late"]), len(calls["late"])
        print(rule, samples.shape, float(np.abs(samples).max()), len(inter), len(calls["early"]), len(calls["late"]))
    np.savez_compressed(OUT / "duodiff_sampler_tiny.npz", **fx)


def ddim_sampler_fixture():
    """Unmodified sampler.get_samples(use_ddim=True): strided steps, eta 0 and 0.5, hand-off at t_switch=300."""
    torch.manual_seed(21)
    early = UViT(**TINY).eval()
    late = UViT(**TINY_FULL).eval()
    heat(early, 22)
    heat(late, 23)
    calls = {"early": [], "late": []}
    early.register_forward_hook(lambda _m, i, _o: calls["early"].append(int(i[1][0])))
    late.register_forward_hook(lambda _m, i, _o: calls["late"].append(int(i[1][0])))
    fx = dict(**params_np(TINY), **{f"q::{k}": np.asarray(v) for k, v in TINY_FULL.items()}, **sd_np(early, "we::"),
              **sd_np(late, "wl::"))
    # (eta large enough makes sqrt(1 - abar_s - sigma^2) NaN near s = 0 in the reference: 20 steps, eta 0.5 -> NaN images)
    for steps, eta in ((50, 0.0), (20, 0.05)):
        calls["early"].clear(), calls["late"].clear()
        samples, inter = ref_sampler.get_samples(
            model=early, batch_size=2, postprocessing=ref_sampler.predict_noise_postprocessing, seed=7,
            num_channels=3, sample_height=8, sample_width=8, use_ddim=True, ddim_steps=steps, ddim_eta=eta,
            timesteps_save=[1, 1000 - int(np.linspace(0, 999, steps).astype(int)[::-1][3])], y=None, autoencoder=None,
            late_model=late, t_switch=300)
        key = f"{steps}_{eta}"
        fx[f"samples_{key}"] = samples
        for i, s in enumerate(inter):
            fx[f"inter_{key}_{i}"] = s
        fx[f"early_ts_{key}"] = np.asarray(calls["early"])
        fx[f"late_ts_{key}"] = np.asarray(calls["late"])
        print("ddim", key, samples.shape, float(np.abs(samples).max()), len(inter), calls["early"][-3:], calls["late"][:3])
    np.savez_compressed(OUT / "ddim_sampler_tiny.npz", **fx)


def ee_sampler_fixture(m):
    samples, err_log, idx_log = ref_ee.get_samples(model=m, batch_size=3, seed=9, num_channels=3, sample_height=8,
                                                   sample_width=8, threshold=0.35, depth=TINY_FULL["depth"])
    fx = dict(samples=samples, err_log=err_log.numpy(), idx_log=idx_log.numpy(), threshold=np.float32(0.35),
              **params_np(TINY_FULL), **sd_np(m))
    np.savez_compressed(OUT / "ee_sampler_tiny.npz", **fx)
    print("ee_sampler", samples.shape, "mean exit", float(idx_log.mean()))


def schedule_fixture():
    np.savez_compressed(OUT / "schedule.npz", betas=ref_sampler.betas.numpy(), alphas=ref_sampler.alphas.numpy(),
                        alphas_bar=ref_sampler.alphas_bar.numpy(),
                        alphas_bar_previous=ref_sampler.alphas_bar_previous.numpy(),
                        betas_tilde=ref_sampler.betas_tilde.numpy())


def ae_decode_fixture():
    """FrozenAutoencoderKL.decode of the unmodified reference (models/utils/autoencoder.py:452-490) on a tiny ddconfig:
    the module only loads weights from a file, so a random-init state_dict (GroupNorm affine and biases perturbed) is
    saved to a temporary path first.  Also runs the decode through sampler.get_samples(autoencoder=...) with a zero
    network so that the fixture pins the `(decode(x) + 1) / 2`, NHWC post-processing (sampler.py:141-146)."""
    import tempfile

    from models.utils.autoencoder import Decoder, Encoder, FrozenAutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
              num_res_blocks=1, attn_resolutions=[], dropout=0.0)
    torch.manual_seed(41)
    enc, dec = Encoder(**dd), Decoder(**dd)
    quant, post = torch.nn.Conv2d(8, 8, 1), torch.nn.Conv2d(4, 4, 1)
    g = torch.Generator().manual_seed(42)
    with torch.no_grad():
        for name, p in list(dec.named_parameters()) + list(post.named_parameters()):
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
    sd = {}
    for pfx, mod in (("encoder.", enc), ("decoder.", dec), ("quant_conv.", quant), ("post_quant_conv.", post)):
        sd.update({pfx + k: v for k, v in mod.state_dict().items()})
    with tempfile.TemporaryDirectory() as d:
        torch.save(sd, Path(d) / "ae.pth")
        ae = FrozenAutoencoderKL(dd, 4, str(Path(d) / "ae.pth"))
    z = torch.randn(2, 4, 8, 8, generator=g) * 0.18215 * 1.5
    with torch.no_grad():
        out = ae.decode(z)
    fx = dict(z=z.numpy(), out=out.numpy(), ch=np.int64(32), ch_mult=np.asarray(dd["ch_mult"]),
              num_res_blocks=np.int64(1), resolution=np.int64(32), scale_factor=np.float32(0.18215))
    fx.update({"w::" + k: v.numpy().copy() for k, v in sd.items()
               if k.startswith("decoder.") or k.startswith("post_quant_conv.")})
    np.savez_compressed(OUT / "ae_decode_tiny.npz", **fx)
    print("ae_decode", tuple(out.shape), float(out.abs().max()))


def ee_probe_types_fixture():
    """The two timestep-indexed MLP probe layouts (models/early_exit.py:199-202, 228-239): the reference model holds
    1000 (x depth) probes; the fixture keeps the backbone / heads and only the probes of the timesteps it runs.  The
    third call mixes timesteps inside the batch: int(timesteps[0]) picks the probes (early_exit.py:269)."""
    fx = {}
    calls = [torch.full((3,), 321.0), torch.full((3,), 7.0), torch.tensor([500.0, 3.0, 999.0])]
    for tag, ctype in (("pt", "mlp_probe_per_timestep"), ("plt", "mlp_probe_per_layer_per_timestep")):
        torch.manual_seed(21)
        m = EarlyExitUViT(UViT(**TINY), ctype).eval()
        heat(m, 22)
        used = sorted({int(t[0]) for t in calls})
        with torch.no_grad():
            for k, t in enumerate(used):
                keys = [f"{t}"] if tag == "pt" else [f"{i}, {t}" for i in range(TINY["depth"])]
                for j, key in enumerate(keys):
                    m.matrix[key].classifier[0].weight.mul_(6.0)
                    m.matrix[key].classifier[0].bias.fill_(0.4 - 0.5 * j - 0.2 * k)
        x = torch.randn(3, 3, 8, 8)
        for c, t in enumerate(calls):
            with torch.no_grad():
                eps, cls, outs = m(x, t, None)
            fx[f"{tag}::t{c}"] = t.numpy()
            fx[f"{tag}::eps{c}"] = eps.numpy()
            fx[f"{tag}::cls{c}"] = torch.stack(cls).numpy()
            fx[f"{tag}::outs{c}"] = torch.stack(outs).numpy()
        fx[f"{tag}::x"] = x.numpy()
        keep = lambda k: not k.startswith("matrix.") or any(  # noqa: E731
            k.startswith(f"matrix.{t}.") or k.split(".")[1].endswith(f", {t}") for t in used)
        for k, v in m.state_dict().items():
            if keep(k):
                fx[f"{tag}::w::{k}"] = v.detach().numpy().copy()
        print("ee_probe_types", ctype, torch.stack(cls).numpy().round(3).tolist())
    np.savez_compressed(OUT / "ee_probe_types_tiny.npz", **params_np(TINY), **fx)


def uvit_variants_fixture():
    """The constructor options no shipped config switches on (models/uvit.py:229-247): A = mlp_time_embed=True with a
    fractional timestep in the batch; B = conv=False + skip=False (+ qk_scale, which the reference ignores),
    class-conditional with raw timesteps."""
    fx = {}
    for tag, params, with_y in (("a", dict(TINY_FULL, mlp_time_embed=True), False),
                                ("b", dict(TINY_CLS, conv=False, skip=False, qk_scale=0.3), True)):
        torch.manual_seed(51)
        m = UViT(**params).eval()
        heat(m, 52)
        x = torch.randn(3, params["in_chans"], 8, 8)
        t = torch.tensor([999.0, 12.5, 3.0])
        y = torch.tensor([2, 9, 0]) if with_y else None
        with torch.no_grad():
            out = m(x, t, y)
        fx.update({f"{tag}::x": x.numpy(), f"{tag}::t": t.numpy(), f"{tag}::out": out.numpy()})
        if with_y:
            fx[f"{tag}::y"] = y.numpy()
        fx.update({f"{tag}::p::{k}": np.asarray(v) for k, v in params.items()})
        fx.update(sd_np(m, f"{tag}::w::"))
        print("uvit_variants", tag, tuple(out.shape), float(out.abs().max()), len(m.state_dict()))
    np.savez_compressed(OUT / "uvit_variants_tiny.npz", **fx)


def ee_attention_probe_fixture():
    """classifier_type = "attention_probe" (models/early_exit.py:40-80, the constructor's default).  The learned query
    is zero-initialised (uniform attention): it is given random values so that the softmax matters; one unconditional
    and one class-conditional model (two extra tokens: x[:, 1:] then drops the LABEL token)."""
    fx = {}
    for tag, params, with_y in (("u", TINY, False), ("c", TINY_CLS, True)):
        torch.manual_seed(41)
        m = EarlyExitUViT(UViT(**params), "attention_probe").eval()
        heat(m, 42)
        g = torch.Generator().manual_seed(43)
        with torch.no_grad():
            for i in range(params["depth"]):
                m.matrix[f"{i}"].q.copy_(torch.randn(m.matrix[f"{i}"].q.shape, generator=g) * 2.0)
        x = torch.randn(3, params["in_chans"], 8, 8)
        t = torch.tensor([321.0, 321.0, 321.0])
        y = torch.tensor([1, 7, 3]) if with_y else None
        with torch.no_grad():
            eps, cls, outs = m(x, t, y)
        fx.update({f"{tag}::x": x.numpy(), f"{tag}::t": t.numpy(), f"{tag}::eps": eps.numpy(),
                   f"{tag}::cls": torch.stack(cls).numpy(), f"{tag}::outs": torch.stack(outs).numpy()})
        if with_y:
            fx[f"{tag}::y"] = y.numpy()
        fx.update({f"{tag}::p::{k}": np.asarray(v) for k, v in params.items()})
        fx.update(sd_np(m, f"{tag}::w::"))
        print("ee_attention_probe", tag, torch.stack(cls).numpy().round(3).tolist())
    np.savez_compressed(OUT / "ee_attention_probe_tiny.npz", **fx)


if __name__ == "__main__":
    torch.set_num_threads(4)
    if len(sys.argv) > 2 and sys.argv[2] == "variants":
        uvit_variants_fixture()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "attention_probe":
        ee_attention_probe_fixture()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "probe_types":  # add the probe-layout fixture without touching the others
        ee_probe_types_fixture()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "ae":  # add the autoencoder fixture without touching the others
        ae_decode_fixture()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "ddim":  # add the DDIM fixture without touching the others
        ddim_sampler_fixture()
        sys.exit(0)
    schedule_fixture()
    uvit_forward_fixture("uvit_forward_tiny", TINY, 1, False)
    uvit_forward_fixture("uvit_forward_tiny_cls", TINY_CLS, 3, True)
    ee_model = ee_forward_fixture()
    duodiff_sampler_fixture()
    ddim_sampler_fixture()
    ee_sampler_fixture(ee_model)
    ae_decode_fixture()
    ee_probe_types_fixture()
    ee_attention_probe_fixture()
    uvit_variants_fixture()
