"""Pins oracle/ (the CPU restatement) to the reference: fixtures in tests/golden/*.npz were produced by the
unmodified reference modules (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import uvit_oracle as O
from tests.helpers import load_fixture, split_fixture

torch.set_num_threads(4)
ATOL = 2e-5  # fp32, different but equivalent op order (explicit softmax vs SDPA, reshape vs einops/conv)


def test_schedule_bit_exact():
    fx = load_fixture("schedule")
    sch = O.ddpm_schedule()
    for k in ("betas", "alphas", "alphas_bar", "alphas_bar_previous", "betas_tilde"):
        assert np.array_equal(sch[k].numpy(), fx[k]), k


def _check_forward(name, with_y):
    fx = load_fixture(name)
    sd, params = split_fixture(fx)
    spec = O.UViTSpec.from_params(params)
    y = torch.from_numpy(fx["y"]) if with_y else None
    cap = []
    out = O.uvit_forward(sd, spec, torch.from_numpy(fx["x"]), torch.from_numpy(fx["t"]), y, capture=cap)
    for i in range(spec.depth):
        np.testing.assert_allclose(cap[i + 1].numpy(), fx[f"hidden_{i}"], atol=ATOL * 10, rtol=1e-5)
    np.testing.assert_allclose(out.numpy(), fx["out"], atol=ATOL, rtol=1e-5)


def test_uvit_forward_unconditional():
    _check_forward("uvit_forward_tiny", False)


def test_uvit_forward_class_conditional_raw_timesteps():
    _check_forward("uvit_forward_tiny_cls", True)


def test_ee_forward_probes_and_heads():
    fx = load_fixture("ee_forward_tiny")
    sd, params = split_fixture(fx)
    spec = O.UViTSpec.from_params(params)
    eps, cls, outs = O.ee_forward(sd, spec, torch.from_numpy(fx["x"]), torch.from_numpy(fx["t"]))
    assert len(cls) == len(outs) == spec.depth  # tests/models/test_early_exit.py:98-115 of the reference
    np.testing.assert_allclose(eps.numpy(), fx["eps"], atol=ATOL, rtol=1e-5)
    np.testing.assert_allclose(torch.stack(cls).numpy(), fx["cls"], atol=1e-6)
    np.testing.assert_allclose(torch.stack(outs).numpy(), fx["outs"], atol=ATOL, rtol=1e-5)


@pytest.mark.parametrize("tag,ctype", [("pt", "mlp_probe_per_timestep"), ("plt", "mlp_probe_per_layer_per_timestep")])
def test_ee_forward_timestep_indexed_probes(tag, ctype):
    """models/early_exit.py:194-204: matrix["t"] / matrix["i, t"] with t = int(timesteps[0]) -- also when the batch
    mixes timesteps (third call of the fixture)."""
    fx = load_fixture("ee_probe_types_tiny")
    sd, params = split_fixture(fx, f"{tag}::w::", "p::")
    spec = O.UViTSpec.from_params(params)
    x = torch.from_numpy(fx[f"{tag}::x"])
    for c in range(3):
        eps, cls, outs = O.ee_forward(sd, spec, x, torch.from_numpy(fx[f"{tag}::t{c}"]), classifier_type=ctype)
        np.testing.assert_allclose(eps.numpy(), fx[f"{tag}::eps{c}"], atol=ATOL, rtol=1e-5)
        np.testing.assert_allclose(torch.stack(cls).numpy(), fx[f"{tag}::cls{c}"], atol=1e-6)
        np.testing.assert_allclose(torch.stack(outs).numpy(), fx[f"{tag}::outs{c}"], atol=ATOL, rtol=1e-5)
    # the probes depend on the timestep: the scores of calls 0 and 1 must differ
    assert np.abs(fx[f"{tag}::cls0"] - fx[f"{tag}::cls1"]).max() > 1e-2


@pytest.mark.parametrize("tag", ["a", "b"])
def test_uvit_forward_constructor_variants(tag):
    """a: mlp_time_embed=True (fractional timestep in the batch); b: conv=False + skip=False (+ qk_scale, ignored by the
    reference), class-conditional (models/uvit.py:229-247, 264-272, 200-204, 329-333)."""
    fx = load_fixture("uvit_variants_tiny")
    sd, params = split_fixture(fx, f"{tag}::w::", f"{tag}::p::")
    spec = O.UViTSpec.from_params(params)
    y = torch.from_numpy(fx[f"{tag}::y"]) if f"{tag}::y" in fx else None
    out = O.uvit_forward(sd, spec, torch.from_numpy(fx[f"{tag}::x"]), torch.from_numpy(fx[f"{tag}::t"]), y)
    np.testing.assert_allclose(out.numpy(), fx[f"{tag}::out"], atol=ATOL, rtol=1e-5)


@pytest.mark.parametrize("tag", ["u", "c"])
def test_ee_forward_attention_probe(tag):
    """models/early_exit.py:40-80 through EarlyExitUViT(..., "attention_probe"): unconditional and class-conditional
    (where x[:, 1:] drops the label token and keeps the time token)."""
    fx = load_fixture("ee_attention_probe_tiny")
    sd, params = split_fixture(fx, f"{tag}::w::", f"{tag}::p::")
    spec = O.UViTSpec.from_params(params)
    y = torch.from_numpy(fx[f"{tag}::y"]) if f"{tag}::y" in fx else None
    eps, cls, outs = O.ee_forward(sd, spec, torch.from_numpy(fx[f"{tag}::x"]), torch.from_numpy(fx[f"{tag}::t"]), y,
                                  classifier_type="attention_probe")
    np.testing.assert_allclose(eps.numpy(), fx[f"{tag}::eps"], atol=ATOL, rtol=1e-5)
    np.testing.assert_allclose(torch.stack(cls).numpy(), fx[f"{tag}::cls"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(torch.stack(outs).numpy(), fx[f"{tag}::outs"], atol=ATOL, rtol=1e-5)


def _replay_reference_rng(seed, shape):
    """seed_everything + x_T draw of sampler.py:99-100; z_t drawn lazily from the same global CPU stream."""
    torch.manual_seed(seed)
    x_T = torch.randn(*shape)
    return x_T, (lambda t: torch.randn(*shape))


def test_duodiff_sampler_all_rules():
    fx = load_fixture("duodiff_sampler_tiny")
    sde, pe = split_fixture(fx, "we::", "p::")
    sdl, pl = split_fixture(fx, "wl::", "q::")
    se, sl = O.UViTSpec.from_params(pe), O.UViTSpec.from_params(pl)
    assert (fx["early_calls"], fx["early_t_min"], fx["late_calls"], fx["late_t_max"]) == (300, 700, 700, 699)  # Q1
    early = lambda x, t, y: O.uvit_forward(sde, se, x, t, y)  # noqa: E731
    late = lambda x, t, y: O.uvit_forward(sdl, sl, x, t, y)  # noqa: E731
    for rule in ("predict_noise", "predict_original", "predict_previous"):
        x_T, noise = _replay_reference_rng(5, (2, 3, 8, 8))
        x0 = O.sample_ddpm(early, late, 300, x_T, noise, rule=rule)
        got = O.to_samples_nhwc(x0).numpy()
        ref = fx[f"samples_{rule}"]
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 2e-4 * scale, rule  # 1000 chained fp32 steps


def test_ddim_sampler():
    """sampler.py:103-126 (use_ddim): strided schedule, hand-off rule `t < 1000 - t_switch` after the step, the
    sigma^2 * z quirk, intermediates."""
    fx = load_fixture("ddim_sampler_tiny")
    sde, pe = split_fixture(fx, "we::", "p::")
    sdl, pl = split_fixture(fx, "wl::", "q::")
    se, sl = O.UViTSpec.from_params(pe), O.UViTSpec.from_params(pl)
    early = lambda x, t, y: O.uvit_forward(sde, se, x, t, y)  # noqa: E731
    late = lambda x, t, y: O.uvit_forward(sdl, sl, x, t, y)  # noqa: E731
    for steps, eta in ((50, 0.0), (20, 0.05)):
        key = f"{steps}_{eta}"
        x_T, noise = _replay_reference_rng(7, (2, 3, 8, 8))
        trace = {}
        x0 = O.sample_ddim(early, late, 300, x_T, noise, steps, eta, trace=trace)
        # which backbone ran at which t
        assert [t for t, l in zip(trace["t"], trace["late"]) if not l] == fx[f"early_ts_{key}"].tolist()
        assert [t for t, l in zip(trace["t"], trace["late"]) if l] == fx[f"late_ts_{key}"].tolist()
        ref = fx[f"samples_{key}"]
        assert np.isfinite(ref).all()
        assert np.abs(O.to_samples_nhwc(x0).numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), key
        # timesteps_save=[1, 1000 - ts[3]] -> x after the pairs starting at t = 999 and t = ts[3], in loop order
        x_T, noise = _replay_reference_rng(7, (2, 3, 8, 8))
        x1 = O.sample_ddim(early, late, 300, x_T, noise, steps, eta, n_pairs=1)
        ref1 = fx[f"inter_{key}_0"]
        assert np.abs(O.to_samples_nhwc(x1).numpy() - ref1).max() <= 1e-5 * np.abs(ref1).max()


def test_duodiff_sampler_intermediates():
    fx = load_fixture("duodiff_sampler_tiny")
    sde, pe = split_fixture(fx, "we::", "p::")
    sdl, pl = split_fixture(fx, "wl::", "q::")
    se, sl = O.UViTSpec.from_params(pe), O.UViTSpec.from_params(pl)
    early = lambda x, t, y: O.uvit_forward(sde, se, x, t, y)  # noqa: E731
    late = lambda x, t, y: O.uvit_forward(sdl, sl, x, t, y)  # noqa: E731
    # timesteps_save=[1,300,990] -> saved after the steps t = 999, 700, 10, appended in loop order
    x_T, noise = _replay_reference_rng(5, (2, 3, 8, 8))
    x = O.sample_ddpm(early, late, 300, x_T, noise, t_first=999, t_last=999)
    ref = fx["inter_predict_noise_0"]
    assert np.abs(O.to_samples_nhwc(x).numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_ee_sampler_logs_and_samples():
    fx = load_fixture("ee_sampler_tiny")
    sd, params = split_fixture(fx)
    spec = O.UViTSpec.from_params(params)
    model = lambda x, t, y: O.ee_forward(sd, spec, x, t, y)  # noqa: E731
    x_T, noise = _replay_reference_rng(9, (3, 3, 8, 8))
    x0, err_log, idx_log = O.ee_sample(model, float(fx["threshold"]), spec.depth, x_T, noise)
    # exit indices must be identical except where a probe sits within 1e-5 of the threshold
    mism = (idx_log.numpy() != fx["idx_log"]).mean()
    assert mism <= 0.002, mism
    np.testing.assert_allclose(err_log.numpy(), fx["err_log"], atol=5e-4)
    ref = fx["samples"]
    assert np.abs(O.to_samples_nhwc(x0).numpy() - ref).max() <= 5e-3 * np.abs(ref).max()


def test_autoencoder_decode_matches_reference():
    """oracle/ae_oracle.py vs FrozenAutoencoderKL.decode of the reference (tests/golden/ae_decode_tiny.npz)."""
    from oracle import ae_oracle as A
    fx = load_fixture("ae_decode_tiny")
    sd, _ = split_fixture(fx)
    spec = A.AESpec(ch=int(fx["ch"]), ch_mult=[int(v) for v in fx["ch_mult"]],
                    num_res_blocks=int(fx["num_res_blocks"]), resolution=int(fx["resolution"]),
                    scale_factor=float(fx["scale_factor"]))
    tap = {}
    out = A.decode(sd, spec, torch.from_numpy(fx["z"]), tap)
    np.testing.assert_allclose(out.numpy(), fx["out"], atol=ATOL, rtol=1e-5)
    assert "mid.attn_1.proj_out+res" in tap and "up.1.upsample" in tap and "norm_out" in tap
    # the random factory used by the GPU parity tests produces exactly the reference decoder's keys and shapes
    rnd = A.random_state_dict(spec, 0)
    assert set(rnd) == set(sd)
    assert all(rnd[k].shape == sd[k].shape for k in sd)
