"""GPU tests of the two command-line entry points (sampler.py:192-352, eesampler.py:114-209): checkpoint file -> yaml
config -> 1000-step (or DDIM) sampling on the B200 kernels -> the reference's output files."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

from tests.helpers import CONFIGS, heat_

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _write(tmp: Path, name: str, params: dict, seed: int, wrap: bool, ee: bool = False, extra=None,
           ctype: str = "mlp_probe_per_layer"):
    """Random-init checkpoint (bare state_dict, or a training checkpoint carrying `model_state_dict`) + yaml config."""
    import duodiff_b200 as ddb
    torch.manual_seed(seed)
    net = ddb.UViT(**params)
    mp = dict(params)
    if ee:
        net = ddb.EarlyExitUViT(net, ctype)
        heat_(net, seed)
        with torch.no_grad():
            if ctype == "mlp_probe_per_layer":
                for i in range(params["depth"]):
                    net.matrix[f"{i}"].classifier[0].bias.fill_(1.5 - 3.0 * i / params["depth"])
            elif ctype == "attention_probe":
                for i in range(params["depth"]):
                    net.matrix[f"{i}"].q.normal_()
        mp["classifier_type"] = ctype
    if extra:
        mp.update(extra)
    sd = net.state_dict()
    ck, cfg = tmp / f"{name}.pth", tmp / f"{name}.yaml"
    torch.save({"model_state_dict": sd, "epoch": 3} if wrap else sd, ck)
    cfg.write_text(yaml.safe_dump({"model_params": mp, "dataset": "synthetic"}))
    return str(ck), str(cfg)


def _png(path: Path):
    from PIL import Image
    return np.asarray(Image.open(path))


def test_sampler_cli_duodiff_ddpm(tmp_path):
    """README.md:104-110: two checkpoints, --t_switch 300, intermediates via --timesteps_save."""
    from duodiff_b200 import sampler as S
    ck_e, cfg_e = _write(tmp_path, "early", CONFIGS["cifar10_3"], 1, wrap=False)
    ck_l, cfg_l = _write(tmp_path, "late", CONFIGS["cifar10"], 2, wrap=True)
    out = tmp_path / "out"
    S.main(["--checkpoint_path", ck_e, "--config_path", cfg_e, "--checkpoint_path_late", ck_l, "--config_path_late",
            cfg_l, "--t_switch", "300", "--batch_size", "3", "--parametrization", "predict_noise", "--output_folder",
            str(out), "--seed", "3", "--timesteps_save", "500", "900"])
    names = sorted(p.name for p in out.iterdir())
    want = sorted(["0.png", "1.png", "2.png", "grid_image.png", "statistics.txt"]
                  + [f"{i}_{t}.png" for i in range(3) for t in (500, 900)])
    assert names == want
    assert (out / "statistics.txt").read_text().startswith("Elapsed time: ")
    assert float((out / "statistics.txt").read_text().split()[2]) > 0
    img = _png(out / "0.png")
    assert img.shape == (32, 32, 4) and img.dtype == np.uint8 and (img[..., 3] == 255).all()  # plt.imsave: RGBA
    assert _png(out / "grid_image.png").shape == (64, 64, 4)  # ceil(sqrt(3)) = 2 tiles per side
    assert not np.array_equal(_png(out / "0_500.png"), _png(out / "0_900.png"))
    # same seed -> same files; the API gives the same samples as the CLI wrote
    import duodiff_b200 as ddb
    early = ddb.UViT(**CONFIGS["cifar10_3"])
    early.load_state_dict(torch.load(ck_e))
    late = ddb.UViT(**CONFIGS["cifar10"])
    late.load_state_dict(torch.load(ck_l)["model_state_dict"])
    samples, inter = S.get_samples(early.eval().cuda(), 3, S.predict_noise_postprocessing, seed=3, num_channels=3,
                                   sample_height=32, sample_width=32, late_model=late.eval().cuda(), t_switch=300,
                                   timesteps_save=[500, 900])
    assert len(inter) == 2
    ref_png = (np.clip(samples[0], 0, 1) * 255).astype(np.uint8)
    assert np.array_equal(img[..., :3], ref_png)


@pytest.mark.parametrize("rule", ["predict_original", "predict_previous"])
def test_sampler_cli_other_rules_single_model(tmp_path, rule):
    from duodiff_b200 import sampler as S
    ck, cfg = _write(tmp_path, "m", CONFIGS["cifar10_3"], 4, wrap=False)
    out = tmp_path / rule
    S.main(["--checkpoint_path", ck, "--config_path", cfg, "--batch_size", "2", "--parametrization", rule,
            "--output_folder", str(out)])
    assert sorted(p.name for p in out.iterdir()) == ["0.png", "1.png", "grid_image.png", "statistics.txt"]


def test_sampler_cli_ddim_class_conditional(tmp_path):
    """--use_ddim + --class_id on a class-conditional config whose yaml carries the stray `classifier_type` key of
    configs/uvit_imagenet64.yaml; checkpoint in the `model_state_dict` wrapper."""
    from duodiff_b200 import sampler as S
    ck, cfg = _write(tmp_path, "in64", CONFIGS["imagenet64_3"], 5, wrap=True, extra={"classifier_type": "attention_probe"})
    out = tmp_path / "ddim"
    S.main(["--checkpoint_path", ck, "--config_path", cfg, "--batch_size", "2", "--parametrization", "predict_noise",
            "--output_folder", str(out), "--use_ddim", "--ddim_steps", "20", "--ddim_eta", "0.02", "--class_id", "7",
            "--seed", "1"])
    assert sorted(p.name for p in out.iterdir()) == ["0.png", "1.png", "grid_image.png", "statistics.txt"]
    assert _png(out / "1.png").shape == (64, 64, 4)


def test_eesampler_cli(tmp_path):
    """README.md:120-124: --threshold, per-sample PNGs, statistics.txt and the two .pt logs (eesampler.py:102-111)."""
    from duodiff_b200 import eesampler as ES
    ck, cfg = _write(tmp_path, "ee", CONFIGS["cifar10"], 6, wrap=False, ee=True)
    out = tmp_path / "ee_out"
    ES.main(["--checkpoint_path", ck, "--config_path", cfg, "--threshold", "0.45", "--batch_size", "3",
             "--output_folder", str(out), "--seed", "2"])
    names = sorted(p.name for p in out.iterdir())
    assert names == sorted(["0.png", "1.png", "2.png", "statistics.txt", "error_prediction_by_timestep.pt",
                            "indices_by_timestep.pt"])
    err = torch.load(out / "error_prediction_by_timestep.pt")
    idx = torch.load(out / "indices_by_timestep.pt")
    assert err.shape == (1000, 13) and idx.shape == (1000, 3) and idx.dtype == torch.float32
    assert 0 <= float(idx.min()) and float(idx.max()) <= 13 and float(idx.min()) < 13  # exits happened
    assert torch.isfinite(err).all() and 0 < float(err.min()) and float(err.max()) < 1
    assert _png(out / "2.png").shape == (32, 32, 4)


@pytest.mark.parametrize("ctype,thr", [("attention_probe", "0.0"), ("mlp_probe_per_timestep", "0.5")])
def test_eesampler_cli_other_classifier_types(tmp_path, ctype, thr):
    """eesampler.py:157 hands config["model_params"]["classifier_type"] to EarlyExitUViT: the other probe layouts run
    through the same command line (checkpoint -> yaml -> 1000 steps -> PNGs + logs)."""
    from duodiff_b200 import eesampler as ES
    ck, cfg = _write(tmp_path, "ee", CONFIGS["cifar10_3"], 8, wrap=True, ee=True, ctype=ctype)
    out = tmp_path / "ee_out"
    ES.main(["--checkpoint_path", ck, "--config_path", cfg, "--threshold", thr, "--batch_size", "2",
             "--output_folder", str(out), "--seed", "3"])
    assert sorted(p.name for p in out.iterdir()) == sorted(
        ["0.png", "1.png", "statistics.txt", "error_prediction_by_timestep.pt", "indices_by_timestep.pt"])
    err = torch.load(out / "error_prediction_by_timestep.pt")
    idx = torch.load(out / "indices_by_timestep.pt")
    assert err.shape == (1000, 3) and idx.shape == (1000, 2) and torch.isfinite(err).all()
    assert 0 <= float(idx.min()) and float(idx.max()) <= 3
    assert _png(out / "1.png").shape == (32, 32, 4)


def test_top_level_scripts_run_like_the_reference(tmp_path):
    """`python sampler.py ...` and `python eesampler.py ...` from the repository root, in fresh processes."""
    ck, cfg = _write(tmp_path, "m", CONFIGS["cifar10_3"], 7, wrap=False)
    out = tmp_path / "cli"
    r = subprocess.run([sys.executable, "sampler.py", "--checkpoint_path", ck, "--config_path", cfg, "--batch_size", "2",
                        "--parametrization", "predict_noise", "--output_folder", str(out), "--use_ddim",
                        "--ddim_steps", "10"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Using device" in r.stdout
    assert sorted(p.name for p in out.iterdir()) == ["0.png", "1.png", "grid_image.png", "statistics.txt"]
    r = subprocess.run([sys.executable, "eesampler.py", "--help"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "--threshold" in r.stdout
