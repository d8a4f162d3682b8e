"""Drop-in for the reference's ``models/early_exit.py`` interface: ``EarlyExitUViT(uvit, classifier_type, exit_threshold)``
with the checkpoint layout of SURVEY.md Q16 (``uvit.*``, ``matrix.{i}.classifier.0.*``, ``in_blocks_heads.*``,
``mid_block_head.*``, ``out_blocks_heads.*``) and ``forward -> (eps, [probe_i], [head_output_i])``
(models/early_exit.py:268-320).  All four classifier types of the reference are supported -- ``mlp_probe_per_layer``
(the type of every configs/deediff_*.yaml), ``mlp_probe_per_timestep``, ``mlp_probe_per_layer_per_timestep``
(models/early_exit.py:194-239: ``matrix["i"]``, ``matrix["t"]``, ``matrix["i, t"]``) and ``attention_probe`` (the
constructor's default; one ``AttentionProbe`` per layer, whose scores are not sigmoids).
"""
from __future__ import annotations

import torch
from torch import nn

from .engine import Engine
from .uvit import UViT


class OutputHead(nn.Module):  # parameter names of models/early_exit.py:9-20
    def __init__(self, embed_dim: int, patch_dim: int, in_chans: int, conv: bool = True):
        super().__init__()
        self.in_chans = in_chans
        self.norm = nn.LayerNorm(embed_dim)
        self.decoder_pred = nn.Linear(embed_dim, patch_dim, bias=True)
        self.final_layer = nn.Conv2d(in_chans, in_chans, 3, padding=1) if conv else nn.Identity()


class MLPProbe(nn.Module):  # models/early_exit.py:31-34
    def __init__(self, embed_dim: int):
        super().__init__()
        self.classifier = nn.Sequential(nn.Linear(embed_dim, 1), nn.Sigmoid())


class AttentionProbe(nn.Module):  # parameter names / shapes of models/early_exit.py:46-60
    def __init__(self, embed_dim: int, num_heads: int = 1):
        super().__init__()
        if num_heads != 1:
            raise NotImplementedError("AttentionProbe: the reference only ever builds num_heads = 1")
        self.num_heads = num_heads
        self.q = nn.Parameter(torch.zeros(1, num_heads, 1, embed_dim // num_heads))
        self.weight_kv = nn.Linear(embed_dim, 2 * embed_dim)
        self.classification = nn.Sequential(nn.Linear(embed_dim, embed_dim), nn.SiLU(), nn.Linear(embed_dim, 1))


# classifier_type -> ddb_uvit_config.early_exit (include/duodiff_b200.h)
PROBE_KINDS = {"mlp_probe_per_layer": 1, "mlp_probe_per_timestep": 2, "mlp_probe_per_layer_per_timestep": 3,
               "attention_probe": 4}


def probe_keys(classifier_type: str, depth: int) -> list[str]:
    """Keys of ``EarlyExitUViT.matrix`` in the reference's construction order (models/early_exit.py:217-239)."""
    if classifier_type in ("mlp_probe_per_layer", "attention_probe"):
        return [f"{i}" for i in range(depth)]
    if classifier_type == "mlp_probe_per_timestep":
        return [f"{t}" for t in range(1000)]
    if classifier_type == "mlp_probe_per_layer_per_timestep":
        return [f"{i}, {t}" for t in range(1000) for i in range(depth)]
    # (the reference constructs the module without probes and fails at the first forward, early_exit.py:203-204)
    raise ValueError(f"Unknown classifier type: {classifier_type}")


class EarlyExitUViT(nn.Module):
    def __init__(self, uvit: UViT, classifier_type="attention_probe", exit_threshold=0.2):
        super().__init__()
        keys = probe_keys(classifier_type, uvit.depth)
        self.uvit = uvit
        self.exit_threshold = exit_threshold
        self.classifier_type = classifier_type
        d, half = uvit.embed_dim, uvit.depth // 2
        probe = AttentionProbe if classifier_type == "attention_probe" else MLPProbe
        self.matrix = nn.ModuleDict({k: probe(d) for k in keys})
        head = lambda: OutputHead(d, uvit.patch_dim, uvit.in_chans)  # noqa: E731
        self.in_blocks_heads = nn.ModuleList([head() for _ in range(half)])
        self.mid_block_head = head()
        self.out_blocks_heads = nn.ModuleList([head() for _ in range(half)])
        self._engine: Engine | None = None
        self._engine_key = None

    @property
    def device(self):
        return next(self.parameters()).device

    def engine(self, batch: int) -> Engine:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or self._engine_key != key or self._engine.max_batch < batch:
            self._engine = None
            cap = max(batch, self.uvit.max_batch or 0)
            self._engine = Engine(self.state_dict(), max_batch=cap, **self.uvit.engine_kwargs(early_exit=PROBE_KINDS[self.classifier_type]))
            self._engine_key = key
        return self._engine

    def forward(self, x, timesteps, y=None):
        """models/early_exit.py:268-320: every probe and every head is evaluated (simulate mode)."""
        with torch.no_grad():
            use_y = y if self.uvit.label_emb is not None else None
            _, _, scores, outputs = self.engine(x.shape[0]).ee_forward(x, timesteps, use_y, threshold=0.0, mode=0)
        depth = self.uvit.depth
        return outputs[depth], [scores[i] for i in range(depth)], [outputs[i] for i in range(depth)]
