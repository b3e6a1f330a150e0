"""Deterministic random-init weights with the reference's state_dict keys and shapes.

The released checkpoints are not available offline (reference download.sh:1-30), so parity runs use
random-init weights (BASELINE.json configs[0]). These generators are reproducible on any machine
with the same torch build (CPU generator), load into the reference `Generator` / `CVAE` with
strict=True (checked by oracle/gen_golden.py) and into this package's drop-in modules.

Scales follow the PyTorch default initialisers of the reference layers (kaiming-uniform with
a=sqrt(5) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear/Conv weights and biases, xavier-uniform
for MultiheadAttention in_proj, N(0,1) for tokens / positional embedding).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import skeleton

DEFAULT_MODEL_CFG = {
    # configs/config.yaml:13-43
    "mot_in_dim": 15, "nframes": 60, "njoints": 24, "nbody": 6, "temporal_patch_size": 4,
    "encoder_dim": 256, "encoder_depth": 2, "encoder_heads": 4, "encoder_dim_head": 128, "encoder_mlp_dim": 512,
    "decoder_dim": 256, "decoder_depth": 2, "decoder_heads": 4, "decoder_dim_head": 256, "decoder_mlp_dim": 512,
    "prj_dim": 1024, "num_patches": -1, "num_classes": 6,
    "graph": {"joint": {"layout": "mocha", "strategy": "distance", "max_hop": 2},
              "bodypart": {"layout": "mocha", "strategy": "distance", "max_hop": 1}},
}


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(sd, gen, name, out_f, in_f, bias=True, conv_shape=None):
    bound = 1.0 / math.sqrt(in_f if conv_shape is None else in_f * conv_shape[0] * conv_shape[1])
    shape = (out_f, in_f) if conv_shape is None else (out_f, in_f, *conv_shape)
    sd[name + ".weight"] = _uniform(gen, shape, bound)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (out_f,), bound)


def generator_state_dict(seed: int = 1777, cfg: dict | None = None) -> "OrderedDict[str, torch.Tensor]":
    """State dict of reference `Generator` (model.py:15-106): 65 parameters + 6 buffers."""
    cfg = cfg or DEFAULT_MODEL_CFG
    gen = torch.Generator(device="cpu").manual_seed(seed)
    D = cfg["encoder_dim"]
    tp = cfg["temporal_patch_size"]
    C0 = D // tp
    Cin = cfg["mot_in_dim"]
    ntok = cfg["nbody"] * (cfg["nframes"] // tp)
    A_j = torch.from_numpy(skeleton.joint_adjacency(cfg["graph"]["joint"]["max_hop"]))
    A_b = torch.from_numpy(skeleton.body_adjacency(cfg["graph"]["bodypart"]["max_hop"]))
    Kj, Kb = A_j.shape[0], A_b.shape[0]
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["pos_emb"] = torch.randn((1, ntok, D), generator=gen)
    _linear(sd, gen, "mot_embedding.1", C0, Cin, conv_shape=(1, 1))
    sd["mot_embedding.2.A_j"] = A_j.clone()
    _linear(sd, gen, "mot_embedding.2.blk.gcn.conv", D * Kj, C0, conv_shape=(1, 1))
    _linear(sd, gen, "mot_embedding.2.blk.tcn", D, D, conv_shape=(5, 1))
    sd["mot_embedding.3.weight"] = torch.from_numpy(skeleton.pool_weight())
    sd["mot_embedding.5.A_b"] = A_b.clone()
    _linear(sd, gen, "mot_embedding.5.blk.gcn.conv", D * Kb, D, conv_shape=(1, 1))
    _linear(sd, gen, "mot_embedding.5.blk.tcn", D, D, conv_shape=(3, 1))
    for l in range(cfg["encoder_depth"]):
        inner = cfg["encoder_heads"] * cfg["encoder_dim_head"]
        p = f"encoder.layers.{l}"
        _linear(sd, gen, p + ".1.to_q.1", inner, D, bias=False)
        _linear(sd, gen, p + ".1.to_k.1", inner, D, bias=False)
        _linear(sd, gen, p + ".1.to_v", inner, D, bias=False)
        _linear(sd, gen, p + ".1.to_out.0", D, inner)
        _linear(sd, gen, p + ".2.net.0", cfg["encoder_mlp_dim"], D)
        _linear(sd, gen, p + ".2.net.3", D, cfg["encoder_mlp_dim"])
    Dd = cfg["decoder_dim"]
    for l in range(cfg["decoder_depth"]):
        inner = cfg["decoder_heads"] * cfg["decoder_dim_head"]
        p = f"decoder.layers.{l}"
        _linear(sd, gen, p + ".0.style.2", 2 * Dd, Dd)
        _linear(sd, gen, p + ".0.style.4", 2 * Dd, 2 * Dd)
        _linear(sd, gen, p + ".1.to_q.1", inner, Dd, bias=False)
        _linear(sd, gen, p + ".1.to_k.1", inner, Dd, bias=False)
        _linear(sd, gen, p + ".1.to_v", inner, Dd, bias=False)
        _linear(sd, gen, p + ".1.to_out.0", Dd, inner)
        _linear(sd, gen, p + ".2.net.0", cfg["decoder_mlp_dim"], Dd)
        _linear(sd, gen, p + ".2.net.3", Dd, cfg["decoder_mlp_dim"])
    sd["to_mot.1.A_b"] = A_b.clone()
    _linear(sd, gen, "to_mot.1.blk.gcn.conv", Dd * Kb, Dd, conv_shape=(1, 1))
    _linear(sd, gen, "to_mot.1.blk.tcn", Dd, Dd, conv_shape=(3, 1))
    sd["to_mot.3.weight"] = torch.from_numpy(skeleton.unpool_weight())
    sd["to_mot.4.A_j"] = A_j.clone()
    _linear(sd, gen, "to_mot.4.blk.gcn.conv", (Dd // tp) * Kj, Dd, conv_shape=(1, 1))
    _linear(sd, gen, "to_mot.4.blk.tcn", Dd // tp, Dd // tp, conv_shape=(5, 1))
    _linear(sd, gen, "to_mot.6", Cin, Dd // tp, conv_shape=(1, 1))
    return sd


def sinusoid_table(max_len: int = 5000, d_model: int = 256) -> torch.Tensor:
    """PositionalEncoding buffer `pe` (model_CVAE.py:173-178), shape [1, max_len, d_model]."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-np.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def _mha(sd, gen, p, D):
    bound = math.sqrt(6.0 / (3 * D + D))  # xavier_uniform_ on [3D, D]
    sd[p + ".in_proj_weight"] = _uniform(gen, (3 * D, D), bound)
    # PyTorch zero-inits these biases; small random values exercise the bias paths in parity tests
    sd[p + ".in_proj_bias"] = _uniform(gen, (3 * D,), 0.02)
    sd[p + ".out_proj.weight"] = _uniform(gen, (D, D), 1.0 / math.sqrt(D))
    sd[p + ".out_proj.bias"] = _uniform(gen, (D,), 0.02)


def _ln(sd, gen, p, D):
    sd[p + ".weight"] = 1.0 + _uniform(gen, (D,), 0.1)
    sd[p + ".bias"] = _uniform(gen, (D,), 0.1)


def cvae_state_dict(seed: int = 1778, latent_dim: int = 256, depth: int = 2, dff: int = 512
                    ) -> "OrderedDict[str, torch.Tensor]":
    """State dict of reference `CVAE` (model_CVAE.py:8-46): 88 parameters + 3 `pe` buffers."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    D = latent_dim
    pe = sinusoid_table(5000, D)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def enc_stack(prefix):
        sd[prefix + ".mu_token"] = torch.randn((1, 1, D), generator=gen)
        sd[prefix + ".logvar_token"] = torch.randn((1, 1, D), generator=gen)
        sd[prefix + ".pos_encoder.pe"] = pe.clone()
        for l in range(depth):
            p = f"{prefix}.encoder.layers.{l}"
            _mha(sd, gen, p + ".self_attn", D)
            _linear(sd, gen, p + ".linear1", dff, D)
            _linear(sd, gen, p + ".linear2", D, dff)
            _ln(sd, gen, p + ".norm1", D)
            _ln(sd, gen, p + ".norm2", D)

    enc_stack("prior_net")
    enc_stack("encoder")
    sd["decoder.pos_encoder.pe"] = pe.clone()
    for l in range(depth):
        p = f"decoder.decoder.layers.{l}"
        _mha(sd, gen, p + ".self_attn", D)
        _mha(sd, gen, p + ".multihead_attn", D)
        _linear(sd, gen, p + ".linear1", dff, D)
        _linear(sd, gen, p + ".linear2", D, dff)
        _ln(sd, gen, p + ".norm1", D)
        _ln(sd, gen, p + ".norm2", D)
        _ln(sd, gen, p + ".norm3", D)
    return sd
