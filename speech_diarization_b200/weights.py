"""Synthetic ECAPA-TDNN (C=1024) weights with speechbrain's state-dict keys and shapes, for
benchmarks and demos when no checkpoint is available (there is no network here; the reference
fetches LanceaKing/spkrec-ecapa-cnceleb from the HF hub, speech_encode.py:66-69).

He-scaled convolution weights and mild BatchNorm statistics keep activations O(1) through the
trunk, as in a trained network.  The parity tests do NOT use this: they take the oracle's
random-init model (oracle/ecapa_oracle.py) so that both sides share identical weights."""
from __future__ import annotations

import math

import torch


def _tdnn(sd: dict, prefix: str, cout: int, cin: int, k: int, g: torch.Generator, in_rms: float = 1.0) -> None:
    # pre-activation ~ N(0, 2) for inputs of RMS `in_rms`; ReLU of that has mean 0.564, variance 0.68
    sd[f"{prefix}.conv.conv.weight"] = torch.randn(cout, cin, k, generator=g) * math.sqrt(2.0 / (cin * k)) / in_rms
    sd[f"{prefix}.conv.conv.bias"] = 0.1 * torch.randn(cout, generator=g)
    sd[f"{prefix}.norm.norm.weight"] = 0.75 + 0.5 * torch.rand(cout, generator=g)
    sd[f"{prefix}.norm.norm.bias"] = 0.2 * torch.randn(cout, generator=g)
    sd[f"{prefix}.norm.norm.running_mean"] = 0.564 * (1.0 + 0.1 * torch.randn(cout, generator=g))
    sd[f"{prefix}.norm.norm.running_var"] = 0.68 * torch.exp(0.1 * torch.randn(cout, generator=g))


def random_ecapa_state_dict(seed: int = 0, C: int = 1024, n_mels: int = 80, att: int = 128,
                            se: int = 128, emb: int = 192) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd: dict = {}
    _tdnn(sd, "blocks.0", C, n_mels, 5, g, in_rms=8.0)          # dB-scale features are O(10)
    for b in (1, 2, 3):
        _tdnn(sd, f"blocks.{b}.tdnn1", C, C, 1, g, in_rms=1.1 + 0.35 * (b - 1))
        for i in range(7):
            _tdnn(sd, f"blocks.{b}.res2net_block.blocks.{i}", C // 8, C // 8, 3, g, in_rms=1.1 if i == 0 else 1.6)
        _tdnn(sd, f"blocks.{b}.tdnn2", C, C, 1, g, in_rms=1.1)
        sd[f"blocks.{b}.se_block.conv1.conv.weight"] = torch.randn(se, C, 1, generator=g) / math.sqrt(C)
        sd[f"blocks.{b}.se_block.conv1.conv.bias"] = 0.1 * torch.randn(se, generator=g)
        sd[f"blocks.{b}.se_block.conv2.conv.weight"] = torch.randn(C, se, 1, generator=g) / math.sqrt(se)
        sd[f"blocks.{b}.se_block.conv2.conv.bias"] = 0.1 * torch.randn(C, generator=g)
    _tdnn(sd, "mfa", 3 * C, 3 * C, 1, g, in_rms=1.5)
    _tdnn(sd, "asp.tdnn", att, 9 * C, 1, g, in_rms=1.0)
    sd["asp.conv.conv.weight"] = torch.randn(3 * C, att, 1, generator=g) / math.sqrt(att)
    sd["asp.conv.conv.bias"] = 0.1 * torch.randn(3 * C, generator=g)
    sd["asp_bn.norm.weight"] = 0.75 + 0.5 * torch.rand(6 * C, generator=g)
    sd["asp_bn.norm.bias"] = 0.2 * torch.randn(6 * C, generator=g)
    sd["asp_bn.norm.running_mean"] = 0.3 * torch.randn(6 * C, generator=g)
    sd["asp_bn.norm.running_var"] = 0.5 + 0.5 * torch.rand(6 * C, generator=g)
    sd["fc.conv.weight"] = torch.randn(emb, 6 * C, 1, generator=g) / math.sqrt(6 * C)
    sd["fc.conv.bias"] = 0.1 * torch.randn(emb, generator=g)
    return sd
