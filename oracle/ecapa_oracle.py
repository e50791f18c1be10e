"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 PyTorch restatement of the speaker-embedding front end the reference
reaches through ``EncoderClassifier.encode_batch`` (call sites
``speech_encode.py:77``, ``ecapa_annote.py:22``, ``diar_diag.py:169``):
speechbrain ``Fbank`` -> ``InputNormalization(sentence)`` -> ``ECAPA_TDNN``.

PARITY PARTLY PINNED for this file: speechbrain is a third-party dependency of the
reference that is absent from /root/reference and not installable here (no
version pin exists in the reference tree either; the import path
``speechbrain.inference`` implies >= 1.0), so the reference's own ECAPA cannot be
run.  What IS pinned: the trunk topology (TDNN / Res2Net / SE / attentive
statistics pooling / fc wiring, reflect 'same' padding, pooled-std epsilon) is
checked bit for bit against an installed third-party port of speechbrain's
``lobes/models/ECAPA_TDNN.py`` — transformers' ``ECAPA_TimeDelayNet`` (Qwen2.5-Omni
token2wav), which is that model minus the BatchNorm layers
(tests/golden/make_ecapa_hf_golden.py, tests/test_oracle_golden.py).  What is
still recalled from SURVEY.md Appendix A: where the BatchNorms sit (after the
ReLU inside every TDNN block; ``asp_bn`` before ``fc``) and the Fbank /
InputNormalization front end (``processing/features.py``).  Also cross-checked:
the parameter count (20 767 552 == the published embedding_model.ckpt size / 4)
and the state-dict key names, which mirror speechbrain's so that a real
checkpoint can validate the rest later.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

N_MELS = 80
EMB_DIM = 192


# --------------------------------------------------------------------------- fbank
def _to_mel(hz: float) -> float:
    return 2595.0 * math.log10(1.0 + hz / 700.0)


def speechbrain_filterbank_matrix(n_mels: int = 80, n_fft: int = 400, sr: int = 16000,
                                  f_min: float = 0.0, f_max: float = 8000.0) -> torch.Tensor:
    """speechbrain ``Filterbank`` triangular filters, [n_fft//2+1, n_mels] (App. A.1 step 3)."""
    mel = torch.linspace(_to_mel(f_min), _to_mel(f_max), n_mels + 2)
    hz = 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    band = (hz[1:] - hz[:-1])[:-1]
    f_central = hz[1:-1]
    all_freqs = torch.linspace(0, sr // 2, n_fft // 2 + 1)
    slope = (all_freqs[None, :] - f_central[:, None]) / band[:, None]
    fb = torch.clamp(torch.minimum(slope + 1.0, -slope + 1.0), min=0.0)
    return fb.transpose(0, 1).contiguous()


def fbank_speechbrain(wavs: torch.Tensor, mean_norm: bool = True) -> torch.Tensor:
    """wavs [B, n] f32 -> [B, T, 80]: STFT(Hamming 400/160, zero centre pad) -> power ->
    mel -> dB (top_db 80 per utterance) -> sentence mean normalisation (App. A.1)."""
    wavs = wavs.float()
    window = torch.hamming_window(400, device=wavs.device)
    spec = torch.stft(wavs, n_fft=400, hop_length=160, win_length=400, window=window, center=True,
                      pad_mode="constant", normalized=False, onesided=True, return_complex=True)
    power = spec.real.pow(2) + spec.imag.pow(2)          # [B, 201, T]
    power = power.transpose(1, 2)                        # [B, T, 201]
    fb = speechbrain_filterbank_matrix().to(wavs.device)
    mel = torch.matmul(power, fb)
    x_db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    floor = x_db.amax(dim=(-2, -1)) - 80.0
    x_db = torch.maximum(x_db, floor.view(-1, 1, 1))
    if mean_norm:
        x_db = x_db - x_db.mean(dim=1, keepdim=True)     # InputNormalization(sentence, std_norm=False)
    return x_db


# ---------------------------------------------------------------------- ECAPA-TDNN
class _Conv1d(nn.Module):
    """speechbrain.nnet.CNN.Conv1d with padding='same', padding_mode='reflect' (App. A.2)."""

    def __init__(self, in_channels, out_channels, kernel_size, dilation=1):
        super().__init__()
        self.kernel_size, self.dilation = kernel_size, dilation
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, dilation=dilation, padding=0)

    def forward(self, x):
        pad = self.dilation * (self.kernel_size - 1) // 2
        if pad:
            x = F.pad(x, (pad, pad), mode="reflect")
        return self.conv(x)


class _BatchNorm1d(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.norm = nn.BatchNorm1d(input_size, eps=1e-5, momentum=0.1)

    def forward(self, x):
        return self.norm(x)


class TDNNBlock(nn.Module):
    """conv -> ReLU -> BatchNorm (BN AFTER the activation)."""

    def __init__(self, in_channels, out_channels, kernel_size, dilation):
        super().__init__()
        self.conv = _Conv1d(in_channels, out_channels, kernel_size, dilation)
        self.activation = nn.ReLU()
        self.norm = _BatchNorm1d(out_channels)

    def forward(self, x):
        return self.norm(self.activation(self.conv(x)))


class Res2NetBlock(nn.Module):
    def __init__(self, channels, scale=8, kernel_size=3, dilation=1):
        super().__init__()
        c = channels // scale
        self.blocks = nn.ModuleList([TDNNBlock(c, c, kernel_size, dilation) for _ in range(scale - 1)])
        self.scale = scale

    def forward(self, x):
        y = []
        y_i = None
        for i, x_i in enumerate(torch.chunk(x, self.scale, dim=1)):
            if i == 0:
                y_i = x_i
            elif i == 1:
                y_i = self.blocks[i - 1](x_i)
            else:
                y_i = self.blocks[i - 1](x_i + y_i)
            y.append(y_i)
        return torch.cat(y, dim=1)


class SEBlock(nn.Module):
    def __init__(self, in_channels, se_channels, out_channels):
        super().__init__()
        self.conv1 = _Conv1d(in_channels, se_channels, 1)
        self.conv2 = _Conv1d(se_channels, out_channels, 1)

    def forward(self, x):
        s = x.mean(dim=2, keepdim=True)   # lengths are all ones in encode_batch
        s = torch.relu(self.conv1(s))
        s = torch.sigmoid(self.conv2(s))
        return s * x


class SERes2NetBlock(nn.Module):
    def __init__(self, channels, kernel_size, dilation, res2net_scale=8, se_channels=128):
        super().__init__()
        self.tdnn1 = TDNNBlock(channels, channels, 1, 1)
        self.res2net_block = Res2NetBlock(channels, res2net_scale, kernel_size, dilation)
        self.tdnn2 = TDNNBlock(channels, channels, 1, 1)
        self.se_block = SEBlock(channels, se_channels, channels)

    def forward(self, x):
        residual = x
        x = self.tdnn1(x)
        x = self.res2net_block(x)
        x = self.tdnn2(x)
        x = self.se_block(x)
        return x + residual


class AttentiveStatisticsPooling(nn.Module):
    def __init__(self, channels, attention_channels=128):
        super().__init__()
        self.eps = 1e-12
        self.tdnn = TDNNBlock(channels * 3, attention_channels, 1, 1)
        self.conv = _Conv1d(attention_channels, channels, 1)

    def _stats(self, x, m):
        mean = (m * x).sum(2)
        std = torch.sqrt((m * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(self.eps))
        return mean, std

    def forward(self, x):
        L = x.shape[-1]
        mean, std = self._stats(x, torch.full_like(x[:, :1, :], 1.0 / L))
        attn = torch.cat([x, mean.unsqueeze(2).expand(-1, -1, L), std.unsqueeze(2).expand(-1, -1, L)], dim=1)
        attn = self.conv(torch.tanh(self.tdnn(attn)))
        attn = F.softmax(attn, dim=2)
        mean, std = self._stats(x, attn)
        return torch.cat((mean, std), dim=1).unsqueeze(2)


class ECAPA_TDNN(nn.Module):
    """speechbrain ECAPA_TDNN(input_size=80, channels=[1024]*4+[3072], kernel_sizes=[5,3,3,3,1],
    dilations=[1,2,3,4,1], attention_channels=128, lin_neurons=192, res2net_scale=8,
    se_channels=128, global_context=True) — SURVEY App. A.3.  ``taps`` collects named
    intermediate activations ([B, C, T]) for the layer-wise parity tests."""

    def __init__(self, input_size=80, C=1024, lin_neurons=192, attention_channels=128):
        super().__init__()
        self.blocks = nn.ModuleList([TDNNBlock(input_size, C, 5, 1)])
        for d in (2, 3, 4):
            self.blocks.append(SERes2NetBlock(C, 3, d))
        self.mfa = TDNNBlock(3 * C, 3 * C, 1, 1)
        self.asp = AttentiveStatisticsPooling(3 * C, attention_channels)
        self.asp_bn = _BatchNorm1d(6 * C)
        self.fc = _Conv1d(6 * C, lin_neurons, 1)

    def forward(self, x, taps: dict | None = None):
        x = x.transpose(1, 2)
        xl = []
        for i, layer in enumerate(self.blocks):
            x = layer(x)
            xl.append(x)
            if taps is not None:
                taps["block0" if i == 0 else f"b{i}.out"] = x
        x = torch.cat(xl[1:], dim=1)
        x = self.mfa(x)
        if taps is not None:
            taps["mfa"] = x
        x = self.asp(x)
        if taps is not None:
            taps["pooled"] = x.squeeze(2)
        x = self.asp_bn(x)
        x = self.fc(x)
        return x.transpose(1, 2)      # [B, 1, 192]; no L2 norm


def synth_features(B: int, T: int, seed: int = 0) -> torch.Tensor:
    """Feature-like input (sentence-mean-normalised dB-scale log-mel) for calibration / tests."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, N_MELS, generator=g) * 8.0
    # smooth along time and mel so it resembles speech features rather than white noise
    x = F.avg_pool1d(x.transpose(1, 2), 5, 1, 2, count_include_pad=False).transpose(1, 2)
    return x - x.mean(dim=1, keepdim=True)


def make_random_ecapa(seed: int = 0, calib_T: int = 151, calib_B: int = 8) -> ECAPA_TDNN:
    """Random-init ECAPA-TDNN C=1024 (default nn.Conv1d init, seed `seed`) whose BatchNorm
    layers carry NON-TRIVIAL statistics: running mean/var are calibrated on synthetic
    features (what a trained network looks like — post-BN activations O(1)) and then
    perturbed, and the affine weight/bias are randomised, so BN bugs are visible
    (SURVEY §8c "randomised BN affine and running stats")."""
    torch.manual_seed(seed)
    m = ECAPA_TDNN()
    g = torch.Generator().manual_seed(seed + 1)
    bns = [mod for mod in m.modules() if isinstance(mod, nn.BatchNorm1d)]
    for bn in bns:
        bn.weight.data = 0.75 + 0.5 * torch.rand(bn.num_features, generator=g)
        bn.bias.data = 0.2 * torch.randn(bn.num_features, generator=g)
        bn.momentum = None            # cumulative average during calibration
    m.train()
    with torch.no_grad():
        m(synth_features(calib_B, calib_T, seed + 2))
    m.eval()
    for bn in bns:
        bn.momentum = 0.1
        bn.running_mean.data *= 1.0 + 0.1 * torch.randn(bn.num_features, generator=g)
        bn.running_var.data *= torch.exp(0.1 * torch.randn(bn.num_features, generator=g))
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())


@torch.inference_mode()
def encode_batch(model: ECAPA_TDNN, wavs: torch.Tensor) -> torch.Tensor:
    """EncoderClassifier.encode_batch(wavs) with wav_lens=None, normalize=False -> [B, 1, 192]."""
    feats = fbank_speechbrain(wavs.float(), mean_norm=True)
    return model(feats)
