"""Generates tests/golden/ecapa_hf_ref.npz: a third-party pin of the ECAPA-TDNN trunk topology.

speechbrain (whose ECAPA_TDNN the reference runs through EncoderClassifier.encode_batch) is not installed here,
but transformers is, and its ``ECAPA_TimeDelayNet`` — the speaker encoder inside Qwen2.5-Omni's token2wav — is a
port of speechbrain's ``lobes/models/ECAPA_TDNN.py`` with the BatchNorm layers removed (same TDNN / Res2Net / SE /
attentive-statistics-pooling / fc wiring, reflect 'same' padding, eps 1e-12 in the pooled std).  Loaded with the
oracle's seed-0 convolution weights it must agree with oracle/ecapa_oracle.py once the oracle's BatchNorms are
exact identities: that pins everything about the oracle's trunk except where the BatchNorms sit.

Run:  python tests/golden/make_ecapa_hf_golden.py     (needs transformers; not run on the GPU box)
"""
import os
import re
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


def neutral_bn_ecapa(seed=0):
    """oracle ECAPA-TDNN with random conv weights and every BatchNorm an exact identity."""
    from oracle import ecapa_oracle
    model = ecapa_oracle.make_random_ecapa(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.fill_(1); m.bias.data.zero_(); m.running_mean.zero_(); m.running_var.fill_(1.0 - m.eps)
    return model


def hf_ecapa_from(model):
    """transformers' ECAPA_TimeDelayNet (the speechbrain-derived speaker encoder inside Qwen2.5-Omni's token2wav;
    same topology as speechbrain's ECAPA_TDNN minus the BatchNorm layers) loaded with `model`'s conv weights."""
    from transformers.models.qwen2_5_omni.configuration_qwen2_5_omni import Qwen2_5OmniDiTConfig
    from transformers.models.qwen2_5_omni.modeling_qwen2_5_omni import ECAPA_TimeDelayNet
    cfg = Qwen2_5OmniDiTConfig(mel_dim=80, enc_dim=192, enc_channels=[1024, 1024, 1024, 1024, 3072],
                               enc_kernel_sizes=[5, 3, 3, 3, 1], enc_dilations=[1, 2, 3, 4, 1],
                               enc_attention_channels=128, enc_res2net_scale=8, enc_se_channels=128)
    hf = ECAPA_TimeDelayNet(cfg).eval()
    sd, new = model.state_dict(), {}
    for k, v in hf.state_dict().items():
        cand = k.replace(".conv.weight", ".conv.conv.weight").replace(".conv.bias", ".conv.conv.bias")
        for pat in (r"(se_block\.conv[12])\.(weight|bias)", r"(asp\.conv)\.(weight|bias)", r"^(fc)\.(weight|bias)"):
            cand = re.sub(pat, r"\1.conv.\2", cand)
        assert sd[cand].shape == v.shape, (k, cand)
        new[k] = sd[cand]
    hf.load_state_dict(new)
    return hf


def make_ecapa_hf_golden():
    """Third-party pin of the ECAPA-TDNN trunk topology: outputs of transformers' ECAPA_TimeDelayNet on seeded
    features with the oracle's seed-0 conv weights (BatchNorm = identity on the oracle side)."""
    hf = hf_ecapa_from(neutral_bn_ecapa(0))
    g = torch.Generator().manual_seed(7)
    feats = torch.randn(3, 61, 80, generator=g)
    with torch.inference_mode():
        out = hf(feats)
    import transformers
    np.savez_compressed(os.path.join(OUT, "ecapa_hf_ref.npz"), feats=feats.numpy(), emb=out.numpy(),
                        transformers_version=np.array(transformers.__version__))



if __name__ == "__main__":
    make_ecapa_hf_golden()
    p = os.path.join(OUT, "ecapa_hf_ref.npz")
    print("ecapa_hf_ref.npz", os.path.getsize(p))
