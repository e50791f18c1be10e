"""Pins rows a4 / a5 (speechbrain Fbank + InputNormalization + ECAPA_TDNN) at the source, WHERE speechbrain is
importable — it is not in this container, so nothing here runs during the build; the script is the
committed recipe a maintainer runs once on a machine that has the reference's own dependency:

    pip install speechbrain            # the reference pins no version; >= 1.0 for speechbrain.inference
    python tests/golden/make_speechbrain_golden.py [--source speechbrain/spkrec-ecapa-voxceleb] [--out tests/golden]

It calls exactly what the reference calls (speech_encode.py:64-78, diar_diag.py:136,169):
``EncoderClassifier.from_hparams(source=...)`` and ``encode_batch(wavs)``, on the seeded synthetic audio of
tests/conftest.py::synth_wave, and stores

    speechbrain_ref.npz      wav seeds / shapes, the features after compute_features + mean_var_norm ([B, T, 80]),
                             the embeddings ([B, 192]), speechbrain.__version__ and the checkpoint source
    speechbrain_ecapa.ckpt   the embedding model's state dict (torch.save; 83 MB — NOT committed: point
                             $SD_ECAPA_CKPT at it, or at the hub's own embedding_model.ckpt, when running the test)

tests/test_gpu_speechbrain_pin.py consumes both and is skipped while either is absent.  With them present it
checks the CUDA front end (|d dB| <= 3e-3 on the normalised features) and the embeddings (1 - cos <= 1e-4) against
speechbrain itself, and the oracle restatement (oracle/ecapa_oracle.py) against the same fixtures — at which point
the "parity unpinned" caveat of rows a4 / a5 / (c) in DESIGN.md can be dropped.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="speechbrain/spkrec-ecapa-voxceleb")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    try:
        import speechbrain
        from speechbrain.inference.classifiers import EncoderClassifier
    except ImportError as e:
        sys.exit(f"speechbrain is not importable here ({e}); run this where the reference's dependency is installed")
    from conftest import synth_wave

    enc = EncoderClassifier.from_hparams(source=args.source, run_opts={"device": "cpu"})
    enc.eval()
    out = {"speechbrain_version": np.array(speechbrain.__version__), "source": np.array(args.source)}
    cases = {"win15": (4, 24000, 101), "win10": (3, 16000, 102), "short": (2, 4000, 103), "long": (1, 160000, 104)}
    with torch.inference_mode():
        for tag, (B, n, seed) in cases.items():
            w = torch.from_numpy(synth_wave(B, n, seed))
            lens = torch.ones(B)
            feats = enc.mods.compute_features(w)
            feats = enc.mods.mean_var_norm(feats, lens)
            emb = enc.encode_batch(w).squeeze(1)
            out[f"{tag}_shape"] = np.array([B, n, seed])
            out[f"{tag}_feats"] = feats.numpy().astype(np.float32)
            out[f"{tag}_emb"] = emb.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(args.out, "speechbrain_ref.npz"), **out)
    torch.save({k: v.detach().cpu() for k, v in enc.mods.embedding_model.state_dict().items()},
               os.path.join(args.out, "speechbrain_ecapa.ckpt"))
    print("wrote speechbrain_ref.npz and speechbrain_ecapa.ckpt to", args.out)


if __name__ == "__main__":
    main()
