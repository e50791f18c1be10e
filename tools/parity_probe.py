"""GPU bring-up probe: prints layer-by-layer parity of the CUDA path against the oracle.
(Development tool; the pytest -m gpu suite is the gate.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ecapa_oracle as eo, fbank_oracle as fo, cluster_oracle as co
from speech_diarization_b200 import speech_encode as se, clustering as cl

dev = torch.device("cuda:0")
what = sys.argv[1:] or ["fbank", "trunk", "e2e", "aff", "ahc"]

def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())

def synth_wave(B, n, seed):
    rng = np.random.default_rng(seed); t = np.arange(n) / 16000.0
    out = np.zeros((B, n), np.float32)
    for b in range(B):
        f0 = 100.0 + 17.0 * (b % 11)
        sig = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) / h for h in range(1, 15))
        out[b] = (0.08 * sig + 0.004 * rng.standard_normal(n)).astype(np.float32)
    return out

if "fbank" in what:
    for n in (24000, 4000, 16000):
        w = synth_wave(5, n, 1)
        for variant, name in ((0, "torchaudio"), (1, "speechbrain")):
            for mn in (True, False):
                ref = torch.from_numpy(fo.fbank_batch(w, mean_nor=mn)) if variant == 0 else eo.fbank_speechbrain(torch.from_numpy(w), mean_norm=mn)
                got = se.fbank_batch_device(torch.from_numpy(w).to(dev), variant=variant, mean_nor=mn)
                r, m = rel(got, ref)
                print(f"fbank {name} n={n} mean_norm={mn}: rel={r:.2e} maxabs={m:.2e} shape={tuple(got.shape)}")

if "trunk" in what or "e2e" in what:
    t0 = time.time(); model = eo.make_random_ecapa(0); print("oracle model", time.time() - t0)
    enc = se.EcapaEncoderB200(model.state_dict(), device=dev, max_batch=64, max_samples=24000)

if "trunk" in what:
    for (B, T) in ((3, 151), (5, 101), (2, 26)):
        x = eo.synth_features(B, T, 7)
        taps = {}
        with torch.inference_mode():
            ref = model(x, taps).squeeze(1)
        got = enc.forward_feats(x.to(dev))
        torch.cuda.synchronize()
        for name in ("feats", "block0", "b1.out", "b2.out", "b3.out", "mfa", "pooled"):
            g = enc.debug_fetch(name, B, T)
            if name == "feats": r_ = x
            elif name == "pooled": r_ = taps[name]
            else: r_ = taps[name].transpose(1, 2)
            r, m = rel(g, r_)
            print(f"  B={B} T={T} {name:8s} rel={r:.2e} maxabs={m:.2e}")
        cos = torch.nn.functional.cosine_similarity(got.cpu(), ref, dim=1)
        r, m = rel(got, ref)
        print(f"trunk B={B} T={T}: 1-cos max={float((1-cos).max()):.2e} rel={r:.2e}")

if "e2e" in what:
    w = synth_wave(6, 24000, 3)
    with torch.inference_mode():
        ref = eo.encode_batch(model, torch.from_numpy(w)).squeeze(1)
    got = enc.encode_batch(torch.from_numpy(w)).squeeze(1).cpu()
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    print(f"e2e encode_batch: 1-cos max={float((1-cos).max()):.2e} rel={rel(got, ref)[0]:.2e}")
    # throughput smoke
    B = 64
    wb = torch.from_numpy(synth_wave(B, 24000, 4)).to(dev)
    for _ in range(2): enc.embed_device(wb, 24000, B, 24000)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): enc.embed_device(wb, 24000, B, 24000)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 5
    print(f"embed B={B}: {dt*1e3:.2f} ms -> {B/dt:.0f} emb/s")

def synth_emb(N, K, sigma, seed):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((K, 192)); c /= np.linalg.norm(c, axis=1, keepdims=True)
    lab = rng.integers(0, K, N)
    return (c[lab] + sigma * rng.standard_normal((N, 192))).astype(np.float32), lab

if "aff" in what:
    for N in (300, 1000, 5000):
        X, _ = synth_emb(N, 8, 0.02, 1)
        ref = co.cosine_distance(X)
        got = cl.cosine_distance_device(torch.from_numpy(X).to(dev)).cpu().numpy()
        print(f"affinity N={N}: max|d|={np.abs(got-ref).max():.2e} sym={np.abs(got-got.T).max():.1e}")
    X, _ = synth_emb(1000, 8, 0.02, 1)
    got = cl.cosine_distance_device(torch.from_numpy(X).to(dev), 256, 300).cpu().numpy()
    print("rowblock max|d|", np.abs(got - co.cosine_distance(X)[256:556]).max())

if "ahc" in what:
    for (N, K, sigma) in ((7, 2, 0.02), (400, 5, 0.02), (300, 6, 0.05), (2000, 8, 0.02), (5000, 8, 0.02)):
        X, lab = synth_emb(N, K, sigma, 3)
        t0 = time.time(); ref = co.cluster_embeddings(X, "agglo", 0.68); tc = time.time() - t0
        xd = torch.from_numpy(X).to(dev)
        got = cl.cluster_embeddings_device(xd, 0.68); torch.cuda.synchronize()
        t0 = time.time(); got = cl.cluster_embeddings_device(xd, 0.68); torch.cuda.synchronize(); tg = time.time() - t0
        got = got.cpu().numpy()
        print(f"ahc N={N} K={K} sigma={sigma}: same_partition={co.same_partition(got, ref)} n_ref={len(set(ref))} n_got={len(set(got))} cpu={tc*1e3:.1f}ms gpu={tg*1e3:.2f}ms")
