"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

PARITY UNPINNED: the reference (/root/reference) holds no golden vectors, fixtures or tests for the
decoder and its engine is not installable here, so these vectors are produced by
this repo's own oracle.  They pin the oracle against regressions and give the GPU tests fixed
known-answer outputs; they are not outputs of the reference itself."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hift_ref as R, tail_ref as TR  # noqa: E402
from gonova_tts_b200.weights import random_state_dict  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for corners in (False, True):
        sd = random_state_dict(0, corners)
        m = R.load_model(sd)
        for T in (8, 32):
            mel = R.synthetic_mel(2, T, seed=1234 + T)
            s = R.synthetic_source(m, mel, seed=4321 + T)
            taps = {}
            with torch.inference_mode():
                wav = m.decode(mel, s, taps=taps)
                f0 = m.f0_predictor(mel)
            name = f"decode_T{T}_{'corners' if corners else 'plain'}.npz"
            np.savez_compressed(
                os.path.join(OUT, name), wav=wav.numpy(), f0=f0.numpy(),
                conv_post=taps["conv_post"].numpy()[:, :, ::7].copy(),
                stage0_mean=taps["stage0"].numpy().mean(axis=2), s_head=s.numpy()[:, 0, :64].copy())
            print(name, wav.shape, float(wav.abs().max()))
    # streaming tail known answers
    x = np.array([0.0, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.5 / 32767, -1.5 / 32767, 0.99, -0.99, 1.0, -1.0,
                  1.5, -1.5, 3.0517578125e-05, 0.123456789, -0.987654321], dtype=np.float32)
    rng = np.random.default_rng(7)
    xr = (rng.standard_normal(4096) * 0.4).astype(np.float32)
    cur = np.stack([xr[:2048], xr[2048:]])
    prev = (rng.standard_normal((2, 480)) * 0.4).astype(np.float32)
    w = TR.fade_window(480)
    f_cf, i_cf = TR.pcm_tail(cur, prev, w, 0.99)
    f_tf, i_tf = TR.pcm_tail(cur, None, TR.trim_fade_window(), 0.99)
    np.savez_compressed(os.path.join(OUT, "tail_kat.npz"), x=x, x_i16=TR.pack_i16(x), cur=cur, prev=prev, w=w,
                        f_cf=f_cf, i_cf=i_cf, f_tf=f_tf, i_tf=i_tf)
    print("tail_kat.npz", TR.pack_i16(x))


if __name__ == "__main__":
    main()
