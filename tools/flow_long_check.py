import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import flow_ref as FR
from gonova_tts_b200 import B200Flow
sd = FR.random_state_dict(0); est = FR.load_estimator(sd)
dev = "cuda:0"
f = B200Flow(sd, device=dev, dtype="bf16")
for B, T, lens in ((1, 1500, None), (2, 1031, [1031, 700]), (1, 3, None), (3, 129, [129, 1, 128])):
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=21, lengths=lens)
    want = FR.solve_euler(est, z * mask, mu, mask, spks, cond, n_timesteps=1).numpy()
    got = f.decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), lengths=lens, n_timesteps=1).cpu().numpy()
    zn = (z * mask).numpy()
    d = got - want
    snr = 10 * np.log10(((want - zn) ** 2).sum() / max((d ** 2).sum(), 1e-30))
    print(B, T, lens, "max-abs", np.abs(d).max(), "SNR of velocity", round(float(snr), 1), "finite", np.isfinite(got).all())
