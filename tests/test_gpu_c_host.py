"""The drop-in boundary is a C ABI: a host written in plain C (examples/c_host.c: gcc, gonova_hift.h, libcudart — no
Python, no torch, no C++) creates the decoder from a flat weight file, decodes a batch on its own stream with its own
cudaMalloc'ed buffers, and gets bit for bit what the Python shim gets from the same library."""
import struct
import subprocess

import numpy as np
import pytest
import torch

from oracle import hift_ref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,code", [("bf16", 1), ("tf32", 0)])
def test_plain_c_host_equals_python_shim(lib, cuda_device, tmp_path, dtype, code):
    from gonova_tts_b200 import B200HiFT, build, random_state_dict
    from gonova_tts_b200.weights import write_flat

    exe = build.C_HOST
    assert exe.exists(), "build() must produce the C host next to the library"
    sd = random_state_dict(0, False)
    n = write_flat(sd, str(tmp_path / "w.bin"))
    assert n > 100
    B, T, seed = 3, 37, 11
    mel = R.synthetic_mel(B, T, seed=5)
    with open(tmp_path / "mel.bin", "wb") as f:
        f.write(struct.pack("<ii", B, T))
        f.write(mel.contiguous().numpy().tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "w.bin"), str(tmp_path / "mel.bin"), str(tmp_path / "wav.bin"),
                        str(code), str(seed)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "plans built 1" in r.stdout, r.stdout                 # two calls, one plan
    got = np.fromfile(tmp_path / "wav.bin", dtype=np.float32).reshape(B, T * 480)
    want, _ = B200HiFT(sd, device=cuda_device, dtype=dtype).inference(mel.to(cuda_device), seed=seed)
    np.testing.assert_array_equal(got, want.cpu().numpy())
    with torch.inference_mode():                                 # ... and both are the oracle's waveform within tolerance
        m = R.load_model(sd)
        _, src = B200HiFT(sd, device=cuda_device, dtype=dtype).inference(mel.to(cuda_device), seed=seed)
        ref = m.decode(mel, src.cpu()).numpy()
    assert np.abs(got - ref).max() <= (1e-3 if dtype == "tf32" else 1e-2)


def test_c_host_reports_errors_through_the_abi(lib, tmp_path):
    from gonova_tts_b200 import build

    (tmp_path / "w.bin").write_bytes(struct.pack("<i", 1) + struct.pack("<i", 3) + b"foo" + struct.pack("<i", 1) +
                                     struct.pack("<q", 2) + struct.pack("<2f", 1.0, 2.0))
    (tmp_path / "mel.bin").write_bytes(struct.pack("<ii", 1, 2) + np.zeros(160, np.float32).tobytes())
    r = subprocess.run([str(build.C_HOST), str(tmp_path / "w.bin"), str(tmp_path / "mel.bin"), str(tmp_path / "o.bin")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "gnv_create" in r.stderr        # a missing tensor is an error string, not a crash
