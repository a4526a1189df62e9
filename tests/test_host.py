"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from gonova_tts_b200 import _cabi, chunk_plan, fold_state_dict, random_state_dict, round_robin, shard_range
from gonova_tts_b200.weights import layer_specs
from oracle import hift_ref as R
from oracle import tail_ref as TR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gonova_hift.h")).read()
    declared = sorted(set(re.findall(r"\b(gnv_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no entry points found in include/gonova_hift.h"
    raw = ctypes.CDLL(str(_cabi.lib_path()))
    for name in declared:
        assert hasattr(raw, name), f"{name} is declared in the header but not exported"
    assert sorted(_cabi.SYMBOLS) == declared
    assert lib.gnv_abi_version() == 2


def test_create_fails_loudly_without_a_gpu_or_weights(lib):
    h = ctypes.c_void_p()
    rc = lib.gnv_create(None, 0, 0, 0, 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert "weights" in _cabi.last_error(None)
    if not torch.cuda.is_available():
        from gonova_tts_b200 import B200HiFT

        with pytest.raises(RuntimeError):
            B200HiFT(random_state_dict(0), device="cuda:0")
        with pytest.raises(RuntimeError):
            B200HiFT(random_state_dict(0), device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gonova_tts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "oracle/" not in src or f.endswith((".cu", ".cuh")), f"{f} references oracle/"


def test_fold_matches_torch_weight_norm_both_key_styles():
    sd = random_state_dict(3, corners=True)
    folded = fold_state_dict(sd)
    m = R.load_model(sd)
    want = R.folded_weights(m)
    assert set(folded) == set(want)
    for k in want:
        np.testing.assert_allclose(folded[k].numpy(), want[k].numpy(), rtol=2e-6, atol=1e-8, err_msg=k)
    old = {}
    for k, v in sd.items():
        k2 = k.replace(".parametrizations.weight.original0", ".weight_g").replace(
            ".parametrizations.weight.original1", ".weight_v")
        old["mel2wav." + k2] = v
    folded_old = fold_state_dict(old, prefix="mel2wav.")
    for k in want:
        assert torch.equal(folded_old[k], folded[k])
    # ConvTranspose1d weight-norm is per INPUT channel (dim 0 of [Cin, Cout, k])
    g = sd["ups.0.parametrizations.weight.original0"]
    assert g.shape == (512, 1, 1)


def test_fold_rejects_missing_and_misshapen_tensors():
    sd = random_state_dict(0)
    bad = dict(sd)
    del bad["conv_post.bias"]
    with pytest.raises(KeyError):
        fold_state_dict(bad)
    bad = dict(sd)
    bad["source_downs.0.weight"] = torch.zeros(256, 18, 29)
    with pytest.raises(ValueError):
        fold_state_dict(bad)


def test_layer_specs_parameter_count():
    total = 0
    for path, kind, shape, wn in layer_specs():
        total += int(np.prod(shape))
        if kind != "snake":
            total += shape[1] if kind == "convT" else shape[0]
    assert total == 20_806_557


def test_chunk_plan_is_the_oracle_plan():
    for T, chunk, halo in [(3000, 100, 16), (117, 100, 16), (99, 100, 16), (1, 100, 16), (500, 64, 20)]:
        assert list(chunk_plan(T, chunk, halo)) == list(TR.chunk_plan(T, chunk, halo))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            rr = round_robin(n, world)
            assert sorted(i for part in rr for i in part) == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_plain_c_host_is_built_and_write_flat_round_trips(lib, tmp_path):
    """build() also produces examples/c_host.c (the C ABI bound from plain C); its weight file format round-trips."""
    import struct
    import subprocess

    import numpy as np

    from gonova_tts_b200 import build, random_state_dict
    from gonova_tts_b200.weights import fold_state_dict, write_flat

    assert build.C_HOST.exists()
    r = subprocess.run([str(build.C_HOST)], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr              # loads (the library resolves) and asks for arguments
    sd = random_state_dict(0, False)
    n = write_flat(sd, str(tmp_path / "w.bin"))
    folded = fold_state_dict(sd)
    assert n == len(folded)
    raw = (tmp_path / "w.bin").read_bytes()
    off = 4
    for name in sorted(folded):
        (ln,) = struct.unpack_from("<i", raw, off); off += 4
        assert raw[off:off + ln].decode() == name; off += ln
        (nd,) = struct.unpack_from("<i", raw, off); off += 4
        shape = struct.unpack_from("<%dq" % nd, raw, off); off += 8 * nd
        assert tuple(shape) == tuple(folded[name].shape)
        cnt = int(np.prod(shape))
        data = np.frombuffer(raw, dtype="<f4", count=cnt, offset=off); off += 4 * cnt
        np.testing.assert_array_equal(data, folded[name].numpy().ravel())
    assert off == len(raw)
