"""The launch-plan cache behind the C ABI under the reference service's traffic shape.

The reference decodes ONE sentence per call (services/tts/server.py:118-182; synthesizer.py:327-359), so the
decoder sees a new T on almost every call.  A plan (tensor maps + tile lists) is per (B, T, workspace); these tests
pin down that a long run of distinct lengths neither grows device memory nor changes results, that the Python
shim's length bucketing is invisible in the output, and that a CUDA graph keeps its plan."""
import pytest
import torch

from oracle import hift_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    from gonova_tts_b200 import random_state_dict

    return random_state_dict(0, False)


def test_many_sentence_lengths_reuse_slots_and_results_survive_eviction(lib, cuda_device, sd):
    from gonova_tts_b200 import B200HiFT

    dec = B200HiFT(sd, device=cuda_device, dtype="bf16")
    dec.reserve(1, 200)                                   # one workspace address for the whole run
    mel = R.synthetic_mel(1, 200, seed=11).to(cuda_device)
    g = torch.Generator().manual_seed(12)
    s = (torch.rand(1, 1, 200 * 480, generator=g) * 0.2 - 0.1).to(cuda_device)

    def decode(T):
        return dec.decode(mel[:, :, :T].contiguous(), s[:, :, : T * 480].contiguous())

    first = {T: decode(T).clone() for T in (3, 5, 40)}
    for T in range(1, 201):                               # 200 distinct lengths through an LRU of 128
        decode(T)
    st = dec.plan_stats()
    assert st["cached"] <= 128, st
    assert st["built"] >= 200, st
    assert st["slots"] <= 129, st                          # evicted plans hand their device slot on: no growth
    for T, want in first.items():                         # rebuilt after eviction: bit-identical
        assert torch.equal(decode(T), want), T
    before = dec.plan_stats()["slots"]
    for T in range(1, 201):
        decode(T)
    assert dec.plan_stats()["slots"] == before


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_bucketed_inference_is_invisible(lib, cuda_device, sd, dtype):
    """bucket_frames rounds T up and masks through `lengths`: waveform and source of the caller's T frames are
    exactly those of the un-bucketed call, and the plan cache sees one shape per bucket."""
    from gonova_tts_b200 import B200HiFT

    exact = B200HiFT(sd, device=cuda_device, dtype=dtype)
    bucketed = B200HiFT(sd, device=cuda_device, dtype=dtype, bucket_frames=8)
    worst = 0.0
    for T in (1, 7, 8, 9, 33, 117, 233):
        mel = R.synthetic_mel(2, T, seed=T).to(cuda_device)
        wav_e, src_e = exact.inference(mel, seed=5)
        wav_b, src_b = bucketed.inference(mel, seed=5)
        assert wav_b.shape == wav_e.shape == (2, T * 480) and src_b.shape == src_e.shape
        assert wav_b.is_contiguous() and src_b.is_contiguous()
        assert torch.equal(src_b, src_e), T
        worst = max(worst, float((wav_b - wav_e).abs().max()))
    print(f"[bucket] {dtype}: max |bucketed - exact| = {worst:.3e}")
    assert worst == 0.0
    assert bucketed.plan_stats()["built"] == 5            # buckets 8, 16, 40, 120, 240


def test_graph_pins_its_plan_and_first_call_in_capture_is_refused(lib, cuda_device, sd, monkeypatch):
    from gonova_tts_b200 import B200HiFT, GraphedInference

    monkeypatch.setenv("GONOVA_MAX_PLANS", "16")          # a small cache, so that the shapes below really evict
    dec = B200HiFT(sd, device=cuda_device, dtype="bf16")
    monkeypatch.delenv("GONOVA_MAX_PLANS")
    dec.reserve(1, 100)
    gi = GraphedInference(dec, 1, 24, seed=3)
    mel = R.synthetic_mel(1, 24, seed=1).to(cuda_device)
    want = gi(mel).clone()
    assert dec.plan_stats()["pinned"] == 1
    for T in range(25, 101):                              # 76 other shapes: enough to evict anything evictable
        dec.inference(R.synthetic_mel(1, T, seed=T).to(cuda_device), seed=3)
    st = dec.plan_stats()
    assert st["pinned"] == 1 and st["cached"] <= 17 and st["built"] >= 77, st
    gi.reseed(3)                                          # (every replay bumps the graph's device seed)
    assert torch.equal(gi(mel), want)                     # the graph's tensor maps were not recycled

    fresh = B200HiFT(sd, device=cuda_device, dtype="bf16")
    fresh.reserve(1, 16)
    m16 = R.synthetic_mel(1, 16, seed=2).to(cuda_device)
    wav = torch.empty(1, 16 * 480, device=cuda_device)
    src = torch.empty(1, 1, 16 * 480, device=cuda_device)
    graph = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="capture"):
        with torch.cuda.graph(graph):
            fresh.inference(m16, seed=1, out=wav, source_out=src)
    torch.cuda.synchronize()
    w2, _ = fresh.inference(m16, seed=1)                  # and the handle still works afterwards
    assert torch.isfinite(w2).all()


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_result_does_not_depend_on_workspace_contents(lib, cuda_device, sd, dtype):
    """The workspace is scratch: a decode (ragged or not) over a workspace full of NaN bit patterns equals the same
    decode over a zeroed one, bit for bit — no layer reads a row that the same call did not write."""
    from gonova_tts_b200 import B200HiFT

    dec = B200HiFT(sd, device=cuda_device, dtype=dtype)
    lens = [40, 17, 1, 33]
    B, T = len(lens), max(lens)
    mel = R.synthetic_mel(B, T, seed=21).to(cuda_device)
    dec.reserve(B, T)
    outs = []
    for fill in (0x00, 0xFF, 0x7F):
        dec._ws.fill_(fill)
        wav, src = dec.inference(mel, seed=9, lengths=lens)
        assert torch.isfinite(wav).all() and torch.isfinite(src).all(), hex(fill)
        outs.append((wav.clone(), src.clone()))
        dec._ws.fill_(fill)
        full, _ = dec.inference(mel, seed=9)
        assert torch.isfinite(full).all(), hex(fill)
        outs[-1] += (full.clone(),)
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    for b, n in enumerate(lens):
        assert not outs[0][0][b, n * 480:].any()
        # ... and every row of the ragged batch is bit-identical to that utterance decoded alone at its own length
        # (tiles wholly past an utterance's end are skipped, not computed and discarded)
        alone, src1 = dec.inference(mel[b:b + 1, :, :n].contiguous(), seed=9)
        if b == 0:                                        # the noise stream is keyed by the row index: row 0 matches
            assert torch.equal(src1[0, 0], outs[0][1][0, 0, : n * 480])
            assert torch.equal(alone[0], outs[0][0][0, : n * 480])
        one = dec.decode(mel[b:b + 1, :, :n].contiguous(), outs[0][1][b:b + 1, :, : n * 480].contiguous())
        assert torch.equal(one[0], outs[0][0][b, : n * 480]), (b, n)


def test_out_of_range_lengths_are_clamped_not_trusted(lib, cuda_device, sd):
    """`lengths` is device data the C ABI cannot inspect before launching: 0, negative and too-large entries behave as
    clamp(lengths, 0, T) — a silent row for 0, the whole row for > T — and never touch memory outside the batch."""
    from gonova_tts_b200 import B200HiFT

    dec = B200HiFT(sd, device=cuda_device, dtype="bf16")
    T = 24
    mel = R.synthetic_mel(4, T, seed=31).to(cuda_device)
    g = torch.Generator().manual_seed(32)
    s = (torch.rand(4, 1, T * 480, generator=g) * 0.2 - 0.1).to(cuda_device)
    guard = torch.full((1 << 20,), 7.0, device=cuda_device)              # allocated next to the outputs
    wav = dec.decode(mel, s, lengths=[0, -5, 1000, 7])
    torch.cuda.synchronize()
    assert torch.isfinite(wav).all() and torch.all(guard == 7.0)
    assert not wav[0].any() and not wav[1].any()
    full = dec.decode(mel[2:3].contiguous(), s[2:3].contiguous())
    assert torch.equal(wav[2], full[0])
    part = dec.decode(mel[3:4, :, :7].contiguous(), s[3:4, :, : 7 * 480].contiguous())
    assert torch.equal(wav[3, : 7 * 480], part[0]) and not wav[3, 7 * 480:].any()
    w2, s2 = dec.inference(mel, lengths=[0, -5, 1000, 7], seed=4)
    assert torch.isfinite(w2).all() and torch.isfinite(s2).all() and not w2[0].any() and not w2[1].any()


def test_large_ragged_batch_rows_equal_single_decodes(lib, cuda_device, sd):
    """A ragged batch big enough for the throughput schedule (B*T > 4096 frames: no forked streams, CTA pairs, two
    accumulators per tile, dead tiles skipped) against the same utterances decoded alone (small-problem schedule:
    narrow tiles, forked streams): bit-identical rows, zeros after each length."""
    from gonova_tts_b200 import B200HiFT

    dec = B200HiFT(sd, device=cuda_device, dtype="bf16")
    B, T = 24, 300
    g = torch.Generator().manual_seed(44)
    lens = torch.randint(20, T + 1, (B,), generator=g).tolist()
    lens[0], lens[1] = T, 1
    mel = R.synthetic_mel(B, T, seed=45).to(cuda_device)
    s = (torch.rand(B, 1, T * 480, generator=g) * 0.2 - 0.1).to(cuda_device)
    wav = dec.decode(mel, s, lengths=lens)
    assert torch.isfinite(wav).all()
    for b in (0, 1, 5, 11, 23):
        n = lens[b]
        one = dec.decode(mel[b:b + 1, :, :n].contiguous(), s[b:b + 1, :, : n * 480].contiguous())
        assert torch.equal(one[0], wav[b, : n * 480]), (b, n)
        assert not wav[b, n * 480:].any()


def test_forked_schedule_back_to_back_calls_equal_the_serial_schedule(lib, cuda_device, sd, monkeypatch):
    """Small problems run their independent branches on streams the handle owns.  Many calls of changing shapes, queued
    without a host sync in between (the next call's side streams start while nothing of the previous call may still be
    read or written), against a decoder with the fork switched off: bit-identical, also through `inference`."""
    from gonova_tts_b200 import B200HiFT

    forked = B200HiFT(sd, device=cuda_device, dtype="bf16")
    monkeypatch.setenv("GONOVA_FORK_MAX_FRAMES", "0")
    serial = B200HiFT(sd, device=cuda_device, dtype="bf16")
    monkeypatch.delenv("GONOVA_FORK_MAX_FRAMES")
    forked.reserve(4, 200)
    serial.reserve(4, 200)
    g = torch.Generator().manual_seed(3)
    shapes = [(int(torch.randint(1, 5, (1,), generator=g)), int(torch.randint(1, 201, (1,), generator=g))) for _ in range(40)]
    mels = [R.synthetic_mel(B, T, seed=i).to(cuda_device) for i, (B, T) in enumerate(shapes)]
    got = [forked.inference(m, seed=7) for m in mels]                 # all queued back to back
    want = [serial.inference(m, seed=7) for m in mels]
    torch.cuda.synchronize()
    for (B, T), (wa, sa), (wb, sb) in zip(shapes, got, want):
        assert torch.equal(sa, sb) and torch.equal(wa, wb), (B, T)
    assert torch.isfinite(torch.cat([w.flatten() for w, _ in got])).all()
