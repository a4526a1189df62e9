#!/usr/bin/env python
"""bench.py — audio-seconds per second of the waveform decoder (mel -> int16 PCM) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W                      # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # the CPU path, same metric/config

One step = one pass of the hot path over one batch of synthetic mel frames per GPU:
  f0 predictor -> NSF source -> HiFT decode (conv GEMMs, STFT, iSTFT head) -> clamp + int16 pack.
Workload per GPU = BASELINE.json configs[2]: 64 concurrent 10 s utterances (T = 500 mel frames each),
bf16 operands / fp32 accumulate; at N GPUs every rank decodes its own 64 streams (request-level data
parallelism, no collective on the data path; N = 8 is configs[3]'s 512 streams) -> "scaling": "weak".
`value` times the device-resident path; `e2e` times the public API call with HOST buffers (pinned
mel in, int16 PCM out, both copies inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 24000
SPF = 480
METRIC = "audio-sec/sec"
UNIT = "audio-s/s (x real-time)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=500, help="mel frames per utterance (50 per audio second)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-first-chunk", action="store_true")
    ap.add_argument("--profile-table", default="", help="write the per-launch timing table to this path")
    ap.add_argument("--streams", type=int, default=512,
                    help="BASELINE configs[3]: this many 10 s streams IN TOTAL, sharded over the ranks in batches of "
                         "--batch (strong scaling; reported as the `streams512` block; 0 = skip)")
    ap.add_argument("--no-tf32", action="store_true", help="skip the tf32 (reference-precision) block")
    ap.add_argument("--no-stock-torch", action="store_true", help="skip the stock-PyTorch-on-this-GPU block")
    ap.add_argument("--no-flow", action="store_true", help="skip the flow-decoder (SURVEY 8f-1) block")
    return ap.parse_args()


# Algorithmic bytes per mel frame of the bandwidth-bound kernels (eb = bytes per conv operand element).
def aux_bytes_per_frame(name: str, eb: int):
    if name == "istft_head":
        return 18 * 120 * 4 + 1920                 # conv_post rows in (fp32), 480 samples out
    if name == "stft":
        return 1920 + 18 * 120 * eb                # source samples in, 18 channels x 120 frames out as conv operands
    if name == "m_source":
        return 4 + 1920                            # f0 in, 480 source samples out; the [9, 480] bank stays on chip
    if name == "pcm_tail":
        return 1920 + 960                          # 6 B per sample: fp32 in, int16 out
    if name == "pack_mel":
        return 80 * 4 + 80 * eb
    if name == "f0_predictor.classifier":
        return 512 * eb + 4
    return None


def measure_tf32_gemm_peak(dev, seconds: float = 1.5):
    """TF32 tensor-core GEMM rate of THIS GPU, measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3
    (2 N^3 flops) on fp32 operands with TF32 allowed; burst = best of 10, sustained = back to back for `seconds`."""
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        per = max(best, 1e-3)
        reps = max(10, int(seconds * 1e3 / per))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(dev)
        sustained_ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"tf32_tflops": fl / (best / 1e3) / 1e12, "tf32_tflops_sustained": fl / (sustained_ms / 1e3) / 1e12,
                "how": f"torch.matmul fp32 {n}^3 with allow_tf32 (2*N^3): best of 10 (burst) and {reps} back to back (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def synthetic_mel(B: int, T: int, seed: int) -> torch.Tensor:
    """Log-mel-like host input (same recipe as the parity tests): clamp(-5 + 2 randn, log 1e-5, 2.5),
    3-tap moving average along time."""
    g = torch.Generator().manual_seed(seed)
    mel = torch.clamp(-5 + 2 * torch.randn(B, 80, T, generator=g), -11.513, 2.5)
    pad = torch.nn.functional.pad(mel, (1, 1), mode="replicate")
    return ((pad[..., :-2] + pad[..., 1:-1] + pad[..., 2:]) / 3).contiguous()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
            "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the reference decoder on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_decode_rate(B: int, T: int, reps: int, warm: int = 1):
    """Oracle (fp32 torch CPU restatement of the engine's HiFTGenerator + trim_fade) timed on all host
    cores.  Returns (audio-s/s, seconds per call, cores)."""
    from oracle import hift_ref as R
    from gonova_tts_b200.weights import random_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = R.load_model(random_state_dict(0, False))
    mel = synthetic_mel(B, T, 1234)
    gen = torch.Generator().manual_seed(1)
    for _ in range(warm):
        R.hift_inference(model, mel, generator=gen)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        wav, _ = R.hift_inference(model, mel, generator=gen)
        wav.numpy().astype("float32").tobytes()            # the reference's float32 wire format
        times.append(time.perf_counter() - t0)
    best = min(times)
    return B * T / 50.0 / best, best, cores, times


def run_reference(args, rank):
    if rank != 0:
        return
    B, T = 1, args.frames
    from oracle import hift_ref as R
    from gonova_tts_b200.weights import random_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = R.load_model(random_state_dict(0, False))
    mel = synthetic_mel(B, T, 1234)
    gen = torch.Generator().manual_seed(1)
    for _ in range(args.warmup):
        R.hift_inference(model, mel, generator=gen)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav, _ = R.hift_inference(model, mel, generator=gen)
        wav.numpy().astype("float32").tobytes()
    dt = time.perf_counter() - t0
    val = args.steps * B * T / 50.0 / dt
    sample = f"{B} utterance x {T} mel frames ({B * T / 50:.0f} audio-s) per step, fp32, torch CPU (oneDNN), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"64 concurrent {T / 50:.0f} s utterances per GPU through the decoder "
                               "(BASELINE configs[2]; N=8 is configs[3]'s 512 streams): f0 -> source -> HiFT decode -> int16",
                   "utterances_per_gpu": 64, "frames_per_utterance": T, "sample_rate": SR,
                   "cpu_arm": "decodes a bounded sample of that workload per step (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference engine (chatterbox) is not installable here; this is oracle/hift_ref.py, the CPU "
                "restatement of its HiFTGenerator, run on the host cores",
    }))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the decoder has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    import torch.distributed as dist

    from gonova_tts_b200 import B200HiFT, pcm_tail, random_state_dict
    from gonova_tts_b200 import _cabi

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    B, T = args.batch, args.frames
    audio_s_per_gpu = B * T / 50.0
    dec = B200HiFT(random_state_dict(0, False), device=dev, dtype=args.dtype)
    # this rank's own streams (shard `rank` of world*B): different seed -> different utterances
    mel_host = synthetic_mel(B, T, 1234 + rank).pin_memory()
    mel = mel_host.to(dev)
    L = T * SPF
    wav = torch.empty(B, L, dtype=torch.float32, device=dev)
    src = torch.empty(B, 1, L, dtype=torch.float32, device=dev)
    pcm = torch.empty(B, L, dtype=torch.int16, device=dev)
    pcm_host = torch.empty(B, L, dtype=torch.int16).pin_memory()
    ws_bytes = dec.workspace_bytes(B, T)
    stream = torch.cuda.current_stream(dev)

    def step_resident(i):
        dec.inference(mel, seed=i + 1, out=wav, source_out=src)
        pcm_tail(wav, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm)

    # End to end: every step moves its own inputs host -> device and its own PCM device -> host (pinned memory, inside
    # the timed region).  The copies run on a second stream, double-buffered, so step i's copies overlap step i+-1's
    # kernels the way a serving loop would; the host "consumes" step i-1's PCM (event sync) while step i is in flight.
    copy_stream = torch.cuda.Stream(device=dev)
    mel_bufs = [torch.empty_like(mel) for _ in range(2)]
    pcm_bufs = [torch.empty_like(pcm) for _ in range(2)]
    pcm_hosts = [torch.empty(B, L, dtype=torch.int16).pin_memory() for _ in range(2)]
    ev_h2d = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_d2h = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"n": 0}

    def step_e2e(i):
        k = e2e_state["n"] & 1
        n = e2e_state["n"]
        with torch.cuda.stream(copy_stream):
            if n >= 2:
                copy_stream.wait_event(ev_done[k])         # step n-2 no longer reads mel_bufs[k] / writes pcm_bufs[k]
            mel_bufs[k].copy_(mel_host, non_blocking=True)
            ev_h2d[k].record(copy_stream)
        stream.wait_event(ev_h2d[k])
        if n >= 2:
            stream.wait_event(ev_d2h[k])                   # pcm_bufs[k] of step n-2 has left the device
        dec.inference(mel_bufs[k], seed=i + 1, out=wav, source_out=src)
        pcm_tail(wav, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm_bufs[k])
        ev_done[k].record(stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[k])
            pcm_hosts[k].copy_(pcm_bufs[k], non_blocking=True)
            ev_d2h[k].record(copy_stream)
        if n >= 1:
            ev_d2h[k ^ 1].synchronize()                    # the host consumes the previous step's PCM
        e2e_state["n"] = n + 1

    def finish_e2e():
        if e2e_state["n"]:
            stream.wait_event(ev_d2h[(e2e_state["n"] - 1) & 1])   # the last step's PCM is on the host before the clock stops

    def timed(fn, steps, warmup, sampler=None, finish=None):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize(dev)
        barrier()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        if finish:
            finish()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, clocks = timed(step_resident, args.steps, args.warmup, sampler)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup), None, finish_e2e)
    value = world * audio_s_per_gpu * args.steps / (ms_total / 1e3)
    e2e_value = world * audio_s_per_gpu * args.steps / (ms_e2e / 1e3)

    # ---- BASELINE configs[3] as SURVEY 8d defines it: `--streams` streams IN TOTAL, sharded over the ranks and decoded in
    # batches of B; aggregate = total audio seconds / max-over-ranks device time (strong scaling: the job is fixed).
    streams512 = None
    if args.streams > 0:
        from gonova_tts_b200 import shard_range

        lo, hi = shard_range(args.streams, world, rank)
        mine = hi - lo
        sizes = [min(B, mine - b0) for b0 in range(0, mine, B)]

        def step_job(i):
            for j, nb in enumerate(sizes):
                dec.inference(mel[:nb], seed=1000 * (i + 1) + j, out=wav[:nb], source_out=src[:nb])
                pcm_tail(wav[:nb], None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm[:nb])

        job_steps = 3
        ms_job, _ = timed(step_job, job_steps, 1)
        streams512 = {
            "streams_total": args.streams, "streams_this_rank": mine, "batches_per_rank": len(sizes), "batch": B,
            "audio_seconds": args.streams * T / 50.0, "ms_per_job": ms_job / job_steps,
            "value": args.streams * T / 50.0 * job_steps / (ms_job / 1e3), "unit": UNIT, "scaling": "strong",
            "how": "SURVEY 8d config 4: the whole job (all streams once) per step, inputs resident, CUDA events, max over "
                   "ranks; no collective on the data path"}

    out = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        # ---- per-launch table: which kernel dominates, and its achieved rate ----
        rows_acc = None
        n_prof = int(os.environ.get("GONOVA_BENCH_NPROF", "3"))
        for r in range(n_prof + 1):
            _, rows = dec.profile_inference(mel, seed=100 + r)
            if r == 0:
                continue                                   # warm the event path
            if rows_acc is None:
                rows_acc = [[n, k, 0.0, f] for (n, k, _, f) in rows]
            for acc, row in zip(rows_acc, rows):
                acc[2] += row[2] / n_prof
        tc = [r for r in rows_acc if r[1] == _cabi.LAUNCH_CONV_TC]
        simt = [r for r in rows_acc if r[1] == _cabi.LAUNCH_CONV_SIMT]
        aux = [r for r in rows_acc if r[1] == _cabi.LAUNCH_AUX]
        tc_ms, tc_flops = sum(r[2] for r in tc), sum(r[3] for r in tc)
        step_ms_profiled = sum(r[2] for r in rows_acc)
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        tf32_peak = None
        if args.dtype == "tf32":
            tf32_peak = measure_tf32_gemm_peak(dev)
            peak_tf = tf32_peak["tf32_tflops_sustained"]
        elif args.dtype != "bf16":
            peak_tf = 75.0                                              # fp32 FMA nominal
        # The family's duration INSIDE the timed region = its share of the step (per-launch CUDA events of the profiled
        # passes above; the ncu launch list in profiles/ gives the same share) x the timed step.  The profiled passes
        # themselves run without launch overlap and, after the timed loops, deeper in the power cap: their absolute
        # sum is reported next to it (kernel_ms_per_step_profiled) but is not what the timed region saw.
        share = tc_ms / step_ms_profiled if step_ms_profiled else 0.0
        tc_ms_profiled = tc_ms
        tc_ms = share * (ms_total / args.steps)
        achieved_tf = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        # DRAM bytes of the same launches from the committed ncu capture (profiles/), valid for the default workload
        traffic, traffic_src = None, None
        try:
            tname = "r02_conv_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_conv_traffic.json")) \
                else "r01_conv_traffic.json"
            with open(os.path.join(ROOT, "profiles", tname)) as f:
                tj = json.load(f)
            if (B, T, args.dtype) == (64, 500, "bf16"):
                traffic, traffic_src = tj["traffic_bytes"], f"profiles/{tname} (ncu dram__bytes_read+write, all conv launches of one step)"
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = {
            "kernel": f"conv_tc2_kernel + conv_pair_kernel <{args.dtype}> (persistent tcgen05 implicit-GEMM convs; "
                      f"{len(tc)} launches per step, timed as a family)",
            "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
            "hbm_view": None if traffic is None else {
                "achieved_GBps": traffic / (tc_ms / 1e3) / 1e9, "peak_GBps": hbm_peak,
                "frac": traffic / (tc_ms / 1e3) / 1e9 / hbm_peak,
                "note": "the family is close to balanced: the same launches at the HBM peak would take "
                        f"{traffic / hbm_peak / 1e6:.1f} ms, at the tensor peak {tc_flops / peak_tf / 1e9:.1f} ms"},
            "peak_source": peak_src + " bf16 sustained" if args.dtype == "bf16" else
                           (tf32_peak["how"] if tf32_peak else "nominal fp32 FMA rate"),
            "algorithmic_flops_per_step": tc_flops, "kernel_ms_per_step": tc_ms,
            "kernel_ms_per_step_profiled": tc_ms_profiled, "share_of_step": share,
            "how": "achieved = algorithmic conv FLOPs of the family / (its share of the step from per-launch CUDA events "
                   "x the timed step)",
        }
        breakdown = {"conv_tc_ms": tc_ms_profiled, "conv_simt_ms": sum(r[2] for r in simt), "aux_ms": sum(r[2] for r in aux),
                     "profiled_step_ms": step_ms_profiled}
        # ---- one roofline entry per kernel family: the conv family against the tensor peak, every byte-moving kernel
        # against the measured HBM copy bandwidth (algorithmic bytes of SURVEY 8d / its profiled launch time)
        eb = 2 if args.dtype == "bf16" else 4
        roofline_kernels = [{"kernel": "conv family (conv_tc2_kernel + conv_pair_kernel + conv_chain_kernel)", "bound": "tensor",
                             "launches": len(tc), "ms": tc_ms, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": achieved_tf / peak_tf if peak_tf else None}]
        # the PCM tail is not part of gnv_inference: time it alone (CUDA events, 20 launches back to back)
        for _ in range(3):
            pcm_tail(wav, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm)
        torch.cuda.synchronize(dev)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record(stream)
        for _ in range(20):
            pcm_tail(wav, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm)
        pe1.record(stream)
        torch.cuda.synchronize(dev)
        aux_rows = [(n, ms_) for (n, k, ms_, _) in rows_acc if k == _cabi.LAUNCH_AUX] + [("pcm_tail", pe0.elapsed_time(pe1) / 20)]
        for n, ms_ in aux_rows:
            bpf = aux_bytes_per_frame(n, eb)
            if bpf is None or ms_ <= 0:
                continue
            nbytes = bpf * B * T
            gbps = nbytes / (ms_ / 1e3) / 1e9
            roofline_kernels.append({"kernel": {"istft_head": "istft_kernel", "stft": "stft_kernel", "m_source": "source_kernel",
                                                "pcm_tail": "pcm_tail_kernel", "pack_mel": "nct_to_nlc_kernel",
                                                "f0_predictor.classifier": "f0_head_kernel"}.get(n, n),
                                     "bound": "hbm", "bytes": nbytes, "ms": ms_, "achieved": gbps, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": gbps / hbm_peak})
        if args.profile_table:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_table)), exist_ok=True)
            with open(args.profile_table, "w") as f:
                f.write(f"# per-launch device time, B={B} T={T} dtype={args.dtype}, mean of {n_prof} runs (CUDA events)\n")
                f.write("idx,name,kind,ms,gflop,tflops\n")
                kn = {0: "aux", 1: "conv_tc", 2: "conv_simt"}
                for i, (n, k, ms_, fl) in enumerate(rows_acc):
                    f.write(f"{i},{n},{kn[k]},{ms_:.4f},{fl / 1e9:.2f},{(fl / (ms_ / 1e3) / 1e12) if ms_ > 0 else 0:.1f}\n")

        # ---- first-audio-chunk latency (BASELINE configs[1]): B=1, 100-frame chunk + 16 look-ahead ----
        first_chunk = None
        if not args.no_first_chunk:
            Tc = 116
            mel1 = synthetic_mel(1, Tc, 77).to(dev)
            wav1 = torch.empty(1, Tc * SPF, dtype=torch.float32, device=dev)
            src1 = torch.empty(1, 1, Tc * SPF, dtype=torch.float32, device=dev)
            pcm1 = torch.empty(1, 100 * SPF, dtype=torch.int16, device=dev)
            host1 = torch.empty(1, 100 * SPF, dtype=torch.int16).pin_memory()
            lat = []
            for it in range(220):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                dec.inference(mel1, seed=it + 1, out=wav1, source_out=src1)
                pcm_tail(wav1[:, : 100 * SPF], None, None, 0.99, True, False, out_i16=pcm1)
                host1.copy_(pcm1, non_blocking=True)
                stream.synchronize()
                if it >= 20:
                    lat.append((time.perf_counter() - t0) * 1e3)
            lat.sort()
            # the same chunk through a captured CUDA graph (GraphedInference): no per-launch host cost
            from gonova_tts_b200 import GraphedInference

            gi = GraphedInference(dec, 1, Tc, emit_frames=100, seed=1)
            glat = []
            for it in range(220):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                pcm_g = gi(mel1)
                host1.copy_(pcm_g, non_blocking=True)
                stream.synchronize()
                if it >= 20:
                    glat.append((time.perf_counter() - t0) * 1e3)
            glat.sort()
            first_chunk = {"p50_ms": glat[len(glat) // 2], "p99_ms": glat[int(len(glat) * 0.99) - 1], "iters": len(glat),
                           "eager_p50_ms": lat[len(lat) // 2], "eager_p99_ms": lat[int(len(lat) * 0.99) - 1],
                           "what": "B=1, 100 mel frames + 16 look-ahead, mel on device -> int16 PCM in pinned host "
                                   "memory, wall clock; p50/p99 = CUDA-graph replay (GraphedInference), eager_* = "
                                   "one C-ABI call per kernel launch", "dtype": args.dtype}

        # ---- the reference-precision path: tf32 operands (what the reference's own GPU path computes in,
        # services/tts/core/synthesizer.py:177-179), same workload, against a tf32 GEMM peak measured here
        tf32_block = None
        if world == 1 and args.dtype == "bf16" and not args.no_tf32:
            try:
                dec32 = B200HiFT(random_state_dict(0, False), device=dev, dtype="tf32")

                def step32(i):
                    dec32.inference(mel, seed=i + 1, out=wav, source_out=src)
                    pcm_tail(wav, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm)

                n32 = max(3, args.steps // 2)
                ms32, _ = timed(step32, n32, 3)
                rows32 = None
                for r in range(3):
                    _, rr = dec32.profile_inference(mel, seed=200 + r)
                    if r == 0:
                        continue
                    if rows32 is None:
                        rows32 = [[n, k, 0.0, f] for (n, k, _, f) in rr]
                    for acc, row in zip(rows32, rr):
                        acc[2] += row[2] / 2
                tc32 = [r for r in rows32 if r[1] == _cabi.LAUNCH_CONV_TC]
                share32 = sum(r[2] for r in tc32) / sum(r[2] for r in rows32)
                fl32 = sum(r[3] for r in tc32)
                fam_ms32 = share32 * ms32 / n32
                pk = measure_tf32_gemm_peak(dev)
                ach32 = fl32 / (fam_ms32 / 1e3) / 1e12
                tf32_block = {"value": audio_s_per_gpu * n32 / (ms32 / 1e3), "unit": UNIT, "ms_per_step": ms32 / n32, "steps": n32,
                              "roofline": {"bound": "tensor", "achieved": ach32, "peak": pk["tf32_tflops_sustained"],
                                           "peak_burst": pk["tf32_tflops"], "unit": "TFLOP/s",
                                           "frac": ach32 / pk["tf32_tflops_sustained"], "share_of_step": share32,
                                           "peak_source": pk["how"]}}
                del dec32
                torch.cuda.empty_cache()
            except Exception as e:      # the headline must not die with an optional block
                tf32_block = {"error": str(e)}

        # ---- second comparator: stock PyTorch on this same GPU — the reference's modules as it runs them (cuDNN
        # autotune + TF32, synthesizer.py:175-179).  A BASELINE leg like cpu_baseline: it runs the oracle's restatement of the
        # engine's nn.Modules moved to cuda; nothing of it is on the product path.
        stock_block = None
        if world == 1 and not args.no_stock_torch:
            try:
                from oracle import hift_ref as R

                old_flags = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
                torch.backends.cudnn.benchmark = True
                torch.backends.cuda.matmul.allow_tf32 = True
                torch.backends.cudnn.allow_tf32 = True
                m = R.load_model(random_state_dict(0, False)).to(dev).eval()
                gen = torch.Generator(device="cpu").manual_seed(6)
                s_in = (torch.rand(B, 1, L, generator=gen) * 0.2 - 0.1).to(dev)

                def stock_step():
                    with torch.inference_mode():
                        m.decode(mel, s_in)

                for _ in range(2):
                    stock_step()
                torch.cuda.synchronize(dev)
                se0, se1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                se0.record(stream)
                for _ in range(3):
                    stock_step()
                se1.record(stream)
                torch.cuda.synchronize(dev)
                ms_stock = se0.elapsed_time(se1) / 3
                de0, de1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(2):
                    dec.decode(mel, s_in, out=wav)
                de0.record(stream)
                for _ in range(5):
                    dec.decode(mel, s_in, out=wav)
                de1.record(stream)
                torch.cuda.synchronize(dev)
                ms_ours = de0.elapsed_time(de1) / 5
                stock_block = {"value": audio_s_per_gpu / (ms_stock / 1e3), "unit": UNIT, "ms_per_decode": ms_stock,
                               "ours_ms_per_decode": ms_ours, "ours_dtype": args.dtype, "speedup": ms_stock / ms_ours,
                               "what": f"decode(x, s) of the same {B} x {T}-frame batch, inputs resident: the engine's nn.Modules "
                                       "(oracle restatement) on cuda:0 with cudnn.benchmark + TF32 vs this repo's decode"}
                torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_flags
                del m, s_in
                torch.cuda.empty_cache()
            except Exception as e:
                stock_block = {"error": str(e)}

        # ---- the step before the vocoder: CFM flow decoder, ten Euler steps with classifier-free guidance
        flow_block = None
        if world == 1 and not args.no_flow:
            try:
                from gonova_tts_b200 import B200Flow

                fg = torch.Generator().manual_seed(0)
                # seeded random weights of the estimator's architecture (names / shapes: gonova_tts_b200/flow.py)
                from gonova_tts_b200.flow import random_flow_state_dict

                flow = B200Flow(random_flow_state_dict(0), device=dev, dtype=args.dtype if args.dtype in ("bf16", "tf32") else "bf16")
                fB, fT = 32, T
                fz = torch.randn(fB, 80, fT, generator=fg).to(dev)
                fmu = (torch.randn(fB, 80, fT, generator=fg) * 0.5 - 1.0).to(dev)
                fsp = torch.nn.functional.normalize(torch.randn(fB, 80, generator=fg), dim=1).to(dev)
                fcond = torch.zeros(fB, 80, fT, device=dev)
                for _ in range(2):
                    flow.decode(fz, fmu, fsp, fcond)
                torch.cuda.synchronize(dev)
                fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fe0.record(stream)
                for _ in range(3):
                    flow.decode(fz, fmu, fsp, fcond)
                fe1.record(stream)
                torch.cuda.synchronize(dev)
                fms = fe0.elapsed_time(fe1) / 3
                gemm_flops = 2.0 * 65.1e6 * 2 * fB * fT * 10 + 10 * 56 * 2 * fB * 8 * 4.0 * fT * fT * 64
                # per-kernel-class device times of ONE estimator evaluation (gnv_flow_profile: an event after every launch)
                import collections
                import re as _re
                flow.profile(fz, fmu, fsp, fcond, n_timesteps=1)
                acc = collections.OrderedDict()
                n_fp = 3
                for _ in range(n_fp):
                    for name, kind, ms_i, fl_i in flow.profile(fz, fmu, fsp, fcond, n_timesteps=1):
                        cls = _re.sub(r"^(down_blocks\.0|mid_blocks\.\d+|up_blocks\.0)\.", "L.", name)
                        cls = _re.sub(r"^L\.1\.\d+\.", "L.1.j.", cls)
                        a_ = acc.setdefault(cls, [0, 0.0, 0.0])
                        a_[0] += 1; a_[1] += ms_i / n_fp; a_[2] += fl_i / n_fp
                eval_ms = sum(v[1] for v in acc.values())
                f_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) if flow.dtype == "bf16" else None
                fk = []
                for cls, (n_l, ms_c, fl_c) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
                    ent = {"kernel": cls, "launches": n_l // n_fp, "ms": ms_c, "share": ms_c / eval_ms if eval_ms else 0.0}
                    if fl_c > 0:
                        ent.update({"bound": "tensor", "achieved": fl_c / (ms_c / 1e3) / 1e12, "unit": "TFLOP/s"})
                        if f_peak:
                            ent.update({"peak": f_peak, "frac": ent["achieved"] / f_peak})
                    fk.append(ent)
                tens_ms = sum(v[1] for v in acc.values() if v[2] > 0)
                tens_fl = sum(v[2] for v in acc.values())
                # end to end: pinned host inputs -> device, decode, mel back to pinned host memory, inside the timed region
                hz, hmu, hcond, hsp = (t.cpu().pin_memory() for t in (fz, fmu, fcond, fsp))
                hmel = torch.empty(fB, 80, fT, dtype=torch.float32).pin_memory()
                dz, dmu, dcond, dsp = (torch.empty_like(t) for t in (fz, fmu, fcond, fsp))

                def flow_e2e():
                    dz.copy_(hz, non_blocking=True); dmu.copy_(hmu, non_blocking=True)
                    dcond.copy_(hcond, non_blocking=True); dsp.copy_(hsp, non_blocking=True)
                    hmel.copy_(flow.decode(dz, dmu, dsp, dcond), non_blocking=True)

                flow_e2e()
                torch.cuda.synchronize(dev)
                fe0.record(stream)
                for _ in range(3):
                    flow_e2e()
                fe1.record(stream)
                torch.cuda.synchronize(dev)
                fms_e2e = fe0.elapsed_time(fe1) / 3
                # one utterance alone (the service's case: one sentence per request)
                z1, mu1, sp1, c1 = fz[:1].contiguous(), fmu[:1].contiguous(), fsp[:1].contiguous(), fcond[:1].contiguous()
                for _ in range(2):
                    flow.decode(z1, mu1, sp1, c1)
                torch.cuda.synchronize(dev)
                fe0.record(stream)
                for _ in range(5):
                    flow.decode(z1, mu1, sp1, c1)
                fe1.record(stream)
                torch.cuda.synchronize(dev)
                fms_b1 = fe0.elapsed_time(fe1) / 5
                # other batch sizes: 8 utterances, and 37 (2 x 37 x 500 rows = 145 pair tiles of 256 rows on 74 CTA pairs: two full
                # waves, where 32 utterances leave a third of the second wave empty)
                sweep = []
                for sb in (8, 37):
                    sz = torch.randn(sb, 80, fT, generator=fg).to(dev)
                    smu = (torch.randn(sb, 80, fT, generator=fg) * 0.5 - 1.0).to(dev)
                    ssp = torch.nn.functional.normalize(torch.randn(sb, 80, generator=fg), dim=1).to(dev)
                    scond = torch.zeros(sb, 80, fT, device=dev)
                    for _ in range(2):
                        flow.decode(sz, smu, ssp, scond)
                    torch.cuda.synchronize(dev)
                    fe0.record(stream)
                    for _ in range(3):
                        flow.decode(sz, smu, ssp, scond)
                    fe1.record(stream)
                    torch.cuda.synchronize(dev)
                    sms = fe0.elapsed_time(fe1) / 3
                    sweep.append({"batch": sb, "ms_per_decode": sms, "value": sb * fT / 50.0 / (sms / 1e3), "unit": UNIT})
                    del sz, smu, ssp, scond
                # the reference's arithmetic class (tf32 operands, one tensor-core launch per projection, mma.sync tf32 attention)
                flow_tf32 = None
                if flow.dtype == "bf16" and not args.no_tf32:
                    try:
                        f32 = B200Flow(random_flow_state_dict(0), device=dev, dtype="tf32")
                        f32.decode(fz, fmu, fsp, fcond)
                        torch.cuda.synchronize(dev)
                        fe0.record(stream)
                        f32.decode(fz, fmu, fsp, fcond)
                        fe1.record(stream)
                        torch.cuda.synchronize(dev)
                        tms = fe0.elapsed_time(fe1)
                        flow_tf32 = {"value": fB * fT / 50.0 / (tms / 1e3), "unit": UNIT, "ms_per_decode": tms, "batch": fB}
                        del f32
                        torch.cuda.empty_cache()
                    except Exception as te:
                        flow_tf32 = {"error": str(te)}
                # the only pre-existing Blackwell path for this step: stock PyTorch (the engine's nn.Modules, here their oracle
                # restatement) on the same GPU with cudnn.benchmark + TF32, as synthesizer.py:175-179 configures it
                flow_stock = None
                if not args.no_stock_torch:
                    try:
                        from oracle import flow_ref as FRs
                        old_flags = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
                        torch.backends.cudnn.benchmark = True
                        torch.backends.cuda.matmul.allow_tf32 = True
                        torch.backends.cudnn.allow_tf32 = True
                        est_t = FRs.load_estimator(random_flow_state_dict(0)).to(dev).eval()
                        ones = torch.ones(fB, 1, fT, device=dev)
                        with torch.inference_mode():
                            FRs.solve_euler(est_t, fz, fmu, ones, fsp, fcond, n_timesteps=2)
                            torch.cuda.synchronize(dev)
                            fe0.record(stream)
                            FRs.solve_euler(est_t, fz, fmu, ones, fsp, fcond, n_timesteps=10)
                            fe1.record(stream)
                            torch.cuda.synchronize(dev)
                        sms = fe0.elapsed_time(fe1)
                        flow_stock = {"value": fB * fT / 50.0 / (sms / 1e3), "unit": UNIT, "ms_per_decode": sms, "batch": fB,
                                      "speedup": sms / fms, "ours_dtype": flow.dtype,
                                      "what": "solve_euler of the same batch (ten steps, CFG) on the estimator's nn.Modules (oracle "
                                              "restatement) on cuda:0, eager, cudnn.benchmark + TF32, inputs resident"}
                        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_flags
                        del est_t
                        torch.cuda.empty_cache()
                    except Exception as se:
                        flow_stock = {"error": str(se)}
                # the two drop-ins chained, as S3Gen.inference chains them: flow decoder -> mel -> vocoder -> int16 PCM in
                # pinned host memory (mu / spks / cond resident: the encoder in front stays the engine's)
                chain = None
                try:
                    pcm_host = torch.empty(fB, fT * 480, dtype=torch.int16).pin_memory()
                    pcm_dev = torch.empty(fB, fT * 480, dtype=torch.int16, device=dev)
                    no_cache = torch.zeros(fB, 1, 0, device=dev)

                    def flow_then_vocoder():
                        mel_f = flow.decode(fz, fmu, fsp, fcond)
                        wav_f, _ = dec.inference(mel_f, cache_source=no_cache)
                        pcm_tail(wav_f, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm_dev)
                        pcm_host.copy_(pcm_dev, non_blocking=True)

                    flow_then_vocoder()
                    torch.cuda.synchronize(dev)
                    fe0.record(stream)
                    for _ in range(3):
                        flow_then_vocoder()
                    fe1.record(stream)
                    torch.cuda.synchronize(dev)
                    cms = fe0.elapsed_time(fe1) / 3
                    chain = {"value": fB * fT / 50.0 / (cms / 1e3), "unit": UNIT, "ms_per_batch": cms, "batch": fB,
                             "what": "mu -> flow decoder (10 Euler steps, CFG) -> mel -> f0 / source / HiFT decode -> int16 PCM in "
                                     "pinned host memory, both on this repo's kernels"}
                except Exception as ce:
                    chain = {"error": str(ce)}
                # the front of the flow step (speech tokens -> mu / spks: embedding + upsampling Conformer encoder), and the whole
                # tokens -> PCM chain on this repo's kernels: pinned int32 tokens + x-vectors H2D, encode, ten Euler steps, vocoder,
                # int16 PCM back to pinned host memory
                front_block = None
                try:
                    from gonova_tts_b200.flow_front import B200FlowFront, random_front_state_dict

                    front = B200FlowFront(random_front_state_dict(0), device=dev, dtype=flow.dtype)
                    tL = fT // 2
                    htok = torch.randint(0, 6561, (fB, tL), generator=fg, dtype=torch.int32).pin_memory()
                    hemb = torch.randn(fB, 192, generator=fg).pin_memory()
                    dtok, demb = htok.to(dev), hemb.to(dev)
                    for _ in range(2):
                        front.encode(dtok, None, demb)
                    torch.cuda.synchronize(dev)
                    fe0.record(stream)
                    for _ in range(5):
                        front.encode(dtok, None, demb)
                    fe1.record(stream)
                    torch.cuda.synchronize(dev)
                    enc_ms = fe0.elapsed_time(fe1) / 5
                    prof_rows = front.profile(dtok, None, demb)
                    prof_rows = front.profile(dtok, None, demb)
                    t_ms = sum(r[2] for r in prof_rows if r[3] > 0 and r[0] != "self_attn")
                    t_fl = sum(r[3] for r in prof_rows if r[3] > 0 and r[0] != "self_attn")
                    a_ms = sum(r[2] for r in prof_rows if r[0] == "self_attn")
                    a_fl = sum(r[3] for r in prof_rows if r[0] == "self_attn")
                    pcm_host2 = torch.empty(fB, fT * 480, dtype=torch.int16).pin_memory()
                    pcm_dev2 = torch.empty(fB, fT * 480, dtype=torch.int16, device=dev)
                    no_cache2 = torch.zeros(fB, 1, 0, device=dev)

                    def tokens_to_pcm():
                        dtok.copy_(htok, non_blocking=True); demb.copy_(hemb, non_blocking=True)
                        mu_t, sp_t = front.encode(dtok, None, demb)
                        mel_t = flow.decode(fz, mu_t, sp_t, fcond)
                        wav_t, _ = dec.inference(mel_t, cache_source=no_cache2)
                        pcm_tail(wav_t, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm_dev2)
                        pcm_host2.copy_(pcm_dev2, non_blocking=True)

                    tokens_to_pcm()
                    torch.cuda.synchronize(dev)
                    fe0.record(stream)
                    for _ in range(3):
                        tokens_to_pcm()
                    fe1.record(stream)
                    torch.cuda.synchronize(dev)
                    t2p_ms = fe0.elapsed_time(fe1) / 3
                    # one sentence alone (the service's case): tokens -> int16 PCM in pinned host memory
                    tok1, emb1, nc1 = dtok[:1].contiguous(), demb[:1].contiguous(), torch.zeros(1, 1, 0, device=dev)
                    pcm1_host = torch.empty(1, fT * 480, dtype=torch.int16).pin_memory()
                    pcm1_dev = torch.empty(1, fT * 480, dtype=torch.int16, device=dev)

                    def one_sentence():
                        mu_1, sp_1 = front.encode(tok1, None, emb1)
                        mel_1 = flow.decode(z1, mu_1, sp_1, c1)
                        wav_1, _ = dec.inference(mel_1, cache_source=nc1)
                        pcm_tail(wav_1, None, None, 0.99, want_i16=True, want_f32=False, out_i16=pcm1_dev)
                        pcm1_host.copy_(pcm1_dev, non_blocking=True)

                    for _ in range(2):
                        one_sentence()
                    torch.cuda.synchronize(dev)
                    fe0.record(stream)
                    for _ in range(5):
                        one_sentence()
                    fe1.record(stream)
                    torch.cuda.synchronize(dev)
                    t2p1_ms = fe0.elapsed_time(fe1) / 5
                    front_cpu = None
                    if not args.no_cpu_baseline:
                        from oracle import flow_enc_ref as ER
                        torch.set_num_threads(os.cpu_count() or 1)
                        ofront = ER.load_front(ER.random_state_dict(0))
                        ctok, clen, _ = ER.synthetic_tokens(1, 250, seed=1)
                        with torch.inference_mode():
                            ofront.encode(ctok[:, :50])
                            t0c = time.perf_counter()
                            ofront.encode(ctok, clen)
                            dtc = time.perf_counter() - t0c
                        front_cpu = {"value": 10.0 / dtc, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                     "sample": f"1 utterance x 250 tokens (10 audio-s) in {dtc:.2f} s, fp32 torch CPU oracle "
                                               "(oracle/flow_enc_ref.py)"}
                    front_block = {"value": fB * fT / 50.0 / (enc_ms / 1e3), "unit": UNIT, "ms_per_encode": enc_ms, "batch": fB,
                                   "tokens": tL, "frames": fT, "dtype": front.dtype, "gpu_launches": front.launches(),
                                   "roofline_kernels": [
                                       {"kernel": "conv_tc2_kernel (the encoder's 46 Linear / Conv1d launches)", "bound": "tensor",
                                        "ms": t_ms, "achieved": t_fl / (t_ms / 1e3) / 1e12 if t_ms else 0.0, "unit": "TFLOP/s",
                                        "peak": f_peak, "frac": (t_fl / (t_ms / 1e3) / 1e12 / f_peak) if (f_peak and t_ms) else None},
                                       {"kernel": "enc_attn_mma_kernel (relative-position attention, mma.sync; 6 B H T^2 d flops)"
                                                  if front.dtype == "bf16" else "enc_attn_kernel (relative-position attention, fp32 CUDA cores)",
                                        "bound": "tensor", "ms": a_ms, "achieved": a_fl / (a_ms / 1e3) / 1e12 if a_ms else 0.0,
                                        "unit": "TFLOP/s", "peak": f_peak,
                                        "frac": (a_fl / (a_ms / 1e3) / 1e12 / f_peak) if (f_peak and a_ms) else None}],
                                   "cpu_baseline": front_cpu,
                                   "tokens_to_pcm": {"value": fB * fT / 50.0 / (t2p_ms / 1e3), "unit": UNIT, "ms_per_batch": t2p_ms,
                                                     "batch": fB, "h2d_bytes_per_step": int(htok.numel() * 4 + hemb.numel() * 4),
                                                     "d2h_bytes_per_step": int(pcm_host2.numel() * 2),
                                                     "what": "pinned speech tokens + x-vectors -> embedding + Conformer encoder -> ten "
                                                             "Euler steps of the CFM decoder -> f0 / source / HiFT decode -> int16 PCM in "
                                                             "pinned host memory: S3Gen.inference without its prompt bookkeeping, every "
                                                             "kernel this repo's"},
                                   "tokens_to_pcm_one_sentence": {"ms": t2p1_ms, "value": fT / 50.0 / (t2p1_ms / 1e3), "unit": UNIT,
                                                                  "what": "the same chain for ONE 10 s sentence (250 tokens): the latency "
                                                                          "a single request sees between its last token and its PCM"},
                                   "what": "gnv_flow_encode: tokens [B, L] + x-vectors -> mu [B, 80, 2L], spks [B, 80]"}
                    del front
                except Exception as fe:
                    front_block = {"error": str(fe)}
                flow_cpu = None
                if not args.no_cpu_baseline:
                    # the oracle (fp32 torch CPU restatement) on all host cores, bounded sample: 1 utterance x 100 frames, 2 Euler steps
                    from oracle import flow_ref as FR
                    torch.set_num_threads(os.cpu_count() or 1)
                    est = FR.load_estimator(FR.random_state_dict(0))
                    cz, cmu, cmask, csp, ccond = FR.synthetic_inputs(1, 100, seed=1)
                    with torch.inference_mode():
                        FR.solve_euler(est, cz, cmu, cmask, csp, ccond, n_timesteps=1)
                        t0c = time.perf_counter()
                        FR.solve_euler(est, cz, cmu, cmask, csp, ccond, n_timesteps=2)
                        dtc = time.perf_counter() - t0c
                    flow_cpu = {"value": 100 / 50.0 / (dtc * 5.0), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"1 utterance x 100 frames, 2 of the 10 Euler steps in {dtc:.2f} s (x5 for the full "
                                          "solve), fp32 torch CPU oracle (oracle/flow_ref.py)"}
                # DRAM bytes of one estimator evaluation from the committed ncu launch list (valid for this default shape)
                flow_traffic, flow_traffic_src = None, None
                try:
                    with open(os.path.join(ROOT, "profiles", "r02_flow_traffic.json")) as tf_:
                        tjf = json.load(tf_)
                    if (fB, fT, flow.dtype) == (32, 500, "bf16"):
                        flow_traffic, flow_traffic_src = tjf["traffic_bytes_per_evaluation"], tjf["source"]
                except Exception:
                    pass
                flow_block = {"value": fB * fT / 50.0 / (fms / 1e3), "unit": UNIT, "ms_per_decode": fms, "batch": fB, "frames": fT,
                              "dtype": flow.dtype, "n_timesteps": 10, "cfg_rate": 0.7, "gpu_launches": flow.launches(10),
                              "algorithmic_tflops": gemm_flops / (fms / 1e3) / 1e12,
                              "e2e": {"value": fB * fT / 50.0 / (fms_e2e / 1e3), "unit": UNIT, "ms_per_decode": fms_e2e,
                                      "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in (hz, hmu, hcond, hsp))),
                                      "d2h_bytes_per_step": int(hmel.numel() * 4)},
                              "one_utterance": {"ms_per_decode": fms_b1, "value": fT / 50.0 / (fms_b1 / 1e3), "unit": UNIT},
                              "estimator_eval_ms_profiled": eval_ms,
                              "roofline": {"kernel": "flow_blk_kernel + flow_attn_tc_kernel + conv_tc2_kernel (every tensor-core "
                                                     "launch of one estimator evaluation)", "bound": "tensor",
                                           "achieved": tens_fl / (tens_ms / 1e3) / 1e12 if tens_ms else 0.0, "peak": f_peak,
                                           "unit": "TFLOP/s",
                                           "frac": (tens_fl / (tens_ms / 1e3) / 1e12 / f_peak) if (f_peak and tens_ms) else None,
                                           "traffic": flow_traffic, "traffic_source": flow_traffic_src,
                                           "how": "algorithmic FLOPs of the projections, convs and attention (4 B2 H T^2 d) / summed "
                                                  "per-launch CUDA-event time of those launches in one evaluation"},
                              "roofline_kernels": fk, "cpu_baseline": flow_cpu, "stock_torch_gpu": flow_stock, "flow_then_vocoder": chain, "front": front_block, "batch_sweep": sweep, "tf32": flow_tf32,
                              "what": "gnv_flow_decode: mu / spks / cond resident -> mel, ten Euler steps x doubled batch "
                                      "(classifier-free guidance); bf16: between two attention launches (flow_attn_tc_kernel, tcgen05) a "
                                      "transformer block is ONE flow_blk_kernel launch (out-proj + residual + LayerNorm + feed-forward + "
                                      "residual + LayerNorm + the next block's q/k/v); ResNet blocks: conv + LayerNorm + Mish fused"}
                del flow
                torch.cuda.empty_cache()
            except Exception as e:
                flow_block = {"error": str(e)}

        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            v, sec, cores, times = cpu_decode_rate(2, T, reps=3)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"2 utterances x {T} frames (20 audio-s) per call, best of 3 after 1 warm-up "
                                      f"({sec:.2f} s per call), fp32 torch CPU oracle"}
        launches_per_step = dec.launches(B, T, inference=True) + 1
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"{B} concurrent {T / 50:.0f} s utterances per GPU through the decoder "
                            "(BASELINE configs[2]; N=8 is configs[3]'s 512 streams): f0 -> source -> HiFT decode -> int16",
                "utterances_per_gpu": B, "frames_per_utterance": T, "sample_rate": SR,
                "audio_seconds_per_step": world * audio_s_per_gpu, "weights": "seeded random init (no checkpoint)",
                "l2": f"inputs larger than L2: {ws_bytes / 2**30:.1f} GiB of activations per step vs 126 MB L2",
                "parallelism": f"dp{world} (request sharding, no collective)",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": world * mel_host.numel() * 4, "d2h_bytes_per_step": world * pcm_host.numel() * 2,
                    "how": "pinned mel H2D and int16 PCM D2H every step on a copy stream, double-buffered (overlapping the "
                           "neighbouring steps' kernels); the host waits for step i-1's PCM while step i runs"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "roofline_kernels": roofline_kernels, "breakdown": breakdown,
            "first_chunk": first_chunk, "cpu_baseline": cpu_baseline, "tf32": tf32_block, "stock_torch_gpu": stock_block,
            "streams512": streams512, "flow_decoder": flow_block,
        }
    barrier()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
