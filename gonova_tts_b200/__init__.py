"""gonova_tts_b200 — B200-native waveform decoder (HiFT vocoder) for the gonova-tts TTS service.

Host side: Python/PyTorch (device memory + streams).  Compute: libgonova_hift.so, hand-written
sm_100a CUDA (tcgen05/TMEM/TMA conv GEMMs, fused STFT / iSTFT+OLA / PCM-tail kernels) behind the C
ABI in include/gonova_hift.h.  No CPU fallback: importing works anywhere, running needs the built
library and a B200."""
from .decoder import B200HiFT, SAMPLES_PER_FRAME, fade_window, mulaw_encode, pcm_tail, trim_fade_window  # noqa: F401
from .streaming import GraphedInference, IncrementalDecoder, StreamingDecoder, chunk_plan  # noqa: F401
from .dispatch import ShardedDecoder, round_robin, shard_range  # noqa: F401
from .batching import MicroBatcher, RequestBatcher  # noqa: F401
from .flow import B200Flow  # noqa: F401
from .flow_front import B200FlowFront, B200FlowInference  # noqa: F401
from .token2wav import B200Token2Wav  # noqa: F401
from .weights import fold_state_dict, random_state_dict  # noqa: F401

__all__ = [
    "B200HiFT", "StreamingDecoder", "IncrementalDecoder", "GraphedInference", "MicroBatcher", "RequestBatcher", "ShardedDecoder", "B200Flow", "B200FlowFront", "B200FlowInference", "B200Token2Wav", "pcm_tail", "mulaw_encode", "fade_window", "trim_fade_window",
    "chunk_plan", "shard_range", "round_robin", "fold_state_dict", "random_state_dict", "SAMPLES_PER_FRAME",
]
