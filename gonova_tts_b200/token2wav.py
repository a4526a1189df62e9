"""B200Token2Wav — speech tokens -> waveform, the engine's `s3gen.inference(...)` call on this library end to end:
B200FlowInference (token embedding + Conformer encoder + ten Euler steps of the CFM decoder) -> B200HiFT (f0, NSF source,
HiFT decode) -> the 960-sample `trim_fade`.

Reference boundary: services/tts/core/synthesizer.py:344-350 `model.generate(...)` -> upstream `S3Token2Wav.inference(
speech_tokens, ref_wav=None, ref_sr=None, ref_dict=..., cache_source=None, finalize=True)` = `flow_inference` +
`hift_inference` + `output_wavs[:, :len(trim_fade)] *= trim_fade`.  The reference clip's side (`embed_ref`: the S3 tokenizer,
the speaker encoder and the mel extractor that produce `ref_dict`) stays the engine's: it runs once per voice, not per
sentence (the service caches it per voice, synthesizer.py:238-262)."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .decoder import B200HiFT, trim_fade_window
from .flow_front import B200FlowInference

REF_KEYS = ("prompt_token", "prompt_token_len", "prompt_feat", "prompt_feat_len", "embedding")


class B200Token2Wav(torch.nn.Module):
    def __init__(self, flow: B200FlowInference, mel2wav: B200HiFT):
        super().__init__()
        if flow.device != mel2wav.device:
            raise ValueError(f"the flow is on {flow.device}, the vocoder on {mel2wav.device}")
        self.flow = flow
        self.mel2wav = mel2wav
        self.device = flow.device
        self.register_buffer("trim_fade", trim_fade_window(self.device), persistent=False)

    @classmethod
    def from_state_dict(cls, state_dict: Dict[str, torch.Tensor], device="cuda:0", dtype: str = "bf16", noise_seed: int = 0):
        """`state_dict` = upstream's S3Token2Wav weights: `flow.*` (front + `flow.decoder.estimator.*`) and `mel2wav.*`."""
        flow = B200FlowInference({k[5:]: v for k, v in state_dict.items() if k.startswith("flow.")}, device=device, dtype=dtype,
                                 noise_seed=noise_seed)
        hift = B200HiFT({k[8:]: v for k, v in state_dict.items() if k.startswith("mel2wav.")}, device=device, dtype=dtype)
        return cls(flow, hift)

    @torch.no_grad()
    def flow_inference(self, speech_tokens: torch.Tensor, ref_dict: Dict[str, torch.Tensor], finalize: bool = True):
        if speech_tokens.dim() == 1:
            speech_tokens = speech_tokens.unsqueeze(0)
        missing = [k for k in ("prompt_token", "prompt_feat", "embedding") if k not in ref_dict]
        if missing:
            raise ValueError(f"ref_dict lacks {missing} (it is what the engine's embed_ref() returns)")
        n = torch.tensor([speech_tokens.shape[1]])
        mels, _ = self.flow.inference(token=speech_tokens, token_len=n, finalize=finalize,
                                      **{k: ref_dict.get(k) for k in REF_KEYS})
        return mels

    @torch.no_grad()
    def hift_inference(self, speech_feat: torch.Tensor, cache_source: Optional[torch.Tensor] = None):
        if cache_source is None:
            cache_source = torch.zeros(1, 1, 0, device=self.device)
        return self.mel2wav.inference(speech_feat=speech_feat, cache_source=cache_source)

    @torch.no_grad()
    def inference(self, speech_tokens: torch.Tensor, ref_wav=None, ref_sr=None, ref_dict: Optional[Dict[str, torch.Tensor]] = None,
                  cache_source: Optional[torch.Tensor] = None, finalize: bool = True):
        """-> (wav [1, 480 * frames of the new tokens], source [1, 1, same])."""
        if ref_dict is None:
            raise ValueError("give ref_dict (the engine's embed_ref(ref_wav, ref_sr)): the reference clip's tokenizer / speaker "
                             "encoder are not part of this library")
        mels = self.flow_inference(speech_tokens, ref_dict, finalize=finalize)
        wavs, sources = self.hift_inference(mels, cache_source)
        n = min(self.trim_fade.shape[0], wavs.shape[1])
        wavs[:, :n] *= self.trim_fade[:n]
        return wavs, sources

    forward = inference
