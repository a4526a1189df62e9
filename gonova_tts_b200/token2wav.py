"""B200Token2Wav — speech tokens -> waveform, the engine's `s3gen.inference(...)` call on this library end to end:
B200FlowInference (token embedding + Conformer encoder + ten Euler steps of the CFM decoder) -> B200HiFT (f0, NSF source,
HiFT decode) -> the 960-sample `trim_fade`.

Reference boundary: services/tts/core/synthesizer.py:344-350 `model.generate(...)` -> upstream `S3Token2Wav.inference(
speech_tokens, ref_wav=None, ref_sr=None, ref_dict=..., cache_source=None, finalize=True)` = `flow_inference` +
`hift_inference` + `output_wavs[:, :len(trim_fade)] *= trim_fade`.  The reference clip's side (`embed_ref`: the S3 tokenizer,
the speaker encoder and the mel extractor that produce `ref_dict`) stays the engine's: it runs once per voice, not per
sentence (the service caches it per voice, synthesizer.py:238-262)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .decoder import B200HiFT, trim_fade_window
from .flow import CFG_RATE, N_TIMESTEPS
from .flow_front import MEL, TOKEN_MEL_RATIO, B200FlowInference

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

    # ---- many requests at once (SURVEY 8f-4 for the whole tokens -> PCM path; upstream runs one utterance per call) ----
    @staticmethod
    def _rows(n: int, pad_pow2: bool) -> int:
        return 1 << (n - 1).bit_length() if pad_pow2 and n > 1 else n

    @torch.no_grad()
    def flow_inference_batch(self, requests: Sequence[Tuple[torch.Tensor, Dict[str, torch.Tensor]]],
                             n_timesteps: int = N_TIMESTEPS, pad_tokens: int = 16, pad_batch_pow2: bool = True) -> List[torch.Tensor]:
        """[(speech_tokens [n_i] or [1, n_i], ref_dict_i), ...] -> [mel_i [1, 80, frames of request i's new tokens], ...].
        ONE ragged batch through the encoder and the ten Euler steps: every utterance is encoded and decoded exactly as if it
        were alone (lengths mask the rest: tests hold the rows against `flow_inference` of each request).
        pad_tokens / pad_batch_pow2: the batch is padded to a multiple of `pad_tokens` tokens and to a power-of-two number of
        rows (zero-length rows, skipped by every kernel), so that a service's batches fall on few distinct (rows, length)
        shapes and find their launch plans built."""
        dev = self.device
        B = len(requests)
        if B == 0:
            return []
        toks, embs, feats = [], [], []
        for tokens, ref in requests:
            t = tokens.reshape(-1).to(dev, torch.int32)
            toks.append(torch.cat([ref["prompt_token"].reshape(-1).to(dev, torch.int32), t]))
            embs.append(ref["embedding"].reshape(1, -1).to(dev, torch.float32))
            feats.append(ref["prompt_feat"].to(dev, torch.float32).reshape(-1, MEL))
        lens = [int(t.numel()) for t in toks]
        L = -(-max(lens) // max(1, pad_tokens)) * max(1, pad_tokens)
        rows = self._rows(B, pad_batch_pow2)
        tokens = torch.zeros(rows, L, dtype=torch.int32, device=dev)
        for b, t in enumerate(toks):
            tokens[b, : lens[b]] = t
        lens_all = lens + [0] * (rows - B)
        token_len = torch.tensor(lens_all, dtype=torch.int32, device=dev)
        emb = torch.zeros(rows, embs[0].shape[1], dtype=torch.float32, device=dev)
        emb[:B] = torch.cat(embs, dim=0)
        mu, spks = self.flow.front.encode(tokens, token_len, emb)
        T = TOKEN_MEL_RATIO * L
        cond = torch.zeros(rows, MEL, T, dtype=torch.float32, device=dev)
        for b, f in enumerate(feats):
            cond[b, :, : f.shape[0]] = f.transpose(0, 1)
        dec = self.flow.decoder
        if T > dec.rand_noise.shape[2]:
            raise ValueError("utterance longer than the noise buffer")
        z = dec.rand_noise[:, :, :T].expand(rows, -1, -1).contiguous()  # every utterance starts from the buffer's first frames
        mel = dec.decode(z, mu, spks, cond, lengths=[TOKEN_MEL_RATIO * n for n in lens_all], n_timesteps=n_timesteps,
                         cfg_rate=CFG_RATE)
        return [mel[b : b + 1, :, feats[b].shape[0] : TOKEN_MEL_RATIO * lens[b]] for b in range(B)]

    @torch.no_grad()
    def inference_batch(self, requests: Sequence[Tuple[torch.Tensor, Dict[str, torch.Tensor]]],
                        n_timesteps: int = N_TIMESTEPS, pad_tokens: int = 16, pad_frames: int = 32,
                        pad_batch_pow2: bool = True) -> List[torch.Tensor]:
        """The same requests -> [wav_i [1, 480 * frames_i], ...]: one ragged flow batch, then one ragged vocoder batch, then each
        utterance's own trim_fade."""
        mels = self.flow_inference_batch(requests, n_timesteps=n_timesteps, pad_tokens=pad_tokens, pad_batch_pow2=pad_batch_pow2)
        if not mels:
            return []
        frames = [int(m.shape[2]) for m in mels]
        Tm = -(-max(frames) // max(1, pad_frames)) * max(1, pad_frames)
        rows = self._rows(len(mels), pad_batch_pow2)
        batch = torch.zeros(rows, MEL, Tm, dtype=torch.float32, device=self.device)
        for b, m in enumerate(mels):
            batch[b, :, : frames[b]] = m[0]
        wavs, _ = self.mel2wav.inference(speech_feat=batch, lengths=frames + [0] * (rows - len(mels)))
        out = []
        for b, n in enumerate(frames):
            w = wavs[b : b + 1, : 480 * n].clone()
            k = min(self.trim_fade.shape[0], w.shape[1])
            w[:, :k] *= self.trim_fade[:k]
            out.append(w)
        return out
