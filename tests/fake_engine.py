"""A stand-in for the engine the service wraps (`chatterbox.tts.ChatterboxTTS`), with the call shape the reference uses
(services/tts/core/synthesizer.py:185, :344-350) and the upstream object layout: a plain engine object holding
`s3gen`, an nn.Module whose registered child `mel2wav` is the HiFT vocoder (here: the oracle's HiFTGenerator).  The
"front end" (T3 + CFM flow, out of scope) is replaced by a deterministic synthetic mel whose length depends on the
text.  Test infrastructure only."""
import torch
from torch import nn

from oracle import hift_ref as R


class FakeS3Gen(nn.Module):
    """Upstream S3Token2Wav's tail: `self.mel2wav.inference(speech_feat=..., cache_source=...)`, then
    `wav[:, :960] *= trim_fade`."""

    def __init__(self, mel2wav: nn.Module):
        super().__init__()
        self.mel2wav = mel2wav
        self.register_buffer("trim_fade", R.trim_fade_window(), persistent=False)

    @torch.inference_mode()
    def inference(self, mel: torch.Tensor):
        cache = torch.zeros(1, 1, 0).to(mel.device)
        wav, src = self.mel2wav.inference(speech_feat=mel, cache_source=cache)
        wav = wav.clone()
        n = min(960, wav.shape[1])
        wav[:, :n] *= self.trim_fade[:n]
        return wav, src


class FakeEngine:
    sr = 24000

    def __init__(self, state_dict, device="cpu"):
        self.device = torch.device(device)
        self.s3gen = FakeS3Gen(R.load_model(state_dict)).to(self.device)
        self.last_mel = None

    @staticmethod
    def frames_for(text: str) -> int:
        return 24 + 5 * len(text)

    def mel_for(self, text: str) -> torch.Tensor:
        return R.synthetic_mel(1, self.frames_for(text), seed=len(text))

    def generate(self, text, audio_prompt_path=None, exaggeration=0.25, cfg_weight=0.5, temperature=0.8):
        mel = self.mel_for(text).to(self.device)
        self.last_mel = mel
        wav, _ = self.s3gen.inference(mel)
        return wav                                   # [1, N], like ChatterboxTTS.generate
