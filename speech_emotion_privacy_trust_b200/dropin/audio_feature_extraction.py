"""Drop-in for feature_extraction/audio_feature_extraction.py: the two extraction callables keep their signatures and
return types (reference :15-46) but run the fused sm_100a kernels; one utterance per call like the reference.  Bulk
work should use speech_emotion_privacy_trust_b200.extraction (ragged batches, one launch per feature)."""
import numpy as np
import torch

from speech_emotion_privacy_trust_b200 import extraction as _ex


def _ragged(audio):
    a = audio if torch.is_tensor(audio) else torch.as_tensor(np.asarray(audio))
    if a.dim() != 2:
        raise ValueError(f"audio must be (channels, samples); got {tuple(a.shape)}")
    dev = a.device if a.is_cuda else torch.device("cuda")
    wav = a[0].to(device=dev, dtype=torch.float32).contiguous()
    return _ex.RaggedAudio(wav, np.array([0, wav.numel()], dtype=np.int64)), a.shape[0]


def mel_spectrogram(audio, n_fft=1024, feature_len=128):
    """log-mel dB, (1, N) float32 -> CPU tensor (1, feature_len, 1 + N // 160)."""
    batch, channels = _ragged(audio)
    if channels != 1:
        raise ValueError("mel_spectrogram drop-in handles mono audio (the reference's corpora are mono)")
    flat, lay = _ex.logmel(batch, n_fft=n_fft, n_mels=feature_len, band_major=True)
    return flat.view(1, feature_len, lay.total_frames).cpu()


def mfcc(audio):
    """MFCC-40 of the waveform and its two numerical derivatives -> float32 ndarray (1, 120, 1 + N // 200)."""
    batch, channels = _ragged(audio)
    if channels != 1:
        raise ValueError("mfcc drop-in handles mono audio (the reference's corpora are mono)")
    flat, lay = _ex.mfcc(batch)
    return flat.view(1, 3 * _ex.N_MFCC, lay.total_frames).cpu().numpy()
