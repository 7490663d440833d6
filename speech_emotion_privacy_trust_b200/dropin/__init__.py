"""Drop-in modules with the reference's module names.

The reference's drivers do ``sys.path.append('../model')`` and ``from cloak_models import ...``
(training/training_cloak_with_grl.py:19-25).  Put THIS directory on ``sys.path`` ahead of the reference's ``model/`` and
``feature_extraction/`` directories (``install()`` does it) and those imports resolve to the B200 implementations:

    audio_feature_extraction   mel_spectrogram(audio, n_fft=1024, feature_len=128), mfcc(audio)
    reversal_gradient          GradientReversalFunction, GradientReversal (+ alias ReverseLayerF)
    cloak_models               cloak_noise, two_d_cnn_lstm_syn, two_d_cnn_lstm_syn_with_grl
    baseline_models            two_d_cnn_lstm, deep_two_d_cnn_lstm
"""
import sys
from pathlib import Path

HERE = str(Path(__file__).resolve().parent)


def install() -> str:
    """Prepend this directory to sys.path (idempotent) so the reference's flat imports pick up the drop-ins."""
    if HERE in sys.path:
        sys.path.remove(HERE)
    sys.path.insert(0, HERE)
    for name in ("audio_feature_extraction", "reversal_gradient", "cloak_models", "baseline_models"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(HERE):
            del sys.modules[name]
    return HERE
