"""Drop-in for model/reversal_gradient.py: same names, same call signatures, backward on the B200 kernel."""
import torch

from speech_emotion_privacy_trust_b200.cloak_ops import GradientReversalFunction  # noqa: F401  (apply(x, lambda_))


class GradientReversal(torch.nn.Module):
    """Identity forward; multiplies the incoming gradient by -lambda_ (reference :26-32)."""

    def __init__(self, lambda_=1):
        super().__init__()
        self.lambda_ = lambda_

    def forward(self, x):
        return GradientReversalFunction.apply(x, self.lambda_)


ReverseLayerF = GradientReversalFunction   # the name BASELINE.json uses for the same function
