"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the timed CPU arm.
The product package (``speech_emotion_privacy_trust_b200``) never imports it and
fails loudly when its CUDA library is missing.

Modules
-------
restate      numpy restatement (fp32 and fp64) of the extraction chain that the
             reference delegates to torchaudio/torch.stft.
ref_port     the reference's two public callables restated on top of the same
             third-party dependency the reference calls (torchaudio 2.11.0);
             this is the timed CPU arm ("kind": "port").
cloak        numpy restatement of cloak_noise / GradientReversal fwd+bwd.
norm         numpy restatement of the per-speaker normalisation block.

Pinning: the reference holds no tests or golden vectors for this path
(SURVEY 8c), so the oracle is pinned against outputs of the reference itself,
imported in the build container by ``oracle/make_golden.py`` and committed under
``tests/golden/``.
"""
