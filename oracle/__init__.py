"""CPU oracle for the audio-visual inference hot path.  TEST INFRASTRUCTURE ONLY.

This package is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package
(``multimodal-emotion-detection_b200/``) must never import anything from here
and fails loudly when its CUDA library is missing.

What is restated, and how each part is pinned:

* ``oracle.philox``    Philox-4x32-10 counter RNG (Random123 algorithm) and
                       the dropout-mask / synthetic-data specs built on it.
                       Pinned by the Random123 known-answer vectors
                       (tests/test_oracle_philox.py).
* ``oracle.frontend``  log-mel + MFCC front-end.  The reference has NO
                       implementation of this stage, so
                       **parity unpinned** against the reference; it is
                       cross-checked against torch.stft /
                       torchaudio.melscale_fbanks / scipy.fft.dct.
* ``oracle.model``     encoders -> cross-modal attention -> fusion ->
                       MC-dropout uncertainty, restated from
                       /root/reference/src/{encoders,attention,fusion,
                       uncertainty,train,data}.py (file:line cited per
                       function).  Pinned by golden vectors produced by running
                       the reference's own modules (tests/golden/make_golden.py
                       imports them from /root/reference/src) - the reference
                       ships no golden vectors or tests of its own.
"""
