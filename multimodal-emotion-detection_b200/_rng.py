"""Where the stochastic passes get their masks from.

The reference's nn.Dropout draws from torch's global generator (src/uncertainty.py:61-65 passes no
generator; src/train.py:472 is the only seeding).  Here every dropout site draws Philox-4x32-10 bits that
are a pure function of (seed, pass index t, global clip index, site, element) - see oracle/philox.py for the
written spec - so masks do not depend on batch size, sharding or launch geometry.  This module holds the
(seed, t, clip0) a forward call should use; MCDropoutUncertainty advances t once per stochastic pass.
"""
from __future__ import annotations

import contextlib
import threading

_state = threading.local()
DEFAULT_SEED = 42          # configs/base.yaml:4


def current():
    """(seed, t, clip0) for the forward call in progress."""
    return getattr(_state, "ctx", (DEFAULT_SEED, 0, 0))


@contextlib.contextmanager
def philox_context(seed: int = DEFAULT_SEED, t: int = 0, clip0: int = 0):
    """Masks drawn inside use (seed, pass t, clips clip0 + b)."""
    old = getattr(_state, "ctx", None)
    _state.ctx = (int(seed) & 0xFFFFFFFFFFFFFFFF, int(t) & 0xFFFFFFFF, int(clip0) & 0xFFFFFFFFFFFFFFFF)
    try:
        yield
    finally:
        if old is None:
            del _state.ctx
        else:
            _state.ctx = old
