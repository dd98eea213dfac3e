"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

A float64 NumPy restatement of the frame-feature hot path of pydrobert-speech, used to check
the CUDA kernels.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``pydrobert-speech_b200/`` does, and the product has no CPU fallback.

Pinning: every function here is checked (``tests/test_oracle_golden.py``) against golden vectors
in ``tests/golden/`` that were produced by running the *real* reference
(``/root/reference/src/pydrobert/speech``, numpy/scipy float64 path) in the build container with
``tests/golden/make_golden.py``, and against the reference's own Kaldi known-answer fixtures
(``tests/data/kaldi_feats.pkl``, ``kaldi_filts.pkl``, ``noise.pkl``) carried over by that script.
"""

from .stft import stft_features, stft_features_looped, frame_signal  # noqa: F401
from .post import cmvn_accumulate, cmvn_apply, deltas  # noqa: F401
from .pre import preemphasize  # noqa: F401
from .si import si_features  # noqa: F401
