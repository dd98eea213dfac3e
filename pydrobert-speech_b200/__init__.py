"""B200-native frame-feature extraction behind the pydrobert-speech API.

Importable as ``pydrobert_speech_b200`` (see ``pydrobert_speech_b200.py`` at the repo root).
Module layout mirrors ``pydrobert.speech`` so existing code only changes its import::

    from pydrobert_speech_b200 import compute, util
    computer = util.alias_factory_subclass_from_arg(compute.FrameComputer, json_config)
    feats = computer.compute_full(signal)            # one utterance (reference API)
    feats = computer.compute_batch(list_of_signals)  # whole batch, one kernel launch
"""

__version__ = "0.1.0"

from . import alias, config, scales, util, filters, pre, post, compute  # noqa: F401,E402
from .alias import alias_factory_subclass_from_arg, AliasedFactory  # noqa: F401,E402

# the reference exposes the factory helper from `util` as well (README.md:33)
util.alias_factory_subclass_from_arg = alias_factory_subclass_from_arg

__all__ = [
    "alias",
    "compute",
    "config",
    "filters",
    "post",
    "pre",
    "scales",
    "util",
]
