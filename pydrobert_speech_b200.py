"""Import shim: exposes the ``pydrobert-speech_b200/`` directory (hyphenated, hence not
directly importable) as the package ``pydrobert_speech_b200``.

``import pydrobert_speech_b200`` (with the repo root on ``sys.path``) replaces this module in
``sys.modules`` with the real package, so ``pydrobert_speech_b200.compute`` etc. resolve to the
files under ``pydrobert-speech_b200/``.
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "pydrobert-speech_b200")
_spec = importlib.util.spec_from_file_location(
    __name__,
    os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir],
)
_module = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _module
_spec.loader.exec_module(_module)
