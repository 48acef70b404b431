"""Importable alias of the ``vocalie-tts_b200/`` package directory (a hyphen is not a valid
Python identifier).  Loads ``vocalie-tts_b200/__init__.py`` under this module's name so that
``import vocalie_tts_b200`` and ``vocalie_tts_b200.<submodule>`` resolve into that directory."""
import importlib.util as _u
import pathlib as _p
import sys as _s

_real = _p.Path(__file__).resolve().parent.parent / "vocalie-tts_b200"
_spec = _u.spec_from_file_location(__name__, _real / "__init__.py", submodule_search_locations=[str(_real)])
_mod = _u.module_from_spec(_spec)
_s.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
