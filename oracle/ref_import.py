"""Load the UNMODIFIED reference functions from /root/reference (test infrastructure).

Only usable in the authoring container: /root/reference does not exist on the GPU box,
so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.

The reference imports ``matplotlib`` (bluenoise/get_noise_recent.py:3, dead code only)
and ``diffusers`` (utils.py:3, only to *construct* the UNet); neither is installed here,
so empty stand-in modules are inserted into ``sys.modules`` before the import.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BNDM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bluenoise", "get_noise_recent.py"))


def _stub(name: str, **attrs) -> None:
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod


def load_reference():
    """Returns (get_noise_v2, noise_padding, utils_module) from the real reference."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _stub("diffusers", UNet2DModel=type("UNet2DModel", (), {}))

    # The reference's bluenoise/ has no __init__.py (a namespace package) while this repo ships a
    # regular `bluenoise` shim package, which would win any path-based import: load by file path.
    def _by_path(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    saved_utils = sys.modules.pop("utils", None)
    try:
        gnr = _by_path("_bndm_reference_get_noise_recent", os.path.join(REFERENCE_ROOT, "bluenoise", "get_noise_recent.py"))
        ref_utils = _by_path("_bndm_reference_utils", os.path.join(REFERENCE_ROOT, "utils.py"))
    finally:
        if saved_utils is not None:
            sys.modules["utils"] = saved_utils
    assert gnr.__file__.startswith(REFERENCE_ROOT), gnr.__file__
    return gnr.get_noise_v2, gnr.noise_padding, ref_utils
