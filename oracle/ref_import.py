"""Load the UNMODIFIED reference functions from /root/reference (test infrastructure).

Only usable in the authoring container: /root/reference does not exist on the GPU box,
so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.

The reference imports ``matplotlib`` (bluenoise/get_noise_recent.py:3, dead code only)
and ``diffusers`` (utils.py:3, only to *construct* the UNet); neither is installed here,
so empty stand-in modules are inserted into ``sys.modules`` before the import.
"""
from __future__ import annotations

import importlib
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

    # our repo also ships a `bluenoise` shim package; make sure the reference's wins here
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "bluenoise" or k.startswith("bluenoise.") or k == "utils"}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        gnr = importlib.import_module("bluenoise.get_noise_recent")
        ref_utils = importlib.import_module("utils")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "bluenoise" or k.startswith("bluenoise.") or k == "utils":
                sys.modules.pop(k)
        sys.modules.update(saved)
    assert gnr.__file__.startswith(REFERENCE_ROOT), gnr.__file__
    return gnr.get_noise_v2, gnr.noise_padding, ref_utils
