"""Import the unmodified reference `models.py` (build container only; /root/reference does not travel)."""
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"


def have_reference():
    return os.path.isfile(os.path.join(REF_ROOT, "models.py"))


def import_reference_models():
    """Returns the reference `models` module under the name `ref_models` (music21 stubbed: humdrum.py:4 imports it
    at module top but the model never uses it)."""
    if "ref_models" in sys.modules:
        return sys.modules["ref_models"]
    sys.modules.setdefault("music21", types.ModuleType("music21"))
    saved_path = list(sys.path)
    saved_dp = sys.modules.pop("data_processing", None)
    sys.path.insert(0, REF_ROOT)
    try:
        spec = importlib.util.spec_from_file_location("ref_models", os.path.join(REF_ROOT, "models.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["ref_models"] = mod
    finally:
        sys.path[:] = saved_path
        if saved_dp is not None:
            sys.modules["data_processing"] = saved_dp
    return mod
