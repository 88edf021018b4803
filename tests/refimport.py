"""Import the unmodified reference `models.py` (build container only; /root/reference does not travel)."""
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"


def have_reference():
    """The reference model is importable: from /root/reference (build container) or from the compiled oracle/_ref (GPU box)."""
    from oracle import build_ref
    return os.path.isfile(os.path.join(REF_ROOT, "models.py")) or build_ref.available()


def have_reference_tree():
    """The whole reference tree (datasets/, pretrain.py, ...) is mounted: build container only."""
    return os.path.isfile(os.path.join(REF_ROOT, "pretrain.py"))


def import_reference_models():
    """Returns the reference `models` module under the name `ref_models` (music21 stubbed: humdrum.py:4 imports it
    at module top but the model never uses it)."""
    if "ref_models" in sys.modules:
        return sys.modules["ref_models"]
    if not os.path.isfile(os.path.join(REF_ROOT, "models.py")):
        from oracle import build_ref
        return build_ref.load()
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


def reference_dataset_stub(max_frame_num=1201, device="cpu"):
    """An instance of the reference's `datasets.syn.SyntheticDataset` WITHOUT running its __init__ (which lists feature folders):
    only the attributes `pad_spectrogram` / `pad_score` / `pad_single_measure` / `key_to_int` read are set.  The third-party modules
    datasets/syn.py and utilities.py import at module top but these methods never touch (pretty_midi, librosa, mido, hyperpyyaml,
    music21) are stubbed when absent."""
    def stub(name, **attrs):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[name] = m
    for name, attrs in (("music21", {}), ("pretty_midi", {}), ("librosa", {}), ("mido", {"MidiFile": object}),
                        ("hyperpyyaml", {"load_hyperpyyaml": None})):
        stub(name, **attrs)
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k, None) for k in ("data_processing", "datasets", "utilities")}
    sys.path.insert(0, REF_ROOT)
    try:
        spec = importlib.util.spec_from_file_location("ref_syn", os.path.join(REF_ROOT, "datasets", "syn.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in ("data_processing", "datasets", "utilities"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    ds = object.__new__(mod.SyntheticDataset)
    ds.hparams = {"max_frame_num": max_frame_num}
    ds.device = device
    ds.labels = mod.LabelsMultiple(extended=True)
    return ds


def import_reference_trainer(name="pretrain"):
    """The reference's `pretrain.py` / `finetune.py` as a module, with the packages its top-level imports need but the functions
    under test never touch replaced by stubs when absent (speechbrain, hyperpyyaml, jiwer, pretty_midi, librosa, mido, music21).
    `speechbrain.Brain` becomes a plain base class and `speechbrain.Stage` an enum with TRAIN / VALID / TEST."""
    import enum
    key = f"ref_{name}"
    if key in sys.modules:
        return sys.modules[key]

    def stub(modname, **attrs):
        if modname in sys.modules:
            return sys.modules[modname]
        try:
            return importlib.import_module(modname)
        except Exception:
            m = types.ModuleType(modname)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[modname] = m
            return m

    class Stage(enum.Enum):
        TRAIN, VALID, TEST = 1, 2, 3

    sb = stub("speechbrain", Brain=type("Brain", (), {}), Stage=Stage)
    utils = stub("speechbrain.utils")
    distm = stub("speechbrain.utils.distributed", run_on_main=lambda f, *a, **k: f(*a, **k), if_main_process=lambda: True)
    if not hasattr(sb, "utils"):
        sb.utils = utils
    if not hasattr(utils, "distributed"):
        utils.distributed = distm
    stub("hyperpyyaml", load_hyperpyyaml=None)
    stub("jiwer", wer=None)
    for n, a in (("music21", {}), ("pretty_midi", {}), ("librosa", {}), ("mido", {"MidiFile": object})):
        stub(n, **a)
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k, None) for k in ("data_processing", "datasets", "utilities")}
    sys.path.insert(0, REF_ROOT)
    # the reference's `datasets/` has no __init__.py: an installed `datasets` distribution would win over the namespace package
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(REF_ROOT, "datasets")]
    sys.modules["datasets"] = pkg
    try:
        spec = importlib.util.spec_from_file_location(key, os.path.join(REF_ROOT, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[key] = mod
    finally:
        sys.path[:] = saved_path
        for k in [m for m in sys.modules if m == "datasets" or m.startswith("datasets.")] + ["data_processing", "utilities"]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    return mod
