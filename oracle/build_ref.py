"""Builds `oracle/_ref/`: the UNMODIFIED reference model, compiled where its sources lie.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/a2s_oracle.py): nothing under piano_a2s_b200/ may import this.

The reference's hot path is two Python files (`/root/reference/models.py`, which imports `LabelsMultiple` from
`/root/reference/data_processing/humdrum.py`).  They are byte-compiled from `/root/reference` -- no source is copied into
this repository -- and only the compiled bytecode files (`*.pyc.bin`) are written to `oracle/_ref/` (git-ignored, NOT gpurun-ignored, so
they travel to the GPU box like our own `.so`; same image => same CPython => the bytecode loads there).  `load()` imports
the result as module `ref_models` with `music21` stubbed (humdrum.py:4 imports it at module top; the model never uses
it).  With it
  * `bench.py --impl reference` times the reference's own `ScoreTranscription` on the host cores (cpu_baseline.kind
    "reference") and `--impl reference-gpu` runs the same module on the B200 through stock PyTorch (cuDNN / cuBLAS);
  * the live-reference pin tests of tests/ also run on the GPU box.
When /root/reference is absent (the GPU box) build() keeps whatever is already in oracle/_ref/.
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
OUT = os.path.join(HERE, "_ref")
# (the compiled files are named *.bin: the gpurun snapshot skips *.pyc)
FILES = (("models.py", "models.pyc.bin"), (os.path.join("data_processing", "humdrum.py"), os.path.join("data_processing", "humdrum.pyc.bin")))


def build(verbose=False):
    """-> True when oracle/_ref/ holds the compiled reference (freshly built or already there)."""
    if os.path.isfile(os.path.join(REF_SRC, "models.py")):
        for src, dst in FILES:
            out = os.path.join(OUT, dst)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            py_compile.compile(os.path.join(REF_SRC, src), cfile=out, dfile=os.path.join("reference", src), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            if verbose:
                print(f"py_compile {os.path.join(REF_SRC, src)} -> {out}", file=sys.stderr)
    return available()


def available():
    return all(os.path.isfile(os.path.join(OUT, dst)) for _, dst in FILES)


def load():
    """The reference `models` module (as `ref_models`), from oracle/_ref/ bytecode."""
    if "ref_models" in sys.modules:
        return sys.modules["ref_models"]
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python -m oracle.build_ref` where /root/reference exists")
    sys.modules.setdefault("music21", types.ModuleType("music21"))
    saved_dp = sys.modules.pop("data_processing", None)
    try:
        pkg = types.ModuleType("data_processing")
        pkg.__path__ = [os.path.join(OUT, "data_processing")]
        sys.modules["data_processing"] = pkg
        for name, rel in (("data_processing.humdrum", FILES[1][1]), ("ref_models", FILES[0][1])):
            loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(OUT, rel))
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            loader.exec_module(mod)
            if name == "data_processing.humdrum":
                pkg.humdrum = mod
    finally:
        sys.modules.pop("data_processing", None)
        sys.modules.pop("data_processing.humdrum", None)
        if saved_dp is not None:
            sys.modules["data_processing"] = saved_dp
    return sys.modules["ref_models"]


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT built (no /root/reference and nothing prebuilt)")
