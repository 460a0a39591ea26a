"""Import the UNMODIFIED reference's model package (test infrastructure; CPU baselines and golden generation only).

Source of the modules, in this order:
  1. /root/reference (this container): imported in place from the read-only tree;
  2. oracle/_ref/<tree>/model/*.pyc: the sourceless bytecode `oracle/build_ref.py` compiled from that tree (what the GPU box
     has: /root/reference does not exist there).
Both are the reference's own code, unmodified; `kind()` says which one is live.  The two training-only / audio-only imports
the package pulls in at import time are stubbed exactly as SURVEY.md 8(c) prescribes (`model.monotonic_align` is a Cython
extension used by `compute_loss` only; `librosa.filters.mel` is used by DiffVC's `FastGL` only) - the sampler never calls them.
Grad-TTS and DiffVC both name their package `model`, so importing one purges the other from `sys.modules`.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_REF = "/root/reference"
_TREES = {"gradtts": "Grad-TTS", "diffvc": "DiffVC"}


def _roots(tree):
    src = os.path.join(_REF, _TREES[tree])
    byte = os.path.join(HERE, "_ref", tree)
    return src, byte


def available(tree="gradtts") -> bool:
    src, byte = _roots(tree)
    return os.path.isdir(os.path.join(src, "model")) or os.path.exists(os.path.join(byte, "model", "diffusion.pyc"))


def kind(tree="gradtts") -> str:
    src, _ = _roots(tree)
    return "source tree /root/reference" if os.path.isdir(os.path.join(src, "model")) else "oracle/_ref bytecode"


def _purge():
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    drop = [os.path.join(_REF, v) for v in _TREES.values()] + [os.path.join(HERE, "_ref", k) for k in _TREES]
    sys.path[:] = [p for p in sys.path if p not in drop]


def import_model(tree="gradtts"):
    """Returns the reference's `model.diffusion` module of `tree` ('gradtts' | 'diffvc')."""
    src, byte = _roots(tree)
    root = src if os.path.isdir(os.path.join(src, "model")) else byte
    if not os.path.isdir(os.path.join(root, "model")):
        raise RuntimeError(f"the reference is not available: neither {src} nor {byte} exists (run oracle/build_ref.py "
                           "in the container that has /root/reference)")
    _purge()
    sys.path.insert(0, root)
    if tree == "gradtts":
        sys.modules["model.monotonic_align"] = types.ModuleType("model.monotonic_align")
    else:
        for n in ("librosa", "librosa.filters"):
            sys.modules.setdefault(n, types.ModuleType(n))
        sys.modules["librosa.filters"].mel = lambda *a, **k: None
        sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    import model.diffusion as md
    return md
