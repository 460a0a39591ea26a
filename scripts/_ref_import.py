"""Import the UNMODIFIED reference modules from /root/reference (container only).
Recipe from SURVEY.md 8(c): stub the training-only Cython extension, then import."""
import sys
import types


def _purge():
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    for p in ("/root/reference/Grad-TTS", "/root/reference/DiffVC"):
        while p in sys.path:
            sys.path.remove(p)


def import_gradtts():
    _purge()
    sys.path.insert(0, "/root/reference/Grad-TTS")
    sys.modules["model.monotonic_align"] = types.ModuleType("model.monotonic_align")
    import model.diffusion as md
    return md


def import_diffvc():
    _purge()
    sys.path.insert(0, "/root/reference/DiffVC")
    for n in ("librosa", "librosa.filters"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["librosa.filters"].mel = lambda *a, **k: None
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    import model.diffusion as md
    return md
