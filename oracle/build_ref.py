"""Recipe: install the UNMODIFIED reference's sampler modules as compiled bytecode under oracle/_ref/ (test infrastructure).

    python oracle/build_ref.py            # needs /root/reference (this container); no-op elsewhere

The reference (huawei-noah/Speech-Backbones) is Python/PyTorch with no build system for this path: its "build" is
CPython's own byte compiler.  This script byte-compiles, WITHOUT modification and WITHOUT copying any source text into the
repository, the modules the hot path lives in

    Grad-TTS/model/{__init__, base, diffusion, text_encoder, tts, utils}.py   -> oracle/_ref/gradtts/model/*.pyc
    DiffVC/model/{__init__, base, diffusion, modules, encoder, postnet, utils, vc}.py -> oracle/_ref/diffvc/model/*.pyc

from where they lie under /root/reference (read-only) straight into oracle/_ref/ (sourceless .pyc files: build outputs only;
oracle/_ref/ is git-ignored so nothing of the reference enters the history, but it is NOT gpurun-ignored, so - like
libsbk.so - it travels to the GPU box, where /root/reference does not exist).  MANIFEST.json records the sha256 of every
source file that was compiled, the interpreter's bytecode magic and the reference commit if available, so a reader can
check that what runs as `cpu_baseline.kind = "reference"` is the unmodified reference.

`oracle/ref_import.py` imports from here (or directly from /root/reference when it exists).  Only tests/, smoke() and
bench.py's CPU legs use it.
"""
import hashlib
import importlib.util
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
TREES = {
    "gradtts": ("Grad-TTS/model", ["__init__", "base", "diffusion", "text_encoder", "tts", "utils"]),
    "diffvc": ("DiffVC/model", ["__init__", "base", "diffusion", "modules", "encoder", "postnet", "utils", "vc"]),
}


def build(force: bool = False) -> str | None:
    if not os.path.isdir(REF):
        return OUT if os.path.exists(os.path.join(OUT, "MANIFEST.json")) else None
    man_path = os.path.join(OUT, "MANIFEST.json")
    manifest = {"python": sys.version.split()[0], "magic": importlib.util.MAGIC_NUMBER.hex(), "files": {}}
    for name, (rel, mods) in TREES.items():
        dst_dir = os.path.join(OUT, name, "model")
        os.makedirs(dst_dir, exist_ok=True)
        for m in mods:
            src = os.path.join(REF, rel, m + ".py")
            dst = os.path.join(dst_dir, m + ".pyc")
            with open(src, "rb") as f:
                digest = hashlib.sha256(f.read()).hexdigest()
            manifest["files"][f"{rel}/{m}.py"] = digest
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                # dfile: the path shown in tracebacks; unchecked-hash pyc so the loader never looks for the source
                py_compile.compile(src, cfile=dst, dfile=f"<reference>/{rel}/{m}.py", doraise=True,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(man_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
