"""TEST INFRASTRUCTURE ONLY -- recipe for ``oracle/_ref``: the UNMODIFIED reference numerics, byte-compiled.

The reference is pure Python, so "building" it means byte-compiling the four source files of the hot path from where
they lie under ``/root/reference`` (nothing is copied: no reference SOURCE enters this repository) into sourceless
``.pyc`` files under ``oracle/_ref/`` (git-ignored, but shipped to the GPU box with the snapshot, like the built
``libb200ms.so``):

    tidy3d/constants.py                  -> oracle/_ref/tidy3d/constants.pyc
    tidy3d/plugins/mode/derivatives.py   -> oracle/_ref/tidy3d/plugins/mode/derivatives.pyc
    tidy3d/plugins/mode/transforms.py    -> oracle/_ref/tidy3d/plugins/mode/transforms.pyc
    tidy3d/plugins/mode/solver.py        -> oracle/_ref/tidy3d/plugins/mode/solver.pyc

``oracle/ref_shim.py`` loads them (stub package) when ``/root/reference`` is absent, which makes
the reference's own ``compute_modes`` (solver.py:941, scipy ARPACK + SuperLU) runnable on the GPU box: it is what
``bench.py --impl reference`` and the ``cpu_baseline`` leg time there (``kind: "reference"``), and what the
``reference``-marked tests check the restatement against.  The byte code is tied to the interpreter that made it
(``importlib.util.MAGIC_NUMBER`` is recorded in the manifest; the GPU box runs this same image).

    python oracle/build_ref.py            # no-op when oracle/_ref is up to date
"""
import hashlib
import importlib.util
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ("constants.py", "plugins/mode/derivatives.py", "plugins/mode/transforms.py", "plugins/mode/solver.py")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def manifest_path():
    return os.path.join(OUT, "MANIFEST.json")


def build(ref_root=None, force=False) -> bool:
    """Byte-compile the reference numerics into ``oracle/_ref``.  Returns True when ``oracle/_ref`` is usable afterwards."""
    ref_root = ref_root or os.environ.get("B200MS_REFERENCE", "/root/reference")
    src_root = os.path.join(ref_root, "tidy3d")
    if not os.path.isfile(os.path.join(src_root, "plugins", "mode", "solver.py")):
        return os.path.isfile(manifest_path())  # GPU box: use what travelled with the snapshot
    want = {f: _sha(os.path.join(src_root, f)) for f in FILES}
    magic = importlib.util.MAGIC_NUMBER.hex()
    try:
        have = json.load(open(manifest_path()))
    except (OSError, ValueError):
        have = {}
    up_to_date = have.get("sha256") == want and have.get("magic") == magic and all(
        os.path.isfile(os.path.join(OUT, "tidy3d", f + "c")) for f in FILES)
    if up_to_date and not force:
        return True
    for f in FILES:
        dst = os.path.join(OUT, "tidy3d", f + "c")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path shown in tracebacks; UNCHECKED_HASH: never looks for the source file again
        py_compile.compile(os.path.join(src_root, f), cfile=dst, dfile=f"<reference>/tidy3d/{f}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    version = None
    try:
        for ln in open(os.path.join(src_root, "version.py")):
            if ln.startswith("__version__"):
                version = ln.split("=")[1].strip().strip("\"'")
    except OSError:
        pass
    json.dump({"what": "byte-compiled, unmodified flexcompute/tidy3d numerics (see oracle/build_ref.py)", "tidy3d_version": version,
               "python": sys.version.split()[0], "magic": magic, "sha256": want}, open(manifest_path(), "w"), indent=1)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "not built (no reference tree here and nothing shipped)")
