"""Build recipe for the C oracle (gcc only; output stays under oracle/_build/)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "liboracle.so")


def lib_path() -> str:
    return _OUT


def build(force: bool = False) -> str:
    """Compile oracle/csrc/oracle.c -> oracle/_build/liboracle.so (idempotent)."""
    if (not force and os.path.exists(_OUT)
            and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC)):
        return _OUT
    os.makedirs(_OUT_DIR, exist_ok=True)
    # -ffp-contract=off: the restatement spells out every fmaf itself.
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
           "-ffp-contract=off", "-fno-fast-math", "-o", _OUT, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _OUT
