"""TEST INFRASTRUCTURE — builds oracle/_ref/: the UNMODIFIED reference's hot path as sourceless byte code, so that the
real `simulate_game` (core/game_simulator.py:14-156, the loop body of cli/run_distribution_simulator.py:368-382) can be
TIMED and SAMPLED on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py          (run by __graft_entry__.build() in the build container)

What it does: imports the reference read-only from /root/reference exactly as oracle/ref_harness.py does, lists the
reference modules that import pulled in (core.game_simulator and everything it reaches: half_inning_simulator,
pa_simulator, bip_resolution, fatigue_modeling, assets.bullpen_utils, ...), and compiles each of those files with
`py_compile` to oracle/_ref/<package>/<module>.refbc (byte code in the .pyc format under another extension: snapshot
tools drop *.pyc).  No reference SOURCE is copied: oracle/_ref/ holds compiler output only, it is git-ignored (never
enters history) and it is not gpurun-ignored (it travels to the GPU box like the other built artefacts).
oracle/ref_harness.py imports from oracle/_ref/ through a small finder when /root/reference is absent.
"""
from __future__ import annotations

import json
import os
import py_compile
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MLBSIM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
EXT = ".refbc"

_LIST = r"""
import json, os, sys
sys.path.insert(0, %r)
os.environ["MLBSIM_REFERENCE_ROOT"] = %r
import ref_harness
ref_harness.load_reference()
root = os.path.realpath(%r) + os.sep
files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(root)})
print(json.dumps(files))
"""


def build(verbose: bool = False) -> bool:
    """Returns True if oracle/_ref/ is usable afterwards (freshly built, or kept from an earlier build)."""
    if not os.path.isdir(os.path.join(SRC, "core")):
        return os.path.isfile(os.path.join(DST, "core", "game_simulator" + EXT))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", _LIST % (HERE, SRC, SRC)], capture_output=True, text=True, env=env, cwd="/tmp")
    if r.returncode != 0:
        raise RuntimeError("could not import the reference:\n" + r.stderr[-2000:])
    files = json.loads(r.stdout.strip().splitlines()[-1])
    tmp = DST + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    root = os.path.realpath(SRC)
    for f in files:
        rel = os.path.relpath(f, root)
        out = os.path.join(tmp, os.path.splitext(rel)[0] + EXT)  # core/game_simulator.py -> core/game_simulator.refbc
        os.makedirs(os.path.dirname(out), exist_ok=True)
        py_compile.compile(f, cfile=out, dfile=os.path.join("<reference>", rel), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as fh:
        json.dump({"built_from": SRC, "python": sys.version.split()[0], "modules": [os.path.relpath(f, root) for f in files]}, fh, indent=1)
    shutil.rmtree(DST, ignore_errors=True)
    os.replace(tmp, DST)
    if verbose:
        print(f"oracle/_ref: {len(files)} reference modules compiled from {SRC}")
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    sys.exit(0 if ok else 1)
