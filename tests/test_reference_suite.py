"""The reference's own test files, unmodified, pointed at the drop-in package (the API-compatibility gate).

The files are NOT in this repository: `__graft_entry__.build()` stages the unmodified reference (its package and its
`test/` directory with `test/games/**`) under the git-ignored `baseline/_ref/`, which travels to the GPU box.  A
plugin (tests/reference_suite/refsuite_plugin.py) aliases `import draughts` to `draughts_b200`; every in-scope file
(boards of all 8 variants, invariants, Move, PDN, ML helpers) must pass as it is.  Engine / hub / server / match
benchmark tests are out of scope (callers, not the hot path).

The CPU half runs two of the files against the reference itself: it proves the harness (package staging, the `test`
package shadowing fix) without a GPU."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
IN_SCOPE = ["test_board.py", "test_standard_board.py", "test_american_board.py", "test_russian_board.py",
            "test_frisian_board.py", "test_frysk_board.py", "test_brazilian_board.py", "test_antidraughts_board.py",
            "test_breakthrough_board.py", "test_board_invariants.py", "test_move.py", "test_pdn.py",
            "test_ai_features.py"]


def _run(name: str, target: str):
    if not (REF / "test" / name).exists():
        pytest.skip("reference not staged: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    env = dict(os.environ, REFSUITE_TARGET=target,
               PYTHONPATH=str(ROOT / "tests" / "reference_suite") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", str(REF / "test" / name), "-q", "-x", "-p", "refsuite_plugin",
                        "-p", "no:cacheprovider", "--rootdir", str(REF), "-c", os.devnull],
                       cwd=str(REF), env=env, capture_output=True, text=True, timeout=1500)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
    assert r.returncode == 0, f"{name} against {target}:\n{tail}"
    assert " passed" in r.stdout, tail


@pytest.mark.gpu
@pytest.mark.parametrize("name", IN_SCOPE)
def test_reference_file_passes_on_the_shim(name):
    _run(name, "shim")


@pytest.mark.parametrize("name", ["test_move.py", "test_american_board.py"])
def test_harness_on_the_reference_itself(name):
    _run(name, "reference")
