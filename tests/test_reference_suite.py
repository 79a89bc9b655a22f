"""The reference's OWN test suite (`tests/test_rankfm.py` of etlundquist/rankfm, 17 tests: shapes, dtypes, NaN counts, id
membership, exception types -- SURVEY.md section 4), run UNCHANGED against rankfm_b200 through the `rankfm` alias package in
tests/ref_alias/ ("IN SCOPE to pass unchanged").

The file is never copied into the repository: here it is read from /root/reference; for the GPU box oracle/build_ref.py
stages a copy under the git-ignored oracle/_ref/tests/ (which travels with the snapshot like the compiled reference)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALIAS = os.path.join(ROOT, "tests", "ref_alias")
CANDIDATES = ["/root/reference/tests/test_rankfm.py", os.path.join(ROOT, "oracle", "_ref", "tests", "test_rankfm.py")]


def _reference_tests():
    for path in CANDIDATES:
        if os.path.exists(path):
            return path
    return None


def _run(backend):
    path = _reference_tests()
    if path is None:
        pytest.skip("the reference's tests/test_rankfm.py is not on this machine")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ALIAS, ROOT, os.environ.get("PYTHONPATH", "")]))
    if backend == "oracle":
        env["RANKFM_ALIAS_BACKEND"] = "oracle"
    else:
        env.pop("RANKFM_ALIAS_BACKEND", None)
    out = subprocess.run([sys.executable, "-m", "pytest", path, "-q", "-p", "no:cacheprovider", "--rootdir", os.path.dirname(path), "-c", os.devnull],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ALIAS)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) >= 17 and "failed" not in out.stdout, tail
    return int(m.group(1))


def test_reference_suite_unchanged_host_logic_on_the_oracle_backend():
    _run("oracle")


@pytest.mark.gpu
def test_reference_suite_unchanged_on_the_cuda_path(gpu_lib):
    _run("cuda")
