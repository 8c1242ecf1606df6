"""The reference's own acceptance tests, UNMODIFIED, against the drop-in `largecrimsoncanine` module built here.

tests/golden/reftests.tar.gz holds the six test files of /root/reference/tests that exercise the hot path and its callers
(test_batch 40, test_algebras 35, test_pga 42, test_cga 35, test_sta 43, test_basic 795 = 990 tests), packed byte for
byte by tests/golden/make_reftests.py; the manifest's SHA-256 digests are re-checked here before anything runs.  They are
unpacked into a scratch directory and run with pytest in a subprocess whose import path starts at the repo root, so
`from largecrimsoncanine import Algebra, Multivector, MultivectorBatch` resolves to the module under test."""
import hashlib
import json
import os
import re
import subprocess
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
EXPECTED = {"test_batch.py": 40, "test_algebras.py": 35, "test_pga.py": 42, "test_cga.py": 35, "test_sta.py": 43, "test_basic.py": 795}


@pytest.fixture(scope="module")
def suite_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("reftests")
    manifest = json.load(open(os.path.join(GOLDEN, "reftests.manifest.json")))["files"]
    with tarfile.open(os.path.join(GOLDEN, "reftests.tar.gz")) as tar:
        tar.extractall(d, filter="data")
    for name, meta in manifest.items():
        assert hashlib.sha256(open(os.path.join(d, name), "rb").read()).hexdigest() == meta["sha256"], f"{name} differs from the reference's file"
    assert set(manifest) == set(EXPECTED)
    return str(d)


def test_archive_matches_its_manifest(suite_dir):
    assert sorted(os.listdir(suite_dir)) == sorted(EXPECTED)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_reference_test_file_passes_unmodified(suite_dir, name):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(suite_dir, name), "-q", "-p", "no:cacheprovider", "--rootdir", suite_dir],
                         capture_output=True, text=True, timeout=900, env=env, cwd=suite_dir)
    tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-500:]
    m = re.search(r"(\d+) passed", tail)
    assert res.returncode == 0 and m and int(m.group(1)) == EXPECTED[name] and "failed" not in tail, f"{name}: {tail}\n{res.stdout[-3000:]}"
