#!/usr/bin/env python3
"""Packs the reference's OWN acceptance tests for the hot path -- tests/test_batch.py, test_algebras.py, test_pga.py,
test_cga.py, test_sta.py, test_basic.py of /root/reference, byte for byte -- into tests/golden/reftests.tar.gz so the
drop-in claim ("the reference's test files pass unmodified against this module") can be re-checked on the GPU box,
where /root/reference does not exist (tests/test_reference_suites.py).  The archive is a test FIXTURE, like
ref_torch_golden.npz: nothing in the product reads it.  reftests.manifest.json records the SHA-256 of every file so a
reviewer can confirm they are unmodified.  test_viz.py (lcc/viz, out of scope) is not included.

    python tests/golden/make_reftests.py          # in the container that has /root/reference
"""
import hashlib
import io
import json
import os
import tarfile

REF = "/root/reference/tests"
FILES = ["test_batch.py", "test_algebras.py", "test_pga.py", "test_cga.py", "test_sta.py", "test_basic.py"]
HERE = os.path.dirname(os.path.abspath(__file__))

manifest = {}
buf = io.BytesIO()
with tarfile.open(fileobj=buf, mode="w:gz", compresslevel=9) as tar:
    for name in FILES:
        data = open(os.path.join(REF, name), "rb").read()
        manifest[name] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
        info = tarfile.TarInfo(name)
        info.size, info.mtime, info.mode = len(data), 0, 0o644  # fixed metadata: the archive is reproducible
        tar.addfile(info, io.BytesIO(data))
open(os.path.join(HERE, "reftests.tar.gz"), "wb").write(buf.getvalue())
json.dump({"source": "LargeCrimsonCanine/largecrimsoncanine tests/ (unmodified)", "files": manifest},
          open(os.path.join(HERE, "reftests.manifest.json"), "w"), indent=1, sort_keys=True)
print({k: v["bytes"] for k, v in manifest.items()}, len(buf.getvalue()), "bytes packed")
