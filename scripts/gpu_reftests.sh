#!/bin/bash
# Run the reference's own Python test files (staged, git-ignored, under _reftests/) against the drop-in module.
mkdir -p gpurun_out
cd _reftests
for t in test_batch test_algebras test_pga test_cga test_sta test_basic; do
  timeout 900 python -m pytest $t.py -q --no-header -p no:cacheprovider --maxfail=2000 --tb=line -rf 2>&1 | tail -400 > ../gpurun_out/ref_$t.log
  echo "== $t: $(tail -1 ../gpurun_out/ref_$t.log)"
done
