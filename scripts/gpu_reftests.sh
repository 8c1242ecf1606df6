#!/bin/bash
# Run the reference's own Python test files UNMODIFIED against the drop-in module on a GPU box.
# The files are not part of this repository.  Stage them right before the call and remove them afterwards:
#   mkdir -p _reftests && cp /root/reference/tests/{conftest,test_batch,test_algebras,test_pga,test_cga,test_sta,test_basic}.py _reftests/
#   gpurun -- 'bash scripts/gpu_reftests.sh'; rm -rf _reftests        (_reftests/ is git-ignored)
# Result of the last run: profiles/r01_reference_tests_unmodified.txt
mkdir -p gpurun_out
cd _reftests
for t in test_batch test_algebras test_pga test_cga test_sta test_basic; do
  timeout 900 python -m pytest $t.py -q --no-header -p no:cacheprovider --maxfail=2000 --tb=line -rf 2>&1 | tail -400 > ../gpurun_out/ref_$t.log
  echo "== $t: $(tail -1 ../gpurun_out/ref_$t.log)"
done
