#!/bin/bash
# full GPU suite + smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -60 > gpurun_out/all_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/all_tests.log 2>&1
tail -45 gpurun_out/all_tests.log
