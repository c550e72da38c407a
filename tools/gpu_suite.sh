#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/suite.log
: > $LOG
echo "=== full gpu suite" >> $LOG
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -n 25 >> $LOG
echo "=== smoke" >> $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
tail -n 40 $LOG
