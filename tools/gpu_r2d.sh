#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
python tools/debug_cf.py > $O/debug_cf.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
echo done > $O/done
