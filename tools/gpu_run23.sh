#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v "Warning\|^  " | tail -15
