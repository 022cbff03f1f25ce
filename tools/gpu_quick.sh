#!/bin/bash
# scratch: quick sanity of the final build
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_graph.py tests/test_gpu_ops.py -q -x --timeout 120 2>&1 | tail -2
