#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
