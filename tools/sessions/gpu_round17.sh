#!/bin/bash
SECONDS=0; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2; echo "seconds: $SECONDS"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
