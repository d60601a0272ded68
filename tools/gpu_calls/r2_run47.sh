#!/bin/bash
# last check of the round's final build: GPU suite + smoke
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
