#!/bin/bash
# First-call costs of the drop-in on a fresh process (GVC_TRACE + GVC_PROFILE), ER graph of $1 vertices.
n=${1:-1000000}
python - <<PY
import sys; sys.path.insert(0, ".")
import gnn_mwvc_b200
from gnn_mwvc_b200 import graphs
g = graphs.er_graph($n, 5 * $n, seed=1)
graphs.write_metis(g, "/tmp/er.graph")
PY
for bin in "$@"; do
  [ "$bin" = "$n" ] && continue
  echo "== $bin"
  GVC_TRACE=1 GVC_PROFILE=1 $bin /tmp/er.graph /tmp/out.txt 0 -1 0 2>&1 | grep "gvc" | head -${LINES_MAX:-40}
done
