"""Write the col array (uint32 neighbour ids) of the R-MAT scale-20 benchmark graph to a file, for gather_bw."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import gnn_mwvc_b200  # noqa
from gnn_mwvc_b200 import graphs
g = graphs.rmat_graph(20, 16, seed=42, device="cuda")
g.col.cpu().numpy().tofile(sys.argv[1])
print("wrote", g.col.numel(), "ids; max degree", int((g.row_ptr[1:] - g.row_ptr[:-1]).max()))
