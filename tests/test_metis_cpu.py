"""CPU: the memory-mapped, parallel METIS reader (gnn-mwvc_b200/host/gvc_metis.cpp)
against the reference's own parse_graph (src/GNN_VC.cpp:34-91, compiled unmodified into oracle/_ref):
well-formed files and the quirks of the reference's reader."""
import numpy as np
import pytest

import gnn_mwvc_b200  # noqa: F401
from gnn_mwvc_b200 import capi, graphs


def write(path, text):
    path.write_text(text)
    return path


def same(ref_out, our_out, what):
    n, w, eu, ev = ref_out
    n2, w2, eu2, ev2 = our_out
    assert n == n2, what
    assert np.array_equal(w, w2), what
    assert np.array_equal(eu, eu2) and np.array_equal(ev, ev2), what


def test_generated_graphs(reference, tmp_path):
    for g in (graphs.er10k_fixture(), graphs.rmat_graph(12, 16, seed=3), graphs.grid_graph(37, 41),
              graphs.er_graph(200_000, 1_000_000, seed=9)):                    # the last one is parsed by several threads
        p = tmp_path / f"{g.name}.graph"
        graphs.write_metis(g, p)
        for threads in (1, 0, 5):
            ours = capi.parse_metis(p, threads)
            same(reference.parse_graph(p), ours, f"{g.name} threads={threads}")
        n, w, eu, ev, (rp, col, nw) = capi.parse_metis(p, 0, csr=True)
        rp0, col0, W0, NW0 = g.numpy()
        assert np.array_equal(rp, rp0) and np.array_equal(col, col0) and np.array_equal(nw, NW0) and np.array_equal(w, W0)


def test_quirks_of_the_reference_reader(reference, tmp_path):
    cases = {
        "readme": "3 2 10\n15 3\n15 3\n20 1 2\n",
        "header_extras_and_crlf": "4 3 10 junk 7\r\n5 2 3\r\n6 1 4\r\n7 1\r\n8 2\r\n",
        "unsorted_neighbours": "4 4 10\n1 4 3 2\n2 1 3\n3 2 1\n4 1\n",
        "header_e_too_large": "3 5 10\n1 2\n2 1 3\n3 2\n",                      # unused slots become one (0,0) self-loop
        "missing_last_lines": "5 2 10\n9 2\n8 1 3\n",
        "blank_line_vertex": "3 1 10\n4 3\n\n6 1\n",
        "garbage_stops_the_line": "3 2 10\n4 2 x 3\n5 1 3\n6 1 2\n",
        "no_trailing_newline": "2 1 10\n1 2\n2 1",
        "isolated_vertices": "4 1 10\n1\n2 3\n3 2\n4\n",
        "plus_signs_and_tabs": "2 1 10\n+7\t+2\n3 1\n",
        "empty": "0 0 10\n",
    }
    for name, text in cases.items():
        p = write(tmp_path / f"{name}.graph", text)
        same(reference.parse_graph(p), capi.parse_metis(p), name)


def test_undefined_behaviour_of_the_reference_is_an_error_here(tmp_path):
    p = write(tmp_path / "too_many.graph", "3 1 10\n1 2 3\n2 1 3\n3 1 2\n")       # 3 edges, header says 1
    with pytest.raises(capi.GvcError, match="more edges"):
        capi.parse_metis(p)
    p = write(tmp_path / "out_of_range.graph", "2 1 10\n1 3\n2 1\n")
    with pytest.raises(capi.GvcError, match="neighbour id"):
        capi.parse_metis(p)
    with pytest.raises(capi.GvcError, match="cannot open"):
        capi.parse_metis(tmp_path / "nope.graph")
