"""The marching-cubes tables (DATA) in both layouts, checked value by value against the reference's Go literals when
/root/reference is present (build container), and against each other always."""
import importlib.util
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/glrender/marchcubes.go"


def _gen():
    spec = importlib.util.spec_from_file_location("gen_mc_tables", os.path.join(ROOT, "oracle", "tools", "gen_mc_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _cuh_tables():
    src = open(os.path.join(ROOT, "gsdf_b200", "csrc", "mc_tables.cuh")).read()
    def arr(name):
        body = re.search(r"%s\[[^\]]*\]\s*=\s*\{(.*?)\};" % name, src, re.S).group(1)
        return [int(x, 0) for x in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)", body)]
    return arr("c_mc_edges"), arr("c_mc_ntri"), arr("c_mc_tris"), arr("c_mc_pairs")


def test_oracle_and_cuda_tables_agree(oracle):
    L = oracle.lib()
    edges = [L.go_mc_edge_table()[i] for i in range(256)]
    tris = [L.go_mc_tri_table()[i] for i in range(256 * 16)]
    pairs = [L.go_mc_pair_table()[i] for i in range(24)]
    ce, cn, ct, cp = _cuh_tables()
    assert edges == ce and tris == ct and pairs == cp
    for i in range(256):
        row = tris[16 * i:16 * i + 16]
        n = sum(1 for v in row if v >= 0)
        assert n % 3 == 0 and n // 3 == cn[i] <= 5
        assert all(v == -1 for v in row[n:])
        used = 0
        for v in row[:n]:
            used |= 1 << v
        assert used == edges[i]  # the edge mask is exactly the set of edges the triangles use
    assert edges[0] == 0 and edges[255] == 0 and cn[0] == 0 and cn[255] == 0


def test_emit_kernel_pair_table_matches():
    """mc_kernels.cuh packs the 12 edge->corner pairs 4 bits per edge into two 64-bit literals."""
    src = open(os.path.join(ROOT, "gsdf_b200", "csrc", "mc_kernels.cuh")).read()
    m = re.search(r"\(0x([0-9a-f]+)ull >> \(4 \* e\)\) & 0xf\), cb = \(int\)\(\(0x([0-9a-f]+)ull >> \(4 \* e\)\)", src)
    assert m
    pa, pb = int(m.group(1), 16), int(m.group(2), 16)
    _, _, _, cp = _cuh_tables()
    assert [v for e in range(12) for v in ((pa >> (4 * e)) & 0xf, (pb >> (4 * e)) & 0xf)] == cp


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_tables_equal_reference_literals(oracle):
    pairs, edges, tris = _gen().parse_reference()
    L = oracle.lib()
    assert [L.go_mc_edge_table()[i] for i in range(256)] == edges
    assert [L.go_mc_pair_table()[i] for i in range(24)] == [v for p in pairs for v in p]
    flat = [v for t in tris for v in t + [-1] * (16 - len(t))]
    assert [L.go_mc_tri_table()[i] for i in range(4096)] == flat
