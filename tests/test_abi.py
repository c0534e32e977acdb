"""CPU-side boundary checks: the C-ABI library loads and exports every symbol include/scn_gpu.h
declares; the ctypes table matches the header; the host-side mirror validates like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "scn_gpu.h")).read()
    return sorted(set(re.findall(r"SCN_API\s+[\w\s\*]+?\b(scn_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("scn_store_create", "scn_store_append", "scn_store_mark_deleted", "scn_graph_upload",
                 "scn_search_flat", "scn_search_hnsw", "scn_rerank", "scn_distance_batch", "scn_vector_ops", "scn_merge_topk_dev",
                 "scn_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__

    __graft_entry__.build()
    from scintirete_b200 import _native

    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in scn_gpu.h but not exported"
    assert sorted(_native.SIGNATURES) == _declared()


def test_no_product_code_touches_the_oracle():
    pkg = os.path.join(ROOT, "scintirete_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "scn_oracle" not in src, f


def test_metric_validation_matches_new_distance_calculator():
    from scintirete_b200 import ScintireteError, new_distance_calculator

    for bad in (0, 999):
        with pytest.raises(ScintireteError) as e:
            new_distance_calculator(bad)
        assert e.value.code == 3007
    for ok in (1, 2, 3):
        c = new_distance_calculator(ok)
        assert int(c.distance_type()) == ok and c.is_similarity() is False
    # mismatched dimensions: +Inf without touching the device (distance.go:22-24)
    assert new_distance_calculator(1).distance([1, 2], [1, 2, 3]) == np.inf


def test_key_encoding_is_order_preserving():
    # mirrors scn::f32_ord (common.cuh): ascending u32 order == ascending float order, -0 == +0
    def ord32(x):
        x = np.float32(x) + np.float32(0)
        u = np.float32(x).view(np.uint32)
        return np.uint32(~u) if u & 0x80000000 else np.uint32(u | 0x80000000)

    xs = np.array([-np.inf, -3.5, -1e-30, -0.0, 0.0, 1e-30, 2.0, np.inf], np.float32)
    o = [int(ord32(x)) for x in xs]
    assert o == sorted(o) and o[3] == o[4]


def test_go_shim_binds_only_declared_entry_points_with_the_declared_arity():
    # go/gpuindex/gpuindex.go cannot be compiled here (no Go toolchain): at least every C.scn_* call
    # in it must name an entry point of include/scn_gpu.h and pass as many arguments as it declares
    header = open(os.path.join(ROOT, "include", "scn_gpu.h")).read()
    go = open(os.path.join(ROOT, "go", "gpuindex", "gpuindex.go")).read()

    def split_args(text, start):          # text[start] == '(' -> top-level comma count of the call
        depth, n, i, any_char = 0, 0, start, False
        while i < len(text):
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    return n + (1 if any_char else 0)
            elif ch == "," and depth == 1:
                n += 1
            elif depth >= 1 and not ch.isspace():
                any_char = True
            i += 1
        raise AssertionError("unbalanced call")

    declared = {}
    header_nc = re.sub(r"/\*.*?\*/", "", header, flags=re.S)   # arity without the comments inside signatures
    for m in re.finditer(r"SCN_API\s+[\w\s\*]+?\b(scn_\w+)\s*\(", header_nc):
        args = header_nc[m.end():header_nc.index(")", m.end())]
        declared[m.group(1)] = 0 if args.strip() in ("", "void") else args.count(",") + 1
    calls = list(re.finditer(r"\bC\.(scn_\w+)\s*\(", go))
    assert len(calls) >= 15
    for m in calls:
        name = m.group(1)
        assert name in declared, f"{name} is not declared in scn_gpu.h"
        assert split_args(go, m.end() - 1) == declared[name], f"{name}: argument count differs from the header"
