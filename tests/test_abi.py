"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/kmcp_gpu.h
declares, refuses to run without a device (no CPU fallback), and its host-only helpers agree with the oracle."""
import ctypes as C
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "kmcp_gpu.h")).read()
    return sorted(set(re.findall(r"\b(kmcpg_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from kmcp_b200 import api
    L = api.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libkmcp_gpu.so does not export " + s
    assert set(syms) == set(api.ABI_SYMBOLS)
    assert L.kmcpg_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kmcp_b200 import api
    with pytest.raises(api.KmcpGpuError) as e:
        api.Context(0)
    assert e.value.code == api.KMCPG_ECUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under kmcp_b200/ may import, link, dlopen or call it"""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|kmcp_oracle|libkmcp_oracle|\bko_[a-z_]+\s*\()")
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "kmcp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), os.path.join(dirpath, f)


def test_host_fpr_matches_oracle_and_golden(oracle):
    from kmcp_b200 import api
    L = api.load()
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "fpr.json"))):
        assert float(L.kmcpg_query_fpr(c["n"], c["c"], c["p"])).hex() == c["fpr_hex"]
    for n, c in ((130, 72), (77, 40), (300, 200), (1500, 900)):
        assert L.kmcpg_query_fpr(n, c, 0.3) == oracle.query_fpr(n, c, 0.3)


def test_default_params_are_the_reference_defaults():
    from kmcp_b200 import api
    L = api.load()
    p = api.SearchParams()
    L.kmcpg_default_params(C.byref(p))
    assert (p.min_query_len, p.min_matched, p.dedup_threshold, p.min_query_cov, p.paired) == (30, 10, 256, 0.55, 0)   # S:1055-1069
    o = api.EngineOpts()
    L.kmcpg_default_engine_opts(C.byref(o))
    assert (o.min_query_cov, o.min_target_cov, o.max_fpr, o.top_n_scores, o.sort_by) == (0.55, 0.0, 0.01, 0, 0)          # S:1066-1093


def test_shard_plan_reads_headers_and_reports_format_errors(oracle, tmp_path):
    """host-only part of kmcpg_open_db: __db.yml + .uniki header parsing, error codes of X:38-56 / util-db-info.go:118"""
    import parity_helpers as helpers
    from kmcp_b200 import api
    O = oracle
    sp = O.sketch_params(21)
    targets = helpers.make_synth_targets(O, sp, 5, 6, 6000, 2, 50)
    r001 = O.build_db(targets, str(tmp_path / "db"), sp, num_hashes=1, fpr=0.3, block_size=8)       # 12 targets → 2 blocks
    assert sorted(api.shard_plan(r001, 2)) == [0, 1]
    blk = os.path.join(r001, "_block001.uniki")
    raw = open(blk, "rb").read()
    cases = {"bad magic": (b"XXXXXXXX" + raw[8:], api.KMCPG_EFORMAT), "version": (raw[:8] + bytes([3]) + raw[9:], api.KMCPG_EFORMAT),
             "truncated rows": (raw[:-100], api.KMCPG_EIO), "truncated header": (raw[:30], api.KMCPG_EIO)}
    for name, (data, code) in cases.items():
        open(blk, "wb").write(data)
        with pytest.raises(api.KmcpGpuError) as e:
            api.shard_plan(r001, 2)
        assert e.value.code == code, name
    open(blk, "wb").write(raw)
    yml = os.path.join(r001, "__db.yml")
    txt = open(yml).read()
    open(yml, "w").write(txt.replace("version: 4", "version: 3", 1))
    with pytest.raises(api.KmcpGpuError) as e:
        api.shard_plan(r001, 2)
    assert e.value.code == api.KMCPG_EFORMAT
    open(yml, "w").write(txt.replace("hashes: 1", "hashes: 2"))          # blocks say 1 hash: incompatible (U:689-695)
    with pytest.raises(api.KmcpGpuError) as e:
        api.shard_plan(r001, 2)
    assert e.value.code == api.KMCPG_EFORMAT
    with pytest.raises(api.KmcpGpuError) as e:
        api.shard_plan(str(tmp_path / "nope"), 2)
    assert e.value.code == api.KMCPG_EIO
