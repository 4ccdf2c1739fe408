"""The C-ABI library loads, exports every symbol include/leann_b200.h declares, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "leann_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lb2_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("lb2_open", "lb2_search", "lb2_search_device", "lb2_close", "lb2_last_error", "lb2_set_passages",
                 "lb2_set_encoder", "lb2_set_vectors", "lb2_encode_ids", "lb2_encode_tokens"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/leann_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from leann_b200 import capi
    assert sorted(capi.EXPORTED_SYMBOLS) == declared_symbols()


def test_no_torch_or_cxx_types_in_signatures():
    text = (ROOT / "include" / "leann_b200.h").read_text()
    assert "std::" not in text and "at::" not in text and "torch" not in text.replace("no C++ or torch types", "")


def test_struct_layouts_match_header(lib):
    from leann_b200 import capi
    assert ctypes.sizeof(capi.SearchParams) == 32
    assert ctypes.sizeof(capi.EncoderConfig) == 40
    assert ctypes.sizeof(capi.SearchStats) == 7 * 8 + 6 * 8 + 8  # + n_encoder_passes
    assert ctypes.sizeof(capi.IndexInfo) == 56
    p = capi.SearchParams()
    lib.lb2_default_params(ctypes.byref(p))
    assert (p.efSearch, p.beam_size, p.batch_size, p.check_relative_distance, p.recompute) == (64, 1, 0, 1, 1)
    # DiskANN structs: 8 x int32 + float + uint32 ; int64 + 8 x int32 + float + int32 + int64
    assert ctypes.sizeof(capi.DiskannParams) == 40 and ctypes.sizeof(capi.DiskannInfo) == 56
    d = capi.DiskannParams()
    lib.lb2_diskann_default_params(ctypes.byref(d))
    assert (d.complexity, d.beam_width, d.deferred_fetch, d.skip_search_reorder, d.recompute_neighbors, d.io_limit) == (64, 1, 1, 0, 0, 0)
    m = capi.make_diskann_params(48, 3, recompute_embeddings=False, skip_search_reorder=True, prune_ratio=0.25, io_limit=9)
    assert (m.complexity, m.beam_width, m.deferred_fetch, m.skip_search_reorder, m.io_limit) == (48, 3, 0, 1, 9) and abs(m.prune_ratio - 0.25) < 1e-7


def test_weight_count_is_pure_host_arithmetic(lib):
    from leann_b200 import synth
    for preset in (synth.MINILM_L6, synth.BGE_BASE, synth.TINY):
        cfg = preset.config()
        n = lib.lb2_encoder_weight_count(ctypes.byref(cfg))
        assert n == sum(int.__mul__(*s) if len(s) == 2 else s[0] for _, s in synth.weight_layout(preset))


def test_open_fails_loudly_without_gpu(lib, golden_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from leann_b200 import capi
    with pytest.raises(capi.Lb2Error, match="no CUDA device|no CPU path"):
        capi.Index(str(golden_dir / "hnsw_small_ip.index"), 0)


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    from leann_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "_LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU path"):
        capi.load()


def test_product_never_imports_the_oracle():
    for f in (ROOT / "leann_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and f.is_file():
            t = f.read_text()
            assert "import oracle" not in t and "from oracle" not in t and "liboracle" not in t, f
            assert '#include "../oracle' not in t and "oracle/" not in t.replace("oracle/canon_dist.h", ""), f
