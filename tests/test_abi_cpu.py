"""CPU: the C-ABI library loads, exports every symbol include/rtkv.h declares, the ctypes
mirrors of its structs have the C layout, and the host-side logic of the drop-in classes."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_header_symbols_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "rtkv.h")).read()
    declared = set(re.findall(r"RTKV_API\s+[\w\s\*]+?\b(rtkv_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    from rtkv import _lib
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert _lib.lib.rtkv_abi_version() == 1


def test_struct_layout_matches_c(built_lib):
    """sizeof/offsetof from a C translation unit that includes the public header."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rtkv.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu\n", sizeof(rtkv_layer_params), offsetof(rtkv_layer_params,bits),
        offsetof(rtkv_layer_params,fallback_topk), sizeof(rtkv_row_stats), offsetof(rtkv_row_stats,cost_bits),
        sizeof(rtkv_layer_stats));
 printf("%zu %zu %zu\n", offsetof(rtkv_row_stats,sel_score_sum), offsetof(rtkv_row_stats,score_min), offsetof(rtkv_row_stats,label_count));
 printf("%zu %zu %zu %zu\n", sizeof(rtkv_quant_options), offsetof(rtkv_quant_options,group_size),
        offsetof(rtkv_quant_options,token_group), offsetof(rtkv_quant_options,outlier_tau_q8));
 return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    from rtkv import _lib
    vals = [int(v) for v in out]
    assert vals[0] == ctypes.sizeof(_lib.LayerParams) == 80
    assert vals[1] == _lib.LayerParams.bits.offset and vals[2] == _lib.LayerParams.fallback_topk.offset
    assert vals[3] == _lib.ROW_STATS.itemsize and vals[4] == _lib.ROW_STATS.fields["cost_bits"][1]
    assert vals[5] == _lib.LAYER_STATS.itemsize == ctypes.sizeof(_lib.LayerStats)
    assert vals[6] == _lib.ROW_STATS.fields["sel_score_sum"][1]
    assert vals[7] == _lib.ROW_STATS.fields["score_min"][1]
    assert vals[8] == _lib.ROW_STATS.fields["label_count"][1]
    assert vals[9] == ctypes.sizeof(_lib.QuantOptions) == 32
    assert vals[10] == _lib.QuantOptions.group_size.offset and vals[11] == _lib.QuantOptions.token_group.offset
    assert vals[12] == _lib.QuantOptions.outlier_tau_q8.offset


def test_argument_errors_do_not_need_a_gpu(built_lib):
    from rtkv import _lib
    assert _lib.lib.rtkv_packed_row_bytes(4096, 2) == 1024
    assert _lib.lib.rtkv_packed_row_bytes(4096, 16) == 8192
    assert _lib.lib.rtkv_packed_row_bytes(10, 4) == 16
    assert _lib.lib.rtkv_packed_row_bytes(10, 3 + 14) == -1
    rc = _lib.lib.rtkv_prompt_mass(None, 0, 1, 1, 1, 0, 0, 0, None, 1, None, None)
    assert rc == -1 and b"null pointer" in _lib.lib.rtkv_last_error()
    with pytest.raises(_lib.RtkvError):
        _lib.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu(built_lib):
    import rtkv
    cfg = rtkv.CompressionConfig()
    comp = rtkv.RealTimePrefillCompressor(cfg)
    k = torch.zeros(1, 10, 16)
    a = torch.softmax(torch.zeros(1, 2, 10, 10), -1)
    with pytest.raises(RuntimeError, match="CUDA"):
        comp.compress_layer_kv_cache(k, k, a, torch.zeros(1, 10, dtype=torch.long), 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        rtkv.PromptGuidedImportanceScorer(cfg).compute_importance_scores(a, torch.arange(2), 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        rtkv.DynamicPrecisionQuantizer(cfg).assign_precision_levels(torch.rand(1, 10))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "realtime-kv-cache-compression_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "librtkv_oracle" not in text, f


def test_config_defaults_and_layer_weights(built_lib):
    import rtkv
    c = rtkv.CompressionConfig()
    assert (c.alpha, c.beta, c.gamma, c.theta_h, c.theta_m) == (0.4, 0.3, 0.3, 0.7, 0.3)
    assert (c.high_precision_bits, c.medium_precision_bits, c.low_precision_bits) == (16, 8, 4)
    assert (c.early_layer_ratio, c.middle_layer_ratio, c.later_layer_ratio) == (0.8, 0.6, 0.4)
    assert len(c.layer_weights) == 32 and c.layer_weights[0] == 1.0 and c.layer_weights[-1] == 0.5
    assert c.layer_weights[5] == 1.0 - 0.5 * (5 / 31)
    assert c.context_lengths == [4096, 8192, 16384, 32768] and c.batch_sizes == [1, 4, 8]
    c40 = rtkv.CompressionConfig(num_hidden_layers=40, num_attention_heads=40, hidden_size=5120)
    assert len(c40.layer_weights) == 40


def test_layer_bands_and_prompt_length(built_lib):
    import rtkv
    p = rtkv.SelectiveTokenPropagator(rtkv.CompressionConfig())
    assert [p.get_layer_propagation_ratio(l) for l in (0, 8, 9, 21, 22, 31)] == [0.8, 0.8, 0.6, 0.6, 0.4, 0.4]
    p40 = rtkv.SelectiveTokenPropagator(rtkv.CompressionConfig(num_hidden_layers=40))
    assert [p40.get_layer_propagation_ratio(l) for l in (11, 12, 27, 28)] == [0.8, 0.6, 0.6, 0.4]
    e = p.estimate_compression_ratio(2, 1000)
    assert abs(e["cumulative_ratio"] - 0.8 ** 3) < 1e-12 and e["estimated_length"] == 512
    comp = rtkv.RealTimePrefillCompressor(rtkv.CompressionConfig())
    assert [comp.prompt_length(s) for s in (1, 4, 10, 640, 4096, 32768)] == [1, 1, 2, 128, 128, 128]
    ids = torch.zeros(2, 50, dtype=torch.long)
    assert torch.equal(comp.identify_prompt_tokens(ids), torch.arange(10))
    assert comp.get_overall_compression_stats() == {}
    assert rtkv.UnifiedCompressor is rtkv.RealTimePrefillCompressor
    assert rtkv.UnifiedCompressor.compress is rtkv.UnifiedCompressor.compress_layer_kv_cache


def test_selection_info_from_rows(built_lib):
    from rtkv import _lib
    from rtkv.selective_propagation import selection_info_from_rows
    rows = np.zeros(3, _lib.ROW_STATS)
    rows[0]["n_selected_greedy"], rows[0]["cost_bits"], rows[0]["sel_score_sum"] = 4, 16, 2.0
    rows[0]["sel_class"] = (1, 2, 1)
    rows[2]["n_selected_greedy"], rows[2]["cost_bits"], rows[2]["sel_score_sum"] = 2, 8, 0.5
    rows[2]["sel_class"] = (0, 0, 2)
    info = selection_info_from_rows(rows, 4.0)
    assert info["selected_counts"] == [4, 2] and info["budget_utilization"] == [0.5, 0.25]
    assert info["avg_importance"] == [0.5, 0.25]
    assert info["cost_distribution"] == {"high": 3, "medium": 2, "low": 1}
    assert info["avg_selected"] == 3 and info["avg_budget_util"] == 0.375
    empty = selection_info_from_rows(rows[1:2], 4.0)
    assert "avg_selected" not in empty and empty["selected_counts"] == []


def test_tracker_cumulative_scores_host_logic(built_lib):
    import rtkv
    t = rtkv.LayerWiseImportanceTracker(rtkv.CompressionConfig(num_hidden_layers=8))
    assert t.get_cumulative_scores(0) is None
    for l in range(3):
        t.layer_scores[l] = torch.full((1, 5), float(l + 1))
    assert torch.allclose(t.get_cumulative_scores(2), torch.full((1, 5), 2.0))
    assert torch.allclose(t.get_cumulative_scores(5), torch.full((1, 5), 1.0))   # divides by layer_idx+1
    t.layer_scores = {3: torch.ones(1, 5)}
    with pytest.raises(KeyError):
        t.get_cumulative_scores(3)                                              # reference: layer_scores[0]


def test_compat_tree_resolves_reference_imports(built_lib):
    """PYTHONPATH=<pkg>/compat:<pkg> makes the reference's own import lines resolve to rtkv."""
    import subprocess, sys
    pkg = os.path.join(ROOT, "realtime-kv-cache-compression_b200")
    code = ("from configs.base_config import CompressionConfig\n"
            "from src.compression.unified_compressor import RealTimePrefillCompressor\n"
            "from src.compression.token_importance import PromptGuidedImportanceScorer, LayerWiseImportanceTracker\n"
            "from src.compression.dynamic_quantization import DynamicPrecisionQuantizer\n"
            "from src.compression.selective_propagation import SelectiveTokenPropagator\n"
            "from src.models.modified_llama import create_compressed_llama_model, CompressedLlamaForCausalLM\n"
            "import rtkv\n"
            "assert RealTimePrefillCompressor is rtkv.RealTimePrefillCompressor\n"
            "c = RealTimePrefillCompressor(CompressionConfig(num_hidden_layers=4))\n"
            "print('ok', len(c.config.layer_weights))\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(pkg, "compat") + os.pathsep + pkg)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip() == "ok 4", out.stderr


def test_model_hook_chain_on_cpu(built_lib):
    """set_compressor reaches every attention module; a decode step (q_len == 1) and training mode
    never call the compressor (modified_llama.py:102), so this runs without a GPU."""
    import rtkv
    from rtkv.llama_hooks import CompressedLlamaForCausalLM
    from transformers import LlamaConfig
    hf = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                     num_key_value_heads=4, vocab_size=97, max_position_embeddings=64)
    model = CompressedLlamaForCausalLM(hf).eval()
    comp = rtkv.RealTimePrefillCompressor(rtkv.CompressionConfig(num_hidden_layers=2, num_attention_heads=4, hidden_size=64))
    model.set_compressor(comp)
    assert all(l.self_attn.compressor is comp for l in model.model.layers)
    with torch.no_grad():
        out = model(input_ids=torch.tensor([[5]]))
    assert out.logits.shape == (1, 1, 97) and out.compression_info == {}
    assert model.get_compression_stats() == {}
    model.reset_compression_state()
