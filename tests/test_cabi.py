"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
that include/etpgt_b200.h declares; the ctypes prototypes cover exactly that set; the host
package mirrors the reference's module API and state-dict keys.  No compute calls (no GPU)."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "etpgt_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(etpgt_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from etpgt_b200 import _lib

    if not _lib.lib_path().exists():
        import importlib.util

        spec = importlib.util.spec_from_file_location("etpgt_build", ROOT / "gat-recommendation_b200" / "build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    handle = ctypes.CDLL(str(lib.lib_path()))
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(lib.exported_symbols()) == names, "ctypes prototypes and header disagree"
    assert lib.load().etpgt_version() >= 100
    assert lib.size("etpgt_tconv_bwd_workspace_bytes", 10, 20, 256, 2) > 10 * 256 * 4


def test_argument_checks_fail_loudly_without_launching(lib):
    # unsupported dim: rejected on the host, nothing is launched, message available
    with pytest.raises(RuntimeError, match="unsupported"):
        lib.call("etpgt_embed_pe_fwd", None, 4, None, 10, None, 0, None, None, 0, 48, None, None)
    with pytest.raises(ValueError, match="Unknown readout type"):
        lib.call("etpgt_readout_fwd", None, None, 1, 32, 9, None, None, None, None)


def test_cpu_tensors_are_rejected():
    from etpgt_b200.model import create_graph_transformer_optimized

    model = create_graph_transformer_optimized(num_items=50, embedding_dim=32, hidden_dim=32, use_laplacian_pe=False)

    class B:
        x = torch.tensor([1, 2, 3])
        edge_index = torch.tensor([[0, 1], [1, 2]])
        batch = torch.tensor([0, 0, 0])

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(B())


def test_module_api_and_state_dict_keys_match_reference():
    from etpgt_b200.model import (create_gat, create_graph_transformer, create_graph_transformer_optimized,
                                  create_graphsage)
    from etpgt_b200.train.losses import BPRLoss, DualLoss, ListwiseLoss, SampledSoftmaxLoss, create_loss_function

    m = create_graph_transformer_optimized(num_items=100, embedding_dim=32, hidden_dim=32)
    assert (m.use_ffn, m.num_layers, m.num_heads) == (False, 2, 2)       # tests/test_models.py:147-149
    keys = set(m.state_dict().keys())
    for layer in (0, 1):
        for lin in ("key", "query", "value", "skip"):
            assert f"convs.{layer}.lin_{lin}.weight" in keys and f"convs.{layer}.lin_{lin}.bias" in keys
        assert f"convs.{layer}.lin_beta.weight" in keys
        assert f"batch_norms.{layer}.running_var" in keys
    assert {"item_embedding.weight", "laplacian_pe.projection.weight", "laplacian_pe.projection.bias"} <= keys
    assert m.item_embedding.padding_idx == 0 and bool((m.item_embedding.weight[0] == 0).all())
    with pytest.raises(RuntimeError, match="Laplacian PE not precomputed"):
        m.laplacian_pe(torch.tensor([1, 2]))
    count = lambda mod: sum(p.numel() for p in mod.parameters())  # noqa: E731
    kw = dict(num_items=188, embedding_dim=64, hidden_dim=64, num_layers=2, dropout=0.1)
    # docs/EXPERIMENTS.md:85-88
    assert count(create_graph_transformer_optimized(**kw, num_heads=2, use_laplacian_pe=False)) == 45952
    assert count(create_graph_transformer(**kw, num_heads=2, use_laplacian_pe=False, use_ffn=True)) == 112128
    assert count(create_gat(**kw, num_heads=2)) == 29312
    assert count(create_graphsage(**kw)) == 28800
    assert isinstance(create_loss_function("bpr"), BPRLoss)
    assert isinstance(create_loss_function("listwise"), ListwiseLoss)
    assert isinstance(create_loss_function("dual"), DualLoss)
    assert isinstance(create_loss_function("sampled_softmax"), SampledSoftmaxLoss)
    with pytest.raises(ValueError, match="Unknown loss type"):
        create_loss_function("nope")


def test_golden_state_dicts_load_strictly():
    from golden_util import Golden

    from etpgt_b200.model import create_gat, create_graph_transformer_optimized, create_graphsage

    for name, factory in (("gt_opt_b32", create_graph_transformer_optimized), ("gat_directed", create_gat),
                          ("sage_directed", create_graphsage)):
        g = Golden(name)
        cfg = g.cfg()
        model = factory(**cfg)
        state = g.group("state")
        pe = state.pop("laplacian_pe._cached_pe", None)
        missing, unexpected = model.load_state_dict(state, strict=False)
        assert not unexpected, unexpected
        assert not missing, missing


def test_step_descriptor_layout_matches_the_header(tmp_path):
    """etpgt_gt_step_t / etpgt_gt_layer_t are passed by address from ctypes: every field of the ctypes mirror
    (etpgt_b200/train/step.py) must sit at the offset the C compiler gives it in include/etpgt_b200.h."""
    import subprocess

    from etpgt_b200.train import step

    def fields(struct):
        return [name for name, _ in struct._fields_]

    lines = ['#include <stddef.h>', '#include <stdio.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, struct in (("etpgt_gt_layer_t", step._GtLayer), ("etpgt_gt_step_t", step._GtStep)):
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields(struct):
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    for cname, struct in (("etpgt_gt_layer_t", step._GtLayer), ("etpgt_gt_step_t", step._GtStep)):
        assert int(got[cname]) == ctypes.sizeof(struct)
        for f in fields(struct):
            assert int(got[f"{cname}.{f}"]) == getattr(struct, f).offset, f"{cname}.{f}"
    # and the header declares no field the mirror lacks
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    body = text[text.index("typedef struct etpgt_gt_step {"):text.index("} etpgt_gt_step_t;")]
    declared = set(re.findall(r"[\s\*,](\w+)\s*(?:\[\w+\])?\s*[;,]", body))
    assert declared == set(fields(step._GtStep)), declared ^ set(fields(step._GtStep))
    assert step.MAX_LAYERS == int(re.search(r"#define ETPGT_GT_MAX_LAYERS (\d+)", text).group(1))


def test_step_phase_exchange_rows():
    """Host-side control flow of the data-parallel step (etpgt_b200/train/step.py): which BatchNorm exchange row
    each driver phase produces.  Forward rows in layer order, backward rows from the last layer down, none for the
    two trailing phases; every row exactly once."""
    from etpgt_b200.train.step import exchange_row

    for layers in (1, 2, 3, 4):
        rows = [exchange_row(p, layers) for p in range(2 * layers + 2)]
        assert rows[:layers] == list(range(layers))
        assert rows[layers:2 * layers] == [layers + l for l in reversed(range(layers))]
        assert rows[2 * layers:] == [None, None]
        assert sorted(r for r in rows if r is not None) == list(range(2 * layers))
    assert exchange_row(-1, 2) is None and exchange_row(99, 2) is None


def test_step_driver_support_matrix_on_cpu_models():
    """Which (model, loss) pairs the trainer sends through the C++ step driver — decided on the host without
    touching a GPU: CPU-resident models are never driven (the driver has no CPU fallback), nor are the FFN variant,
    the attention readout, GAT / GraphSAGE."""
    from etpgt_b200.model import (create_gat, create_graph_transformer, create_graph_transformer_optimized,
                                  create_graphsage)
    from etpgt_b200.train.step import FusedTrainStep

    reason = FusedTrainStep.unsupported_reason
    assert "CUDA" in reason(create_graph_transformer_optimized(50, 64, 64))
    assert "use_ffn" in reason(create_graph_transformer(50, 64, 64))
    assert "readout" in reason(create_graph_transformer_optimized(50, 64, 64, readout_type="attention"))
    assert "GraphTransformer" in reason(create_gat(50, 64, 64))
    assert "GraphTransformer" in reason(create_graphsage(50, 64, 64))
    assert "hidden_dim" in reason(create_graph_transformer_optimized(50, 48, 48)) or \
        "embedding_dim" in reason(create_graph_transformer_optimized(50, 48, 48))
    with pytest.raises(NotImplementedError):
        FusedTrainStep(create_graph_transformer_optimized(50, 64, 64))     # CPU parameters
