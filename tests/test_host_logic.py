"""CPU-side checks: the C-ABI library loads and exports every symbol include/kgb200.h declares,
the layer API mirrors the reference's signatures / config keys / error conventions, and the product
refuses to run without a CUDA device (no CPU fallback).  No kernel is launched here."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

import keras_geometric_b200 as kg
from keras_geometric_b200 import _build, _lib
from keras_geometric_b200.layers import AggregatorFactory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kgb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in kgb200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) <= set(names)
    lib.kgb_version.restype = ctypes.c_int
    assert lib.kgb_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    """ctypes mirror of kgb_gather_reduce_args: same field order as the header."""
    src = open(os.path.join(ROOT, "include", "kgb200.h")).read()
    body = src[src.index("typedef struct kgb_gather_reduce_args {"):src.index("} kgb_gather_reduce_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b([a-z_A-Z0-9]+);", body)
    assert fields == [f[0] for f in _lib.GatherReduceArgs._fields_]


def test_sass_has_vector_loads():
    """The hot kernels use 128-bit global loads (LDG.E.128) - checked on the built library."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "code for sm_100a" in sass
    assert sass.count("LDG.E.128") > 100  # float4 feature-row loads in the gather kernels
    # the dense transforms run on the 5th-generation tensor cores: tcgen05 MMAs (also the CTA-pair form),
    # TMA tile loads, TMEM reads, multicast commits
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "LDTM", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    layer = kg.GCNConv(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer([np.zeros((5, 3), np.float32), np.zeros((2, 6), np.int32)])
    with pytest.raises(RuntimeError, match="CUDA"):
        kg.MessagePassing("sum").aggregate(np.ones((3, 2), np.float32), np.zeros(3, np.int32), num_nodes=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        kg.add_self_loops(np.zeros((2, 3), np.int32), 4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "keras_geometric_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


# ---- API surface -------------------------------------------------------------------
def params(fn):
    return list(inspect.signature(fn).parameters)


def test_constructor_signatures():
    assert params(kg.MessagePassing.__init__)[:2] == ["self", "aggregator"]
    assert params(kg.GCNConv.__init__)[1:12] == ["output_dim", "use_bias", "kernel_initializer", "bias_initializer",
                                                 "kernel_regularizer", "bias_regularizer", "kernel_constraint",
                                                 "bias_constraint", "add_self_loops", "normalize", "dropout_rate"]
    assert params(kg.SAGEConv.__init__)[1:16] == ["output_dim", "aggregator", "normalize", "root_weight", "use_bias",
                                                  "activation", "pool_activation", "pool_hidden_dim",
                                                  "kernel_initializer", "bias_initializer", "kernel_regularizer",
                                                  "bias_regularizer", "kernel_constraint", "bias_constraint",
                                                  "dropout_rate"]
    assert params(kg.GINConv.__init__)[1:11] == ["output_dim", "mlp_hidden", "aggregator", "eps_init", "train_eps",
                                                 "use_bias", "dropout", "kernel_initializer", "bias_initializer",
                                                 "activation"]
    assert params(kg.GATv2Conv.__init__)[1:11] == ["output_dim", "heads", "concat", "negative_slope", "dropout",
                                                   "use_bias", "kernel_initializer", "bias_initializer",
                                                   "att_initializer", "add_self_loops"]
    assert params(kg.MessagePassing.message) == ["self", "x_i", "x_j", "edge_attr", "edge_index", "size", "kwargs"]
    assert params(kg.MessagePassing.aggregate) == ["self", "messages", "target_idx", "num_nodes", "dim_size"]
    assert params(kg.MessagePassing.propagate) == ["self", "x", "edge_index", "edge_attr", "size", "kwargs"]
    assert params(kg.MessagePassing.call) == ["self", "inputs", "edge_attr", "training"]
    assert params(kg.GCNConv.call) == ["self", "inputs", "training", "mask"]
    assert params(kg.SAGEConv.call) == ["self", "inputs", "training", "mask"]


def test_invalid_aggregator_errors():
    with pytest.raises(ValueError, match="Invalid aggregator"):
        kg.MessagePassing(aggregator="invalid")
    with pytest.raises(ValueError, match="Invalid aggregator"):
        kg.SAGEConv(4, aggregator="median")
    with pytest.raises(ValueError, match="Invalid aggregator"):
        kg.GINConv(4, aggregator="min")
    assert AggregatorFactory.get_available_aggregators() == ["mean", "max", "sum", "min", "std"]
    for a in ["mean", "max", "sum", "min", "std"]:
        lyr = kg.MessagePassing(aggregator=a)
        assert lyr.aggregator == a and lyr._aggregator.name == a


def test_input_container_errors():
    for layer in (kg.MessagePassing(), kg.GINConv(4)):
        with pytest.raises(ValueError, match="list or tuple"):
            layer.call(np.zeros((3, 2), np.float32))
        with pytest.raises(ValueError, match="at least"):
            layer.call([np.zeros((3, 2), np.float32)])
    with pytest.raises(ValueError, match="GCNConv expects"):
        kg.GCNConv(4).call([np.zeros((3, 2), np.float32)])
    with pytest.raises(ValueError, match="SAGEConv expects"):
        kg.SAGEConv(4).call(np.zeros((3, 2), np.float32))
    with pytest.raises(ValueError, match="Expected inputs"):
        kg.GATv2Conv(4).call([np.zeros((3, 2), np.float32)])


def test_build_weights_names_shapes():
    g = kg.GCNConv(7, use_bias=True)
    g.build([(10, 5), (2, 20)])
    assert tuple(g.kernel.shape) == (5, 7) and tuple(g.bias.shape) == (7,)
    assert [w.keras_name for w in g.weights] == ["kernel", "bias"]
    assert kg.GCNConv(7, use_bias=False).compute_output_shape([(10, 5), (2, 20)]) == (10, 7)
    s = kg.SAGEConv(6, aggregator="pooling", pool_hidden_dim=9)
    s.build([(10, 5), (2, 20)])
    assert tuple(s.lin_neigh.kernel.shape) == (9, 6) and tuple(s.lin_self.kernel.shape) == (5, 6)
    assert tuple(s.pool_mlp.kernel.shape) == (5, 9) and s.lin_neigh.bias is None
    assert s.lin_neigh.name == "linear_neigh" and s.lin_self.name == "linear_self" and s.pool_mlp.name == "pool_mlp"
    assert s.aggregator == "mean" and s.actual_aggregator == "pooling"
    gin = kg.GINConv(4, mlp_hidden=[8, 6], train_eps=True, eps_init=0.3, dropout=0.1)
    gin.build([(10, 5), (2, 20)])
    names = [l.name for l in gin.mlp.layers if hasattr(l, "kernel")]
    assert names == ["mlp_hidden_0", "mlp_hidden_1", "mlp_output"]
    assert tuple(gin.eps.shape) == (1,) and abs(float(gin.eps) - 0.3) < 1e-7
    assert gin.compute_output_shape([(10, 5), (2, 20)]) == (10, 4)
    gat = kg.GATv2Conv(3, heads=4, concat=False)
    gat.build([(10, 5), (2, 20)])
    assert tuple(gat.att.shape) == (1, 4, 3) and tuple(gat.bias.shape) == (3,)
    assert tuple(gat.linear_transform.kernel.shape) == (5, 12) and gat.add_self_loops_flag is True
    gat2 = kg.GATv2Conv(3, heads=4)
    gat2.build((10, 5))
    assert tuple(gat2.bias.shape) == (12,)
    with pytest.raises(ValueError):
        kg.GCNConv(4).build([(10,), (2, 3)])


def test_config_round_trip_keys():
    g = kg.GCNConv(8, add_self_loops=False, normalize=False, dropout_rate=0.25)
    cfg = g.get_config()
    for k in ["output_dim", "use_bias", "kernel_initializer", "bias_initializer", "kernel_regularizer",
              "bias_regularizer", "kernel_constraint", "bias_constraint", "add_self_loops", "normalize",
              "dropout_rate", "aggregator"]:
        assert k in cfg
    assert cfg["aggregator"] == "sum"
    g2 = kg.GCNConv.from_config(cfg)
    assert (g2.output_dim, g2.add_self_loops, g2.normalize, g2.dropout_rate) == (8, False, False, 0.25)
    s = kg.SAGEConv(5, aggregator="pooling", activation="tanh", normalize=True, root_weight=False)
    cfg = s.get_config()
    assert cfg["aggregator"] == "pooling" and cfg["activation"] == "tanh"
    s2 = kg.SAGEConv.from_config(cfg)
    assert s2.actual_aggregator == "pooling" and s2.normalize and not s2.root_weight
    gin = kg.GINConv(5, mlp_hidden=[7], aggregator="mean", eps_init=0.5, train_eps=True, dropout=0.2)
    cfg = gin.get_config()
    assert set(["output_dim", "mlp_hidden", "eps_init", "train_eps", "use_bias", "dropout", "kernel_initializer",
                "bias_initializer", "activation", "aggregator"]) <= set(cfg)
    assert kg.GINConv.from_config(cfg).mlp_hidden == [7]
    gat = kg.GATv2Conv(6, heads=3, concat=False, negative_slope=0.1, dropout=0.3, add_self_loops=False)
    cfg = gat.get_config()
    assert cfg["add_self_loops"] is False and cfg["heads"] == 3 and cfg["dropout"] == 0.3
    gat2 = kg.GATv2Conv.from_config(cfg)
    assert gat2.add_self_loops_flag is False and gat2.concat is False
    assert kg.MessagePassing.from_config(kg.MessagePassing("max").get_config()).aggregator == "max"


def test_message_passing_base_output_shape():
    lyr = kg.MessagePassing()
    assert lyr.compute_output_shape([(None, 5, 32), (None, 2, None)]) == (None, 5, 32)
    assert lyr.compute_output_shape(((None, 5, 32), (None, 2, None))) == (None, 5, 32)


def test_processed_npz_format_round_trip(tmp_path):
    """The reference's processed-dataset layout (datasets/base.py:124-182): x_i / edge_index_i / y_i / num_graphs /
    num_classes.  Host-side parse only (no device)."""
    from keras_geometric_b200.data_utils import read_processed_npz

    rng = np.random.default_rng(0)
    ref_file = {}
    for i, (n, e) in enumerate([(5, 8), (3, 0)]):
        ref_file[f"x_{i}"] = rng.standard_normal((n, 4)).astype(np.float32)
        ref_file[f"edge_index_{i}"] = rng.integers(0, n, (2, e)).astype(np.int32)
        ref_file[f"y_{i}"] = rng.integers(0, 3, n).astype(np.int64)
    ref_file["edge_attr_0"] = rng.standard_normal((8, 2)).astype(np.float32)
    ref_file["num_graphs"] = 2
    ref_file["num_classes"] = 3
    path = str(tmp_path / "toy.npz")
    np.savez(path, **ref_file)  # written exactly like Dataset._save_processed does
    graphs, num_classes = read_processed_npz(path)
    assert num_classes == 3 and len(graphs) == 2
    np.testing.assert_array_equal(graphs[0]["x"], ref_file["x_0"])
    np.testing.assert_array_equal(graphs[1]["edge_index"], ref_file["edge_index_1"])
    assert graphs[1]["edge_attr"] is None and graphs[0]["edge_attr"].shape == (8, 2)
    np.testing.assert_array_equal(graphs[0]["y"], ref_file["y_0"])


def test_scripts_parse_and_partition_bounds():
    """bench / tools scripts are at least syntactically valid here (they need a GPU to run), and the weak-scaling
    arm's cost-balanced contiguous ranges cover [0, n) monotonically."""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in [os.path.join(root, "bench.py"), os.path.join(root, "bench_dist.py"), os.path.join(root, "bench_extra.py"),
                 os.path.join(root, "__graft_entry__.py")] + glob.glob(os.path.join(root, "tools", "*.py")):
        ast.parse(open(path).read(), filename=path)
    import sys
    sys.path.insert(0, root)
    import bench_dist
    from keras_geometric_b200.dist import cost_balanced_bounds
    gen = torch.Generator().manual_seed(0)
    n = 1000
    dst = (torch.rand(20000, generator=gen) ** 3 * n).long().clamp_(max=n - 1)   # skewed in-degrees
    for world in (1, 2, 3, 8):
        b = cost_balanced_bounds(dst, n, world, bench_dist.NODE_WEIGHT)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert all(b[i] <= b[i + 1] for i in range(world))
        if world > 1:
            deg = torch.bincount(dst, minlength=n) + bench_dist.NODE_WEIGHT
            costs = [float(deg[b[i]:b[i + 1]].sum()) for i in range(world)]
            assert max(costs) <= 1.5 * (sum(costs) / world)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle port on a bounded sample) prints ONE JSON line with the keys the
    measurement contract names; under torchrun every rank but 0 exits silently."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-sample-div", "2048"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GTEPS" and d["higher_is_better"] is True
    assert d["metric"] == "aggregated edges/sec per layer fwd+bwd" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    quiet = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""
