"""CPU-side checks of the drop-in boundary: the C-ABI shared library loads and exports every
symbol include/textgcn_b200.h declares (no compute calls without a GPU), the ctypes mirrors of
the argument structs have the C layout, and the host refuses CPU tensors loudly."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "textgcn_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tgcn_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pytextgcn_b200 import _native
    if not os.path.exists(_native.lib_path()):
        import __graft_entry__ as ge
        ge.build()
    return _native.load()


def test_header_declares_the_documented_entry_points():
    syms = _declared_symbols()
    for must in ["tgcn_csr_from_coo_gcn_norm", "tgcn_spmm", "tgcn_masked_nll", "tgcn_dense_bwd", "tgcn_adam_step",
                 "tgcn_spmm_plan", "tgcn_project", "tgcn_hier_forward", "tgcn_hier_backward", "tgcn_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_signatures_cover_the_header(lib):
    from pytextgcn_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared_symbols()


def test_host_only_calls(lib):
    assert lib.tgcn_version() >= 100
    assert isinstance(lib.tgcn_last_error(), bytes)
    n = ctypes.c_size_t(0)
    assert lib.tgcn_masked_nll_workspace_bytes(1000, ctypes.byref(n)) == 0 and n.value >= 8000 + 4096
    assert lib.tgcn_spmm_plan_workspace_bytes(10, ctypes.byref(n)) == 0
    # argument validation happens before any CUDA call
    assert lib.tgcn_spmm(None, None) != 0
    assert b"args null" in lib.tgcn_last_error()
    assert lib.tgcn_adam_step(None, None, None, None, None, 4, 0.1, 0.9, 0.999, 1e-8, 0, 1, None, None, None) != 0


def test_struct_layout_matches_c(lib, tmp_path):
    """sizeof/offsetof of the argument structs as seen by the C compiler == ctypes mirror."""
    from pytextgcn_b200 import _native
    src = tmp_path / "layout.c"
    src.write_text(f'''
#include <stdio.h>
#include <stddef.h>
#include "{HEADER}"
int main(void) {{
  printf("%zu %zu %zu %zu %zu\\n", sizeof(tgcn_spmm_args), offsetof(tgcn_spmm_args, F), offsetof(tgcn_spmm_args, philox_offset_dev),
         offsetof(tgcn_spmm_args, ldp), offsetof(tgcn_spmm_args, bias_len));
  printf("%zu %zu %zu %zu\\n", sizeof(tgcn_dense_bwd_args), offsetof(tgcn_dense_bwd_args, H), offsetof(tgcn_dense_bwd_args, dZ1),
         offsetof(tgcn_dense_bwd_args, db_out));
  printf("%zu %zu\\n", offsetof(tgcn_spmm_args, split_counters), offsetof(tgcn_spmm_args, adam_param_mirror_mc));
  return 0;
}}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    S, D = _native.SpmmArgs, _native.DenseBwdArgs
    mine = [ctypes.sizeof(S), S.F.offset, S.philox_offset_dev.offset, S.ldp.offset, S.bias_len.offset,
            ctypes.sizeof(D), D.H.offset, D.dZ1.offset, D.db_out.offset,
            S.split_counters.offset, S.adam_param_mirror_mc.offset]
    assert [int(v) for v in out] == mine


def test_no_cpu_fallback():
    from pytextgcn_b200 import GCN, ops
    from pytextgcn_b200.graph import upload_graph
    from helpers import karate_graph
    g = karate_graph()
    with pytest.raises(RuntimeError, match="CUDA"):
        GCN(34, 4)(g)
    with pytest.raises(RuntimeError, match="CUDA"):
        upload_graph(g.edge_index, g.edge_attr, 34)
    with pytest.raises(RuntimeError):
        ops.adam_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), None, lr=0.1, step=1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pytextgcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"


def test_module_checkpoint_compatibility(tmp_path):
    """Whole-module th.save / th.load (flat_amazon.py:126-128, eval_perlabel.py:13-19) and the reference's
    state_dict layout: layers.{i}.weight (in,out), layers.{i}.bias (out)."""
    from pytextgcn_b200 import GCN
    m = GCN(40, 5, n_hidden_gcn=16, dropout=0.7)
    sd = m.state_dict()
    assert list(sd) == ["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"]
    assert tuple(sd["layers.0.weight"].shape) == (40, 16) and tuple(sd["layers.1.weight"].shape) == (16, 5)
    p = tmp_path / "gcn.nn"
    torch.save(m, str(p))
    m2 = torch.load(str(p), weights_only=False)
    assert isinstance(m2, GCN) and m2.dropout == 0.7
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    m3 = GCN(40, 5, n_hidden_gcn=16)
    m3.load_state_dict(sd)
