"""Load the UNMODIFIED reference implementation of the hot path, when it is present.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  ``/root/reference`` exists only in
the build container; on the GPU box the loader falls back to the copy of the same files that
``oracle/ship_reference.py`` places under ``baseline/_ref`` (git-ignored, shipped with the
snapshot).  Everything here is optional:
``available()`` says whether it can be used; tests that need it skip otherwise, and
``tests/golden/make_golden.py`` uses it to produce the committed golden vectors.

``import codebase...`` fails without fairseq (codebase/__init__.py:6 imports every
sub-package, which register fairseq models).  The hot-path files themselves only need
torch, so they are loaded by file path under stub parent packages:

  codebase/utils/functions.py, monotonic_attention.py, p_choose_strategy.py
  codebase/models/torch_cif/cif.py
  codebase/modules/monotonic_multihead_attention.py   (with a minimal stand-in for
      fairseq.modules.MultiheadAttention and the registry decorator)

No reference source is copied; the files are executed where they lie.
"""
import importlib.util
import os
import sys
import types
from typing import Optional

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SHIPPED_ROOT = os.path.join(_REPO_ROOT, "baseline", "_ref")


def _default_root() -> str:
    """/root/reference in the build container; on the GPU box the copy that
    ``oracle/ship_reference.py`` placed under the git-ignored ``baseline/_ref``."""
    env = os.environ.get("SIMULST_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/codebase/utils/monotonic_attention.py"):
        return "/root/reference"
    return _SHIPPED_ROOT


REF_ROOT = _default_root()

_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "codebase/utils/monotonic_attention.py"))


def _stub_package(name: str):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        mod.__path__ = []  # mark as package
        sys.modules[name] = mod
    return sys.modules[name]


def _load(mod_name: str, rel_path: str):
    if mod_name in _cache:
        return _cache[mod_name]
    path = os.path.join(REF_ROOT, rel_path)
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod
    spec.loader.exec_module(mod)
    _cache[mod_name] = mod
    return mod


def load_utils():
    """Returns (functions, monotonic_attention, p_choose_strategy) reference modules."""
    _stub_package("codebase")
    _stub_package("codebase.utils")
    fn = _load("codebase.utils.functions", "codebase/utils/functions.py")
    ma = _load("codebase.utils.monotonic_attention", "codebase/utils/monotonic_attention.py")
    pc = _load("codebase.utils.p_choose_strategy", "codebase/utils/p_choose_strategy.py")
    return fn, ma, pc


def load_cif():
    """Returns the reference torch_cif.cif module."""
    return _load("_simulst_ref_cif", "codebase/models/torch_cif/cif.py")


def _install_fairseq_stub():
    """Smallest stand-in for what monotonic_multihead_attention.py needs from fairseq:
    a MultiheadAttention base with q/k/v/out projections and incremental-state accessors."""
    if "fairseq.modules" in sys.modules and hasattr(sys.modules["fairseq.modules"], "_simulst_stub"):
        return
    import torch.nn as nn

    class MultiheadAttention(nn.Module):
        def __init__(self, embed_dim, num_heads, kdim=None, vdim=None, dropout=0.0,
                     bias=True, encoder_decoder_attention=False, **kw):
            super().__init__()
            self.embed_dim = embed_dim
            self.kdim = kdim if kdim is not None else embed_dim
            self.vdim = vdim if vdim is not None else embed_dim
            self.qkv_same_dim = self.kdim == embed_dim and self.vdim == embed_dim
            self.num_heads = num_heads
            self.head_dim = embed_dim // num_heads
            self.scaling = self.head_dim ** -0.5
            self.k_proj = nn.Linear(self.kdim, embed_dim, bias=bias)
            self.v_proj = nn.Linear(self.vdim, embed_dim, bias=bias)
            self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
            self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

        def get_incremental_state(self, incremental_state, key):
            if incremental_state is None:
                return None
            return incremental_state.get((id(self), key))

        def set_incremental_state(self, incremental_state, key, value):
            if incremental_state is not None:
                incremental_state[(id(self), key)] = value
            return incremental_state

    fairseq = _stub_package("fairseq")
    fmods = _stub_package("fairseq.modules")
    fmods.MultiheadAttention = MultiheadAttention
    fmods._simulst_stub = True
    fairseq.modules = fmods


def load_mma_module():
    """Returns the reference modules.monotonic_multihead_attention module (classes
    MonotonicAttention, MonotonicInfiniteLookbackAttention, ...)."""
    load_utils()
    _install_fairseq_stub()
    pkg = _stub_package("codebase.modules")
    if not hasattr(pkg, "register_monotonic_attention"):
        def register_monotonic_attention(name):
            def deco(cls):
                return cls
            return deco
        pkg.register_monotonic_attention = register_monotonic_attention
    return _load("codebase.modules.monotonic_multihead_attention",
                 "codebase/modules/monotonic_multihead_attention.py")


def make_attention(kind: str = "infinite_lookback", embed_dim: int = 64, heads: int = 4,
                   mass_preservation: bool = True, eps: float = 1e-6,
                   energy_bias: bool = True, seed: Optional[int] = 0):
    """Instantiate a reference attention module on CPU with deterministic weights."""
    import argparse
    import torch
    mod = load_mma_module()
    cls = {"hard_aligned": mod.MonotonicAttention,
           "infinite_lookback": mod.MonotonicInfiniteLookbackAttention}[kind]
    args = argparse.Namespace(
        decoder_embed_dim=embed_dim, decoder_attention_heads=heads,
        encoder_embed_dim=embed_dim, attention_dropout=0.0,
        attention_eps=eps, mass_preservation=mass_preservation,
        noise_mean=0.0, noise_var=1.0, energy_bias_init=-2.0, energy_bias=energy_bias)
    if seed is not None:
        torch.manual_seed(seed)
    return cls(args)


def load_fixed_pre_decision():
    """Returns the reference modules.fixed_pre_decision module (classes
    MonotonicAttentionFixedStride, MonotonicInfiniteLookbackAttentionFixedStride,
    WaitKAttentionFixedStride -- modules/fixed_pre_decision.py:175-190)."""
    load_mma_module()
    return _load("codebase.modules.fixed_pre_decision", "codebase/modules/fixed_pre_decision.py")


def make_fixed_pre_decision_attention(kind: str = "infinite_lookback", ratio: int = 8,
                                      pool: str = "average", embed_dim: int = 64, heads: int = 4,
                                      mass_preservation: bool = True, eps: float = 1e-6,
                                      seed: Optional[int] = 0):
    """Instantiate a reference `*_fixed_pre_decision` attention module (exp/2-mma.sh:56-57)."""
    import argparse
    import torch
    mod = load_fixed_pre_decision()
    cls = {"hard_aligned": mod.MonotonicAttentionFixedStride,
           "infinite_lookback": mod.MonotonicInfiniteLookbackAttentionFixedStride}[kind]
    args = argparse.Namespace(
        decoder_embed_dim=embed_dim, decoder_attention_heads=heads,
        encoder_embed_dim=embed_dim, attention_dropout=0.0,
        attention_eps=eps, mass_preservation=mass_preservation,
        noise_mean=0.0, noise_var=1.0, energy_bias_init=-2.0, energy_bias=True,
        fixed_pre_decision_type=pool, fixed_pre_decision_ratio=ratio,
        fixed_pre_decision_pad_threshold=0.3)
    if seed is not None:
        torch.manual_seed(seed)
    return cls(args)


def load_ssnt():
    """Returns the reference criterion/ssnt_loss/ssnt_loss.py module (torch only)."""
    return _load("_simulst_ref_ssnt_loss", "codebase/criterion/ssnt_loss/ssnt_loss.py")


def _function_source(rel_path: str, class_name: Optional[str], func_name: str) -> str:
    """Source text of one (method or) function of a reference file, located with ``ast`` so
    the file itself never has to be imported (its module-level imports need fairseq)."""
    import ast
    import textwrap
    path = os.path.join(REF_ROOT, rel_path)
    text = open(path).read()
    tree = ast.parse(text)
    body = tree.body
    if class_name is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == func_name)
    lines = text.splitlines()[fn.lineno - 1:fn.end_lineno]
    return textwrap.dedent("\n".join(lines))


def load_mma_latency_loss(latency_metrics: dict):
    """``MMACriterion.compute_latency_loss`` (criterion/mma_criterion.py:138-207) as a plain
    function ``f(self, model, sample, net_output)``, executed from the unmodified reference file.
    The file's own imports (fairseq, SimulEval) are absent here, so only this method's source is
    compiled; its one free name, ``LATENCY_METRICS`` (SimulEval's latency functions,
    mma_criterion.py:15-26), is supplied by the caller -- see ``oracle.latency``."""
    import torch
    src = _function_source("codebase/criterion/mma_criterion.py", "MMACriterion", "compute_latency_loss")
    ns = {"torch": torch, "LATENCY_METRICS": latency_metrics}
    exec(compile(src, os.path.join(REF_ROOT, "codebase/criterion/mma_criterion.py"), "exec"), ns)
    return ns["compute_latency_loss"]


def load_best_alignment_python():
    """The reference's ``best_alignment`` Python wrapper (criterion/best_alignment/__init__.py:25-111:
    final-state selection + the S-iteration back-trace) as a function that takes the native
    extension as an argument instead of JIT-building it at import time (:10-17), so that it can
    run here on CPU over the oracle's forward pass, and on the GPU box over the reference's own
    kernel when ``load_best_alignment_extension`` succeeds."""
    import torch
    src = _function_source("codebase/criterion/best_alignment/__init__.py", None, "best_alignment")
    ns = {"torch": torch}
    exec(compile(src, os.path.join(REF_ROOT, "codebase/criterion/best_alignment/__init__.py"), "exec"), ns)
    fn = ns["best_alignment"]

    def call(extension, *args, **kw):
        ns["extension"] = extension
        return fn(*args, **kw)
    return call


def load_best_alignment_extension(build_dir: Optional[str] = None):
    """JIT-build the reference's only native kernel (best_alignment.cu/.cpp) with
    torch.utils.cpp_extension, exactly as criterion/best_alignment/__init__.py:10-17 does --
    needs a CUDA device and a few minutes of nvcc; used by the `-m gpu` parity test and by the
    bench's reference leg for the CTC alignment row.  Outputs go to oracle/_ref (git-ignored)."""
    import torch.utils.cpp_extension as ext
    src_dir = os.path.join(REF_ROOT, "codebase/criterion/best_alignment")
    build_dir = build_dir or os.path.join(_REPO_ROOT, "oracle", "_ref", "best_alignment_build")
    os.makedirs(build_dir, exist_ok=True)
    return ext.load("best_alignment_fn_ref",
                    sources=[os.path.join(src_dir, "best_alignment.cpp"),
                             os.path.join(src_dir, "best_alignment.cu")],
                    build_directory=build_dir, verbose=False)


def _class_source(rel_path: str, class_name: str) -> str:
    import ast
    path = os.path.join(REF_ROOT, rel_path)
    text = open(path).read()
    tree = ast.parse(text)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    first = min([node.lineno] + [d.lineno for d in node.decorator_list])
    return "\n".join(text.splitlines()[first - 1:node.end_lineno])


def load_cif_layer():
    """The reference's ``CIFLayer`` class (models/cif_transformer.py:110-261), compiled from its
    own source.  The file's module-level imports pull in fairseq and the Emformer encoder, so only
    the class statement is executed, with stand-ins for the four fairseq names it uses:
    ``LayerNorm`` / ``Linear`` (torch.nn), ``with_incremental_state`` (adds the get/set accessors
    fairseq's decorator adds) and ``CausalConvTBC`` (a left-padded Conv over time-major input with
    the incremental input cache of modules/causal_conv.py:80-98).  ``cif_function`` is the
    reference's.  Returns (CIFLayer, CausalConvTBC)."""
    import torch
    import torch.nn as nn
    from typing import Dict, List, Optional

    def with_incremental_state(cls):
        def get_incremental_state(self, incremental_state, key):
            if incremental_state is None:
                return None
            return incremental_state.get((id(self), key))

        def set_incremental_state(self, incremental_state, key, value):
            if incremental_state is not None:
                incremental_state[(id(self), key)] = value
            return incremental_state
        cls.get_incremental_state = get_incremental_state
        cls.set_incremental_state = set_incremental_state
        return cls

    @with_incremental_state
    class CausalConvTBC(nn.Module):
        """(T, B, C_in) -> (T, B, C_out), output t sees inputs <= t; streaming calls keep the
        last kernel_size-1 input frames."""

        def __init__(self, in_channels, out_channels, kernel_size):
            super().__init__()
            self.kernel_size = kernel_size
            self.conv = nn.Conv1d(in_channels, out_channels, kernel_size)

        def forward(self, x, incremental_state=None):
            t, b, c = x.shape
            pad = self.kernel_size - 1
            if incremental_state is None:
                left = x.new_zeros(pad, b, c)
            else:
                left = self.get_incremental_state(incremental_state, "conv_state")
                if left is None:
                    left = x.new_zeros(pad, b, c)
            full = torch.cat([left, x], dim=0)
            if incremental_state is not None:
                self.set_incremental_state(incremental_state, "conv_state", full[full.shape[0] - pad:])
            return self.conv(full.permute(1, 2, 0)).permute(2, 0, 1)

    ns = {"torch": torch, "nn": nn, "Tensor": torch.Tensor, "Optional": Optional, "Dict": Dict, "List": List,
          "LayerNorm": nn.LayerNorm, "Linear": nn.Linear, "CausalConvTBC": CausalConvTBC,
          "with_incremental_state": with_incremental_state, "cif_function": load_cif().cif_function}
    src = _class_source("codebase/models/cif_transformer.py", "CIFLayer")
    exec(compile(src, os.path.join(REF_ROOT, "codebase/models/cif_transformer.py"), "exec"), ns)
    return ns["CIFLayer"], CausalConvTBC
