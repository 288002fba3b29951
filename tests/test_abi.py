"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol the header declares,
and the host-side guards fail loudly (no silent CPU fallback).  No compute calls (no GPU here)."""
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from noisynet_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "noisynet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_version(lib):
    assert lib.nn_abi_version() == _lib.ABI_VERSION
    m = re.search(r"#define NN_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "noisynet_b200.h")).read())
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """Every struct of include/noisynet_b200.h, compiled by gcc, against its ctypes mirror: total size, the offset of
    every field and the field ORDER (names are allowed to differ only by the trailing underscore of `in_`)."""
    import ctypes as C
    import subprocess
    assert C.sizeof(_lib.ConvGeom) == 36
    assert C.sizeof(_lib.Rng) == 24
    pairs = {"nn_rng": _lib.Rng, "nn_conv_geom": _lib.ConvGeom, "nn_conv_fwd_args": _lib.ConvFwdArgs,
             "nn_conv_dgrad_args": _lib.ConvDgradArgs, "nn_conv_wgrad_args": _lib.ConvWgradArgs,
             "nn_adamw_tensor": _lib.AdamWTensor, "nn_wprep_job": _lib.WPrepJob, "nn_stage_args": _lib.StageArgs,
             "nn_stage_bwd_args": _lib.StageBwdArgs, "nn_tail_args": _lib.TailArgs}
    hdr = open(os.path.join(ROOT, "include", "noisynet_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"typedef struct (nn_[a-z0-9_]+)\s*\{", body))
    assert declared == set(pairs), declared ^ set(pairs)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "noisynet_b200.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        m = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (cname, cname), body, flags=re.S)
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const float *gamma, *beta" / "int32_t B, C, H, W, pool" / "nn_rng rng" / "void* xp"
            names = [re.sub(r"[^A-Za-z0-9_]", "", part.split()[-1]) for part in decl.split(",")]
            fields += names
        for f in fields:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
        ct_names = [n.rstrip("_") for n, _ in ct._fields_]
        assert ct_names == fields, (cname, ct_names, fields)
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True)
    for line in out.splitlines():
        cname, field, val = line.split()
        ct = pairs[cname]
        if field == "size":
            assert C.sizeof(ct) == int(val), (cname, C.sizeof(ct), val)
        else:
            cf = field if hasattr(ct, field) else field + "_"
            assert getattr(ct, cf).offset == int(val), (cname, field, getattr(ct, cf).offset, val)


def test_cpu_tensors_are_rejected_loudly(lib):
    x = torch.zeros(4, 4)
    with pytest.raises(_lib.NoisyNetLibraryError):
        ops.quantize_fwd(x, 4, 0.0, 1.0)
    from noisynet_b200.hardware_model import NoisyConv2d, QuantMeasure
    m = NoisyConv2d(3, 4, 3)
    with pytest.raises(_lib.NoisyNetLibraryError):
        m(torch.zeros(1, 3, 8, 8))
    with pytest.raises(_lib.NoisyNetLibraryError):
        QuantMeasure(4, max_value=1.0)(torch.zeros(3))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnoisynet_b200.so")
    with pytest.raises(_lib.NoisyNetLibraryError):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "noisynet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", ""), os.path.join(dirpath, f)


def test_boundary_surface_matches_reference_signatures():
    import inspect
    from noisynet_b200 import hardware_model as hm, quant as q
    sig = inspect.signature(hm.NoisyConv2d.__init__)
    assert list(sig.parameters)[1:] == ["in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation",
                                        "groups", "bias", "num_bits", "num_bits_weight", "noise", "test_noise",
                                        "stochastic", "debug"]
    assert sig.parameters["noise"].default == 0.5 and sig.parameters["bias"].default is False
    sig = inspect.signature(hm.NoisyLinear.__init__)
    assert list(sig.parameters)[1:] == ["in_features", "out_features", "bias", "num_bits", "num_bits_weight", "noise",
                                        "test_noise", "stochastic", "debug"]
    sig = inspect.signature(hm.QuantMeasure.__init__)
    assert sig.parameters["pctl"].default == 90. and sig.parameters["num_bits"].default == 8
    assert inspect.signature(q.QuantMeasure.__init__).parameters["pctl"].default == .999
    sig = inspect.signature(hm.add_noise_calculate_power)
    assert list(sig.parameters) == ["self", "args", "arrays", "input", "weights", "output", "layer_type", "i",
                                    "layer_num", "merged_dac"]
    with pytest.raises(SystemExit):
        hm.QuantMeasure(4, pctl=0.5)                      # hardware_model.py:222-225
    m = hm.NoisyConv2d(3, 8, 5, num_bits=4, num_bits_weight=4)
    keys = set(m.state_dict())
    assert {"weight", "quantize_input.running_min", "quantize_input.running_max",
            "quantize_weights.running_min", "quantize_weights.running_max"} <= keys
    assert m.state_dict()["quantize_weights.running_min"].shape == (1,)
    assert m.state_dict()["quantize_weights.running_max"].shape == ()
    assert isinstance(m, torch.nn.Conv2d) and isinstance(hm.NoisyLinear(4, 2), torch.nn.Linear)


def test_net_state_dict_keys_match_reference_fixture(golden):
    from noisynet_b200.net import NoisyNet, default_args, with_quant
    g = golden("net_step")
    ref_keys = {k[len("q4_sd0_"):] for k in g if k.startswith("q4_sd0_")}
    m = NoisyNet(with_quant(default_args(fm1=9, fm2=12, fc=24)))
    assert set(m.state_dict()) == ref_keys
