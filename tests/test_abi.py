"""The C-ABI library loads on a machine without a GPU, exports every symbol include/contrast_b200.h declares, agrees with
the ctypes mirror on struct layouts, reports the reference's error variants, and refuses to compute without a device."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from contrast_renderer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "contrast_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_mirror_list_the_same_symbols():
    assert declared_functions() == sorted(_abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(cr):
    lib = C.CDLL(cr.library_path())
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/contrast_b200.h but not exported"
    assert cr.lib().cr_abi_version() == 3


def test_struct_layouts_match_the_header():
    names = ["cr_stroke_options", "cr_dash_interval", "cr_dynamic_stroke_options", "cr_path_soa", "cr_config", "cr_shape_layout", "cr_draw_command", "cr_stats"]
    mirrors = [_abi.StrokeOptionsC, _abi.DashIntervalC, _abi.DynamicStrokeOptionsC, _abi.PathSoAC, _abi.ConfigC, _abi.ShapeLayoutC, _abi.DrawCommandC, _abi.StatsC]
    src = '#include <stdio.h>\n#include "contrast_b200.h"\nint main(void){' + "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c_path, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(c_path, "w").write(src)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c_path, "-o", exe])   # the header is plain C
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(m) for m in mirrors]
    assert C.sizeof(_abi.StrokeOptionsC) == 24   # ~24 B per path


def test_status_strings_follow_error_rs(cr):
    # enum Error, src/error.rs:5-16, in declaration order
    names = ["Ok", "NumberOfStencilBitsIsUnsupported", "ClipStackOverflow", "TooManyNestedOpacityGroups", "TooManyDashIntervals",
             "DynamicStrokeOptionsIndexOutOfBounds"]
    for code, name in enumerate(names):
        assert cr.lib().cr_status_string(code).decode() == name


def test_argument_validation_needs_no_device(cr):
    # Renderer::new validates the stencil bit split before touching the device (src/renderer.rs:433)
    out = C.c_void_p()
    cfg = cr.Configuration(winding_counter_bits=0).to_c()
    assert cr.lib().cr_renderer_create(C.byref(cfg), C.byref(out)) == _abi.CR_ERR_NUMBER_OF_STENCIL_BITS_IS_UNSUPPORTED
    cfg = cr.Configuration(winding_counter_bits=6, clip_nesting_counter_bits=3).to_c()
    assert cr.lib().cr_renderer_create(C.byref(cfg), C.byref(out)) == _abi.CR_ERR_NUMBER_OF_STENCIL_BITS_IS_UNSUPPORTED
    cfg = cr.Configuration(msaa_sample_count=3).to_c()
    assert cr.lib().cr_renderer_create(C.byref(cfg), C.byref(out)) == _abi.CR_ERR_INVALID_ARGUMENT
    assert cr.lib().cr_renderer_create(None, C.byref(out)) == _abi.CR_ERR_INVALID_ARGUMENT


def test_no_cpu_fallback(cr):
    """Without a CUDA device the product path fails loudly instead of computing anywhere else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(cr.Error) as e:
        cr.Renderer()
    assert e.value.status == _abi.CR_ERR_NO_DEVICE


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/ (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "contrast_renderer_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f"{f} uses the oracle"
                assert not re.search(r'#include\s+"[^"]*oracle/', text), f"{f} includes oracle sources"
