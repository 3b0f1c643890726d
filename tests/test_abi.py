"""CPU-side checks of the C-ABI boundary: libvex.so builds for sm_100a, loads without a GPU, exports
every symbol include/vex.h declares, and the ctypes mirror of vexGemmArgs has the C layout."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vex.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vex_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_listed():
    from mmmm_b200 import _lib
    lib = _lib.lib()
    decl = _declared_symbols()
    assert decl == sorted(_lib.SYMBOLS), (decl, sorted(_lib.SYMBOLS))
    for name in decl:
        assert hasattr(lib, name), f"{name} not exported by libvex.so"
    assert lib.vex_abi_version() == 7
    assert lib.vex_error_string(-2).decode() == "unsupported shape"


def test_no_gpu_calls_fail_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    from mmmm_b200 import _lib
    assert _lib.lib().vex_device_check() == -4  # VEX_E_NO_DEVICE
    from mmmm_b200 import ops
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.silu_mul(torch.zeros(2, 8, dtype=torch.bfloat16), torch.zeros(2, 8, dtype=torch.bfloat16),
                     torch.zeros(1, dtype=torch.int32), torch.zeros(2, 8, dtype=torch.bfloat16))


def test_gemm_args_struct_layout_matches_c():
    from mmmm_b200._lib import GemmArgs
    fields = [f[0] for f in GemmArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
            'printf("%zu\\n", sizeof(vexGemmArgs));']
    prog += [f'printf("%zu\\n", offsetof(vexGemmArgs, {f}));' for f in fields]
    prog += ["return 0;}"]
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write("\n".join(prog))
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", c, "-o", exe], check=True)
        vals = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(GemmArgs)
    for f, off in zip(fields, vals[1:]):
        assert getattr(GemmArgs, f).offset == off, f


def test_sass_uses_tcgen05_and_tma():
    """The built library really contains Blackwell tensor-core / TMA instructions (SASS mnemonics from
    B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG)."""
    from mmmm_b200 import build
    lib = build.build()
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "LDTM" in sass and "UTMALDG" in sass
