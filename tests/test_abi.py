"""CPU: the C-ABI library loads and exports every symbol include/clipgp.h declares; compute entry points
fail loudly (status + message, never a silent fallback) when misused or when no GPU is present."""
import ctypes
import os
import re

import pytest
import torch

from clip_gp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "clipgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clipgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/clipgp.h but not exported by libclipgp.so"


def test_python_binding_covers_the_header():
    assert set(_lib.exported_symbols()) == set(declared_symbols())
    lib = _lib.load()
    assert lib.clipgp_version() >= 100


def test_bad_arguments_report_an_error_message():
    lib = _lib.load()
    rc = lib.clipgp_ece_hist(None, None, 5, None, 10, None, None, None, None)
    assert rc != 0 and b"NULL" in lib.clipgp_last_error()
    rc = lib.clipgp_ece_hist(None, None, 5, None, 1000, None, None, None, None)
    assert rc != 0 and b"n_bins" in lib.clipgp_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "clipgp_ece_hist")


def test_struct_layout_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "clipgp.h")).read()
    body = hdr[hdr.index("typedef struct clipgp_gp_args {"):hdr.index("} clipgp_gp_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("*", " ").split()
        # "int64_t C, T, n, d, S" style
        tail = decl.split(None, 1)[1] if " " in decl else decl
        tail = re.sub(r"^(const\s+)?(unsigned\s+)?\w+\s*\**", "", decl).strip() if "," in decl else names[-1]
        fields += [x.strip().lstrip("*") for x in (tail.split(",") if "," in decl else [names[-1]])]
    assert fields == [f[0] for f in _lib.GpArgs._fields_]


def test_peer_args_layout_matches_a_c_compiler(tmp_path):
    """clipgp_peer_args (the fused NVLink optimiser step): ctypes mirror == what gcc lays out from the header."""
    import ctypes, subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "clipgp.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(clipgp_peer_args), offsetof(clipgp_peer_args, g), offsetof(clipgp_peer_args, flags), offsetof(clipgp_peer_args, m), '
                   'offsetof(clipgp_peer_args, lr_dev), offsetof(clipgp_peer_args, step), offsetof(clipgp_peer_args, kl_scale)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    P = _lib.PeerArgs
    assert got == [ctypes.sizeof(P), P.g.offset, P.flags.offset, P.m.offset, P.lr_dev.offset, P.step.offset, P.kl_scale.offset]
    lib = _lib.load()
    assert lib.clipgp_peer_adamw(None, None) != 0 and b"NULL" in lib.clipgp_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from clip_gp_b200 import metrics
    with pytest.raises(RuntimeError):
        metrics.compute_ece(torch.randn(8, 4), torch.zeros(8, dtype=torch.long))


def test_gp_bwd_args_layout_matches_a_c_compiler(tmp_path):
    """clipgp_gp_bwd_args: ctypes mirror == what gcc lays out from the header."""
    import subprocess
    names = ["dkl_scalar", "dZ_last", "dmean_x", "proto_dP_stride_s", "proto_dP_scale", "proto_norm", "proto_D", "dw_out", "tl_Z", "tl_dlT_ld",
             "tl_B", "tl_mode", "tl_scale"]
    src = tmp_path / "szb.c"
    fmt = " ".join(["%zu"] * (len(names) + 1))
    args = ", ".join(["sizeof(clipgp_gp_bwd_args)"] + [f"offsetof(clipgp_gp_bwd_args, {n})" for n in names])
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "clipgp.h"\nint main(void) {{ printf("{fmt}\\n", {args}); return 0; }}\n')
    exe = tmp_path / "szb"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    B = _lib.GpBwdArgs
    assert got == [ctypes_sizeof(B)] + [getattr(B, n).offset for n in names]


def ctypes_sizeof(t):
    import ctypes
    return ctypes.sizeof(t)
