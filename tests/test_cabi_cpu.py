"""CPU-only checks of the drop-in boundary: libbbmpc.so loads, exports every symbol that
include/bbmpc.h declares, its ctypes mirror agrees with the header, the host-side Philox matches
the Random123 known-answer vectors, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from blackbox_mpc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bbmpc.h")


def _header_text():
    with open(HEADER) as f:
        txt = f.read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def header_functions():
    return sorted(set(re.findall(r"\b(bbmpc_\w+)\s*\(", _header_text())))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"include/bbmpc.h declares {n} but libbbmpc.so does not export it"


def test_ctypes_signatures_cover_header():
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_header_constants_match_python_mirror():
    txt = _header_text()
    consts = dict(re.findall(r"#define\s+BBMPC_(\w+)\s+\(?(-?\d+)\)?", txt))
    for k, v in consts.items():
        if k == "H_":
            continue
        assert getattr(_lib, k) == int(v), k


def test_opt_config_layout_matches_header():
    """Field order of struct bbmpc_opt_config == _lib.OptConfig._fields_."""
    txt = _header_text()
    body = re.search(r"typedef struct bbmpc_opt_config \{(.*?)\} bbmpc_opt_config;", txt, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("*", " ").split(None, 1)[1] if not decl.startswith("const") else decl.replace("*", " ").split(None, 2)[2]
        fields += [n.strip() for n in names.split(",")]
    assert fields == [f[0] for f in _lib.OptConfig._fields_]


@pytest.mark.parametrize("ctr,key,expect", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox4x32_10_known_answers(ctr, key, expect):
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    lib = _lib.load()
    c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    lib.bbmpc_philox4x32_host(c, k, o)
    assert list(o) == expect


def test_no_cpu_fallback():
    """Without a CUDA device ctx_create fails with BBMPC_ECUDA and says so (run on the CPU box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.bbmpc_ctx_create(0, C.c_uint64(0), C.byref(h))
    assert rc == _lib.ECUDA and not h.value
    assert b"no CPU fallback" in lib.bbmpc_last_error(None)
    from blackbox_mpc_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()
