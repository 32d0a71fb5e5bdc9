"""CPU-side checks of the drop-in boundary: the C-ABI library loads (no GPU needed) and exports
every function include/qe_engine.h declares; the host-side stream hash equals the oracle's."""
import os
import re

from oracle import rng as orng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "qe_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qe_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from dist_classicrl_b200 import capi

    lib = capi.lib()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"libqe_b200.so does not export {name}"
        assert name in capi.SIGNATURES, f"capi.py has no signature for {name}"
    assert b"sm_100a" in lib.qe_build_info()


def test_host_stream_hash_matches_oracle():
    from dist_classicrl_b200 import capi

    lib = capi.lib()
    for args in [(0, 0, 0, 0), (5, 3, 77, 4), (123, 0xFFFFFFFF, (1 << 24) - 1, 7)]:
        assert lib.qe_stream_u32(*args) == orng.stream_u32(*args)


def test_sass_is_sm_100a():
    import subprocess

    from dist_classicrl_b200 import capi

    out = subprocess.run(["cuobjdump", "-lelf", capi.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
