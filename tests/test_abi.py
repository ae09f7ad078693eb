"""The C-ABI library loads without a GPU and exports every symbol include/sayal.h declares."""
import ctypes
import re
from pathlib import Path

from opensayal_b200 import LIB_PATH, load
from opensayal_b200._abi import SYMBOLS, SayalArrow, SayalConfig, SayalSlab, SayalSource, SayalVisual

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "sayal.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sayal_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(str(LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sayal.h but not exported"
    bound = {n for n, _, _ in SYMBOLS}
    assert set(names) == bound, f"ctypes table and header disagree: {set(names) ^ bound}"


def test_load_and_version():
    lib = load()
    assert lib.sayal_abi_version() == 3
    assert lib.sayal_last_error() is not None


def test_struct_sizes_match_header():
    # 31 4-byte members, 5 and 4 respectively (include/sayal.h); a mismatch would corrupt every call
    assert ctypes.sizeof(SayalConfig) == 31 * 4
    assert ctypes.sizeof(SayalSource) == 5 * 4
    assert ctypes.sizeof(SayalSlab) == 4 * 4
    assert ctypes.sizeof(SayalVisual) == 17 * 4
    assert ctypes.sizeof(SayalArrow) == 9 * 4


def test_no_oracle_in_product():
    """The product must not import, link or call the oracle (a CPU fallback would void parity claims)."""
    for path in list((ROOT / "opensayal_b200").rglob("*.py")) + list((ROOT / "opensayal_b200" / "csrc").glob("*.*")):
        if path.suffix in (".o", ".so", ".log"):
            continue
        text = path.read_text(errors="ignore")
        assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, path


def test_host_only_entry_points_do_not_need_a_gpu():
    """Argument checks and host-side helpers answer without a device (no compute call here)."""
    import ctypes as C
    lib = load()
    assert lib.sayal_stream_hold(None) < 0 and lib.sayal_stream_release(None) < 0
    assert lib.sayal_get_fields(None, 0, None, None) < 0 and lib.sayal_set_fields(None, 0, None, None) < 0
    assert lib.sayal_slab_ipc_export(None, None) < 0
    assert lib.sayal_plan_log(None, None, 0) < 0
    assert b"null" in lib.sayal_last_error()
    from opensayal_b200 import _abi
    assert _abi.LINK_INFO_BYTES == 256 and _abi.SAYAL_ELINK == -7
