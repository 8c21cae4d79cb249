"""Loader for the hyphen-named package directory `rtl-sdr-rs_b200/` (registers it as `rtl_sdr_rs_b200`)."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "rtl-sdr-rs_b200"
NAME = "rtl_sdr_rs_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(NAME, None)
        raise
    return mod
