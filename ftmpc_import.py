"""Import helper: registers the package directory ``fault-tolerant-mpc_b200/`` as ``ft_mpc_b200``."""
import importlib.util
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
_PKG = _ROOT / "fault-tolerant-mpc_b200"


def load():
    if "ft_mpc_b200" in sys.modules:
        return sys.modules["ft_mpc_b200"]
    spec = importlib.util.spec_from_file_location("ft_mpc_b200", _PKG / "__init__.py",
                                                  submodule_search_locations=[str(_PKG)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ft_mpc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
