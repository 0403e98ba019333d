"""Import shim: the package directory is named ``stm32f7-rtlsdr_b200`` (with a hyphen, as the
build layout asks), which the ``import`` statement cannot spell.  ``import b200sdr`` gives the
same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("stm32f7-rtlsdr_b200")
sys.modules[__name__] = _pkg
