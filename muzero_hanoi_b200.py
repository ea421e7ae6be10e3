"""Import shim: the package directory is named ``muzero-hanoi_b200`` (not a valid Python
identifier), so this module makes it importable as ``muzero_hanoi_b200`` by pointing its
``__path__`` at that directory and executing the package ``__init__`` in this namespace."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "muzero-hanoi_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
