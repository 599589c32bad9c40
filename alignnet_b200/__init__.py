"""Import shim: the product package lives in the directory ``alignnet-3d_b200/`` (a name Python
cannot import directly); this module exposes it as ``alignnet_b200``."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "alignnet-3d_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
