"""Import alias for the product package.

The product lives in ``video-based-gait-analysis-for-dementia_b200/`` (a directory
name Python cannot import because of the hyphens).  This shim makes it importable
as ``gaitb200``: sub-modules resolve inside the real directory, and the real
``__init__`` body runs in this module's namespace.
"""
from pathlib import Path as _Path

_REAL = _Path(__file__).resolve().parent.parent / "video-based-gait-analysis-for-dementia_b200"
__path__ = [str(_REAL)]
exec(compile((_REAL / "__init__.py").read_text(), str(_REAL / "__init__.py"), "exec"))
