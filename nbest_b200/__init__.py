"""Importable alias of the `n-best-asr-transformer_b200/` package (its directory name is not a Python identifier).

`import nbest_b200` resolves every submodule (`nbest_b200.ops`, `nbest_b200.model`, ...) inside
`n-best-asr-transformer_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "n-best-asr-transformer_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
