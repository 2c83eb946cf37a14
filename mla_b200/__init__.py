"""Importable alias of the product package, whose directory name
(`multimodal-learning-with-alternating-unimodal-adaptation_b200/`) is not a valid Python
identifier. `import mla_b200` executes that package's __init__ with this package's name, so
`mla_b200.ops`, `mla_b200.engine`, ... resolve to the files in that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "multimodal-learning-with-alternating-unimodal-adaptation_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
