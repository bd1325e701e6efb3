"""Importable alias of the ``shermbot-navigation_b200/`` package directory.

The product directory carries the repository's hyphenated name, which is not a valid Python
identifier; this stub makes ``import shermbot_navigation_b200`` resolve to it.
"""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "shermbot-navigation_b200"
__path__ = [str(_real)]
_init = _real / "__init__.py"
exec(compile(_init.read_text(), str(_init), "exec"))
