"""`import node2vec` drop-in: put node2vec_by_ecc_b200/dropin first on sys.path (PYTHONPATH) and
the reference's src/main.py / src/main_link.py pick this module up instead of src/node2vec.py."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from node2vec_by_ecc_b200.walker import Graph, WalkCorpus, alias_draw, alias_setup  # noqa: E402,F401
