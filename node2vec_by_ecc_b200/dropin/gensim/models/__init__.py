import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from node2vec_by_ecc_b200.word2vec import KeyedVectors, LineSentence, Word2Vec  # noqa: E402,F401
from . import word2vec  # noqa: E402,F401
