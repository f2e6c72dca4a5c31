"""Minimal `gensim` stand-in exposing what the reference imports (gensim==3.2.0 is pinned in its
requirements.txt:17 but absent here): gensim.models.Word2Vec, gensim.models.KeyedVectors,
gensim.models.word2vec.LineSentence -- all backed by node2vec_by_ecc_b200 (GPU, no CPU fallback)."""
__version__ = "3.2.0+n2v_b200"
from . import models  # noqa: F401
