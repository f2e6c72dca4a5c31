"""node2vec-by-ecc embedding hot path on B200: alias tables, second-order walks, SGNS.

Host side in Python mirroring the reference's interfaces (``walker`` == src/node2vec.py,
``word2vec`` == the gensim 3.2.0 surface used by learn_embeddings); all compute in
``libn2v_b200.so`` (hand-written sm_100a CUDA behind the C ABI of include/n2v_b200.h).
There is no CPU fallback: without the library or a CUDA device every compute call raises.
"""
from ._lib import N2VError, SO_PATH, lib  # noqa: F401
from .graph import AliasTables, DeviceGraph  # noqa: F401
from .walker import Graph, WalkCorpus, alias_draw, alias_setup  # noqa: F401
from .word2vec import BlockSgnsTrainer, KeyedVectors, LineSentence, PeerSgnsTrainer, SgnsTrainer, Vocab, Word2Vec  # noqa: F401

__all__ = ["Graph", "WalkCorpus", "alias_setup", "alias_draw", "DeviceGraph", "AliasTables",
           "Word2Vec", "KeyedVectors", "LineSentence", "Vocab", "N2VError", "lib", "SO_PATH"]
