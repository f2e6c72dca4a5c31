from node2vec_by_ecc_b200.word2vec import KeyedVectors, LineSentence, Vocab, Word2Vec  # noqa: F401
