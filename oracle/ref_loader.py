"""Loader for the UNMODIFIED reference walker -- TEST INFRASTRUCTURE, build container only.

Imports /root/reference/src/node2vec.py as it lies (read-only) and offers the harness that
injects Philox uniforms into its ``np.random.rand()`` calls. /root/reference does not exist
on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this; it is
used by ``oracle/make_golden.py`` (which writes tests/golden/) and by not-gpu tests that
skip when the reference is absent.

Two shims, neither touching the file:
* ``numpy.int = int`` -- node2vec.py:248 uses the alias numpy removed in 1.24.
* ``module.sum = naive left-to-right sum`` (optional, default on) -- the reference's pinned
  stack is Python 2.7 whose ``sum()`` adds floats left to right; Python >= 3.12 compensates
  (Neumaier), which can move ``norm_const`` by one ulp for non-dyadic weights. With the shim
  the module computes what it computes under its own requirements.txt.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os

import numpy as np

REF_SRC = "/root/reference/src/node2vec.py"


def available() -> bool:
    return os.path.exists(REF_SRC)


def _naive_sum(xs, start=0):
    s = start
    for x in xs:
        s = s + x
    return s


def load(naive_sum: bool = True):
    """-> the reference's node2vec module object (fresh instance each call)."""
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001 - shim for node2vec.py:248
    spec = importlib.util.spec_from_file_location("_ref_node2vec", REF_SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if naive_sum:
        mod.sum = _naive_sum
    return mod


class PhiloxInjector:
    """Feeds alias_draw (node2vec.py:271-281) the uniforms of (walk_id, step): the walk's
    n-th pair of rand() calls gets step n (1-based == index of the token being drawn)."""

    def __init__(self, seed: int, walk_id_base: int = 0):
        from . import walk_uniforms
        self._u = walk_uniforms
        self.seed = seed
        self.walk_id = walk_id_base - 1
        self.calls = 0

    def begin_walk(self):
        self.walk_id += 1
        self.calls = 0

    def rand(self, *a):
        assert not a
        step = self.calls // 2 + 1
        which = self.calls % 2
        self.calls += 1
        return self._u(self.seed, self.walk_id, step)[which]


@contextlib.contextmanager
def injected(G, seed: int, walk_id_base: int = 0):
    """Patch numpy.random.rand and wrap G's per-walk methods so the reference's own
    simulate_walks / simulate_walks_on_the_fly loops (node2vec.py:81-111) run unmodified."""
    inj = PhiloxInjector(seed, walk_id_base)
    saved = np.random.rand
    o1, o2 = G.node2vec_walk, G.node2vec_walk_on_the_fly

    def w1(walk_length, start_node):
        inj.begin_walk()
        return o1(walk_length=walk_length, start_node=start_node)

    def w2(walk_length, start_node):
        inj.begin_walk()
        return o2(walk_length=walk_length, start_node=start_node)

    G.node2vec_walk, G.node2vec_walk_on_the_fly = w1, w2
    np.random.rand = inj.rand
    try:
        yield inj
    finally:
        np.random.rand = saved
        del G.node2vec_walk, G.node2vec_walk_on_the_fly
