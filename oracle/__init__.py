"""CPU oracle for the ESRecsys hot path (TEST INFRASTRUCTURE, not product code).

A NumPy restatement of the arithmetic the reference's three trainers execute
through jax/flax/optax (absent from this image):

* ``oracle.glove``   -- wikipedia/models.py:21-55 + wikipedia/train_cooccurence.py:71-112
* ``oracle.spotify`` -- spotify/models.py:33-91 + spotify/train_spotify.py:77-150
* ``oracle.stl``     -- pinterest/models.py:63-74 + pinterest/train_shop_the_look.py:93-122
* ``oracle.optim``   -- optax adam / sgd(momentum) / adagrad semantics
* ``oracle.index``   -- our own integer bookkeeping contract (stable slot sort, segments,
                        cyclic owner routing); bit-exact target for the CUDA path

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and
jax/flax/optax cannot be imported here, so the oracle cannot be checked against
the reference's own outputs. It is validated instead against an independent
float64 torch-autograd transcription of each forward (tests/test_oracle_*.py)
and hand-computed micro cases.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package. The product
(``esrecsys_b200``) never does.
"""
