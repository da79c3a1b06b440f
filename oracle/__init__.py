"""CPU oracle for the ESRecsys hot path (TEST INFRASTRUCTURE, not product code).

A NumPy restatement of the arithmetic the reference's three trainers execute
through jax/flax/optax (absent from this image):

* ``oracle.glove``   -- wikipedia/models.py:21-55 + wikipedia/train_cooccurence.py:71-112
* ``oracle.spotify`` -- spotify/models.py:33-91 + spotify/train_spotify.py:77-150
* ``oracle.stl``     -- pinterest/models.py:63-74 + pinterest/train_shop_the_look.py:93-122
* ``oracle.optim``   -- optax adam / sgd(momentum) / adagrad semantics
* ``oracle.index``   -- our own integer bookkeeping contract (stable slot sort, segments,
                        cyclic owner routing); bit-exact target for the CUDA path

PINNING: the reference ships no tests, golden vectors or fixtures, and jax / flax /
optax cannot be installed here.  The oracle is pinned (to 1e-12, tests/test_ref_golden.py)
against vectors produced by EXECUTING the reference's own unmodified source files on
a torch-float64 stand-in for the jax / flax / optax surface they call
(tests/golden/make_ref_golden.py, tests/golden/refshim/): every statement of the
reference on the path is pinned; the third-party rules underneath (jax VJPs, optax
formulas) are restated in that stand-in, not executed -- in that sense parity with
real jax stays UNPINNED.  Also validated against an independent float64
torch-autograd transcription (tests/golden/make_golden.py) and hand-computed cases.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package. The product
(``esrecsys_b200``) never does.
"""
