"""CPU oracle for the keras-geometric message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``keras_geometric_b200/`` may import this
package; the only allowed users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (as the checker or the
reported CPU baseline, never as the thing shipped).

Contents
--------
keras_ops.py       restatement of the Keras-3 *torch backend* primitives the reference
                   reduces to (``take``, ``segment_sum``, ``segment_max`` ...).  Keras is a
                   third-party, un-vendored dependency of the reference (``keras>=3.0``,
                   /root/reference/pyproject.toml:33) and is not installed in this image,
                   so these are restated from Keras' published torch backend
                   (keras/src/backend/torch/{numpy,math}.py).
reference_path.py  line-by-line functional restatement of the reference's layer
                   algorithms, each function citing the reference file:line it follows.
kink.py            pins the oracle's GEMM outputs to given values (straight-through gradient) so that
                   LeakyReLU sign patterns coincide in full-size GATv2 gradient comparisons.
keras_shim/        a minimal stand-in ``keras`` package (built on keras_ops.py) that lets
                   the UNMODIFIED reference sources under /root/reference/src be imported
                   in the build container.  Used by tests/golden/make_golden.py to
                   generate the committed golden vectors; never shipped to users.

Parity pin: reference_path.py is checked (tests/test_oracle_golden.py) against
(1) the reference's own hand-computed KATs (tests/test_message_passing.py:54-155,
tests/test_graphsage_conv.py:431-537) and (2) golden vectors produced by running the
reference's own layer code through keras_shim (tests/golden/*.npz).  The Keras primitive
semantics themselves are restated from memory of Keras' source (no Keras available
offline), so that innermost layer is "parity unpinned" beyond the KATs.
"""
