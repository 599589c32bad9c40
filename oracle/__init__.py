"""CPU oracle for the AlignNet-3D tp8 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported, linked or
executed by the product path (``alignnet-3d_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may use it, and there only as the checker / the timed CPU arm.

PARITY UNPINNED for the network path: the reference's arithmetic lives in
TensorFlow 1.8 (un-vendored; ``README.md:31``), which cannot be imported in this
environment, and the reference ships no golden vectors, tests or checkpoints
for ``models/tp8.py``.  The oracle is therefore a *restatement* written from
``models/tp8.py`` + ``utils/tf_util.py`` (two independent restatements, NumPy
fp32 and torch fp64/fp32, cross-checked against each other).  The only
known-answer vectors in the reference -- the ``euler2mat`` doctests,
``utils/eulerangles.py:152-159`` -- pin the Rz convention and are checked in
``tests/test_oracle_rigid.py``.
"""
