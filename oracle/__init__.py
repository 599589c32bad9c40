"""CPU oracle for the AlignNet-3D tp8 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported, linked or
executed by the product path (``alignnet-3d_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may use it, and there only as the checker / the timed CPU arm.

PARITY STATUS.  The reference's arithmetic lives in TensorFlow 1.8 (third-party,
un-vendored; ``README.md:31``), which cannot be installed in this environment, and
the reference ships no golden vectors, tests or checkpoints for ``models/tp8.py``.
The oracle is pinned one level below that:

* ``oracle/tf1_shim`` is an eager stand-in for the ~60 TensorFlow symbols the path
  uses, and ``oracle/reference_run.py`` imports the reference's OWN, UNMODIFIED
  ``models/tp8.py`` / ``utils/tf_util.py`` / ``config.py`` / ``tp_utils/pointcloud.py``
  / ``utils/eulerangles.py`` from /root/reference and executes them on it.  The
  fixtures ``tests/golden/reference_*.npz`` are outputs of that run (generator:
  ``tests/golden/make_reference_golden.py``).  So graph structure, variable
  naming/sharing, op order, the decode, every loss term and quirk come from the
  reference's code, not from a reading of it; only the TF primitives (conv2d,
  moments, batch_normalization, max_pool, floormod, argmax, EMA, ...) are restated,
  each from its published TF 1.8 definition (shim header).
* the two restatements here (``np_forward.py`` NumPy fp32, ``torch_ref.py`` torch
  fp64/fp32 autograd) reproduce those fixtures: fp64 outputs, loss, EMA shadows and
  all 92 gradients to <= 1e-9 (``tests/test_reference_run.py``).
* the only known-answer vectors in the reference -- the ``euler2mat`` doctests,
  ``utils/eulerangles.py:152-159`` -- pin the Rz convention (``tests/test_oracle.py``).

What remains unpinned is TensorFlow's own kernels (their fp32 summation order), i.e.
agreement with a TF 1.8 binary at the last few ulps.
"""
