"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ionic-mpnn MPNN hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker (or as the CPU baseline that
is timed *beside* the CUDA path), never as the thing shipped.

Parity status: the reference (goalheart/ionic-mpnn) ships no tests, golden vectors,
weights or data for this path, and TensorFlow is not installable here, so the oracle is
pinned as follows (see DESIGN.md "Oracle"):

* ``oracle/tf_shim.py`` is a numpy stand-in for the handful of ``tf.*`` / Keras calls the
  reference makes.  ``tests/golden/make_golden.py`` uses it to execute the reference's OWN
  source (``models/layers.py`` unmodified, and the ``build_model`` /
  ``preprocess_edges_and_bonds`` / ``pad_sequences_1d`` functions lifted by AST from
  ``train_viscosity.py`` / ``train_melting_point.py``) in this container and records
  inputs, weights, every intermediate and the outputs under ``tests/golden/``.
* ``oracle/ref_model.py`` (torch, fp64/fp32) and ``oracle/ref_inputs.py`` are the
  standalone restatement that travels to the GPU box; ``tests/test_oracle_golden.py``
  holds them to those vectors.

That pins the reference's graph wiring, index conventions and quirks to
its source, but NOT TensorFlow's own kernels: with respect to a real TF run the oracle is
"parity unpinned".
"""
