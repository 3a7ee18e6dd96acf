"""CPU oracle for the HD-GNN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or as the timed CPU baseline.
The product package (``hdgnn_b200``) never imports this package.

Parity status (see DESIGN.md section "Oracle"):

* loader / pair enumeration / pooling indices (utils2.py) and the evaluation
  functions (EvaluationFuncs.py): PINNED -- golden vectors in ``tests/golden`` were
  produced by executing the reference's own files from /root/reference
  (``oracle/gen_golden.py``).
* network forward / loss / Adam (model_1..4.py, model.py): the TensorFlow runtime
  is not installable here, so the reference graph cannot be executed by TF itself
  -- "parity unpinned" at the TF-kernel boundary.  What pins it instead: the
  reference's UNMODIFIED model files are executed over a torch-backed shim of the
  handful of ``tf.*`` calls they make (``oracle/tf1_shim.py``), the result is
  committed as golden vectors, and both restatements in
  ``oracle/hdgnn_oracle.py`` (dense one-hot transcription, closed index form) are
  checked against those vectors.
"""
