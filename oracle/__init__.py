"""CPU oracle for the TextBoxGAN training-step hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch-CPU / NumPy, the arithmetic of the reference's
``training_step.py`` path (every function cites the reference file:line it follows).  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``textboxgan_b200/``
imports it, and the product path fails loudly when ``libtbg.so`` is missing.

Pinning status
--------------
* The reference (TensorFlow 2.8 + tensorflow_addons + easydict) cannot be imported in this image
  and ships no tests, golden vectors or fixtures (tests/test_unit.py:1-2 is a placeholder).
* StyleGAN2 / word-encoder / loss arithmetic: pinned against the reference's OWN source files
  executed here under a small TensorFlow-API shim (``oracle/tf_shim``; the generating script and
  the resulting fixtures live in ``tests/golden/``), plus hand-derivable known answers
  (tokeniser table, ``compute_paddings`` outputs, parameter counts).  See DESIGN.md §oracle.
* ASTER recogniser: weights and architecture are NOT in the reference repository
  (aster_ocr_utils/aster_inferer.py:24-26 loads an external SavedModel) — **parity unpinned**.
  ``oracle/aster.py`` is a seeded synthetic-weight restatement of the published ASTER topology.
"""
