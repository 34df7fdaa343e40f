"""CPU oracle for the simulst streaming-alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``simulst_b200/`` imports this
package.  The only legal importers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and
there only as the checker / the reported CPU baseline, never as the product.

What it is: a restatement, in plain CPU PyTorch, of the algorithms in

* ``codebase/utils/functions.py``            (scan / window helpers)
* ``codebase/utils/monotonic_attention.py``  (MMA expected alignment, soft attention,
                                             mass preservation)
* ``codebase/utils/p_choose_strategy.py``    (p_choose strategies)
* ``codebase/modules/monotonic_multihead_attention.py:152-352`` (train / infer bodies)
* ``codebase/models/torch_cif/cif.py``       (CIF) and
  ``codebase/models/cif_transformer.py:143-261`` (CIFLayer forward / infer bodies)

of the reference (George0828Zhang/simulst @ 3f0e65e).  The reference is
Python/PyTorch, so the oracle is PyTorch too: the fp32 restatement executes the
same primitive sequence and is therefore bit-comparable with the reference on
CPU, and every function takes ``compute_dtype`` so the same formulas can be
evaluated in fp64 (the reference force-casts to fp32 internally and cannot).

Parity pinning:
* CIF  -- pinned by the reference's own property test (sequential checker
  ``torch_cif/test.py:24-91``, restated in ``oracle.cif.cif_sequential``) and by
  golden vectors generated from the unmodified reference.
* MMA  -- the reference ships NO test or golden vector for the alignment /
  soft-attention / incremental-step functions ("parity unpinned by the
  reference's tests").  It is pinned here by golden vectors produced by
  importing the unmodified reference files in the build container
  (``tests/golden/make_golden.py``) and by a live comparison whenever
  ``/root/reference`` is present (``oracle.ref_loader``).
"""
