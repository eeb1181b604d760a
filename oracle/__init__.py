"""CPU oracle for the stencil hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  The
product path (``geosongpu-ci_b200/b200stencil``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status ("Oracle"):

* S1 ``top_of_column``, S2 ``while_in_function``: pinned by the reference's own
  asserts (``dsl_patterns/Do__get_top_of_the_column.py:59-68``,
  ``dsl_patterns/Do__while_in_gt_functions.py:52-62``) -- see
  ``tests/test_oracle_golden.py``.
* S3 ``hybrid_index_2dout``: the reference demo prints only
  (``dsl_patterns/WIP__hybrid_index_2dout.py:68-90``); expected values follow by
  inspection of the stencil body -> **parity unpinned**.
* S4/S5/S6/halo exchange: no source in the reference, this oracle is the
  definition -> **parity unpinned**.

The arithmetic of the reference lives in gt4py.cartesian (numpy backend) reached
through NDSL 2024.04.00 (``sw_stack/discover/sles15/src/2024.04.00/basics.sh:19``);
neither is vendored in ``/root/reference`` nor installable here, so this is a
restatement of the published gt4py numpy-backend semantics.
"""
