"""CPU oracle for the padertorch separation hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference functions that
SURVEY.md section 8(a) lists (``padertorch/ops/_stft.py``,
``padertorch/ops/losses/source_separation.py``, ``padertorch/ops/losses/regression.py``
and the ``review`` loops of the three hot-path models).  Every function cites the
reference ``file:line`` it follows.

Rules (see the task statement, item 3):

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import this package, and only as the checker or as the
  timed CPU baseline -- never as the thing shipped.  ``padertorch_b200`` (the product)
  must not import it; ``tests/test_capi_cpu.py::test_product_never_imports_oracle`` enforces that.
* parity status: **pinned**.  ``tests/test_oracle_golden.py`` checks this package against
  (a) the known-answer vectors the reference's own tests / doctests hold for the path and
  (b) fixtures under ``tests/golden/`` that ``oracle/make_golden.py`` recorded by running
  the UNMODIFIED reference (imported from ``/root/reference`` with the stand-ins in
  ``oracle/ref_standins``) in the build container.
  One function is *parity unpinned*: ``sample_index_to_frame_index`` (no reference test,
  doctest or call site exercises it).

Third-party arithmetic that is not under ``/root/reference``: ``paderbox`` (unpinned
``install_requires`` of the reference, ``setup.py:135``) supplies the analysis window,
the biorthogonal synthesis window and the frame arithmetic.  Its published behaviour is
restated in ``oracle/stft.py`` and anchored on the reference's call sites and known
answers (literal STFT matrix ``padertorch/contrib/cb/transform.py:219-232``, frame counts
``tests/test_ops/test_stft.py:44-70,139-165``, STFT->iSTFT round trip ``:36-42``).
"""
from . import stft, losses, path  # noqa: F401
