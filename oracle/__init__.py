"""CPU oracle for the SENSE data-consistency hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package (``deep_cine_cardiac_mri_b200``); only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it, and there only as the checker / reported baseline.

Parity status: **unpinned by the reference's own tests** (the reference ships no
tests, fixtures or golden vectors — SURVEY.md §4).  The oracle is pinned instead
against outputs of the reference's own torch functions, generated in the
authoring container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed under ``tests/golden/``.
"""
from . import sense_oracle  # noqa: F401
