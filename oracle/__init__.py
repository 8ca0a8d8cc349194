"""CPU oracle for the nekStab Arnoldi / Newton-Krylov hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nekstab_next_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or as the
CPU arm that is timed beside the GPU path.

Parity status
-------------
* In-tree Fortran (``core/krylov_decomposition.f90``, ``krylov_subspace.f90``,
  ``nek_vectors.f90``, ``eigensolvers.f90``, ``lapack_wrapper.f90``,
  ``newton_krylov.f90``, ``utils.f90:mth_rand``): restated literally (same loop
  order, same two-pass MGS, same LAPACK routines through scipy).  The reference
  cannot be compiled here (no Fortran compiler, Nek5000/LightKrylov not
  vendored) and ships no golden vectors for H / Ritz values:
  **parity unpinned** for those outputs.
* SEM kernels (``ax``/``axhelm``, ``local_grad3``, geometry, ``dssum``,
  ``glsc3``) live in un-vendored Nek5000 (``master``, unpinned, see
  ``Nek5000_setup.sh:71-73``): restated from the published algorithm.
  Pinned by the reference's own field-file fixtures through derived known
  answers (sum(bm1) = domain area, <U,U>_bm1, unique-node counts; SURVEY.md
  section 8c) -- see ``tests/test_oracle_fixtures.py``.
"""
