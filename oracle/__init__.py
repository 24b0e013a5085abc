"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CraftingWorld hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import or execute it, and there only as
the checker or as the timed CPU baseline -- never on the product path (``gym_craftingworld_b200`` does not
import this package and fails loudly when its CUDA library is missing).

Contents
--------
``ref_shim``     loads the UNMODIFIED reference (``/root/reference`` in the builder container, else its install
                 under ``oracle/_ref``) under a throw-away ``gym`` / ``matplotlib`` import shim; used to generate
                 the frozen golden traces under ``tests/golden/``, for live differential tests (CPU and GPU) and as
                 the reference arm of ``bench.py``.
``build_ref``    installs the unmodified reference package into ``oracle/_ref`` (pip, offline, ``--no-deps``;
                 git-ignored, travels to the GPU box with the built ``.so`` files).
``pyenv``        per-world env loops for the CPU-baseline legs: the reference's own class (``run_worker_reference``)
                 and the Python/NumPy port of round 1 (``run_worker``).
``compact``      NumPy/pure-Python restatement of the reference algorithm on the compact state encoding
                 (SURVEY.md Appendix A); every function cites the reference file:line it follows.
``cw_oracle.c``  plain-C restatement of the same algorithm (fast enough to check 4096 envs x hundreds of
                 steps incl. pixels in seconds, and multi-threaded as the CPU baseline); built by
                 ``oracle/build.py`` into ``oracle/libcw_oracle.so``.
``native``       ctypes binding of ``libcw_oracle.so``.

Parity pin: the reference ships no golden vectors or known-answer tests for this path
(``tests/craftingworld/test_core.py:1-14`` is stale and tests nothing on it).  The oracle is therefore
pinned against OUTPUTS OF THE REFERENCE ITSELF, run here: ``tests/golden/make_golden.py`` drives the
unmodified reference and freezes its per-step grids / positions / held item / achieved vector / reward /
done / pixels; ``tests/test_oracle_golden.py`` checks both restatements against those traces and
``tests/test_oracle_live_reference.py`` re-runs the differential live wherever the reference exists
(``/root/reference`` or ``oracle/_ref``), and ``tests/test_gpu_live_reference.py`` steps the CUDA kernels beside it.
"""
