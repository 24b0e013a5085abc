"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CraftingWorld hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import or execute it, and there only as
the checker or as the timed CPU baseline -- never on the product path (``gym_craftingworld_b200`` does not
import this package and fails loudly when its CUDA library is missing).

Contents
--------
``ref_shim``     loads the UNMODIFIED reference (``/root/reference``) under a throw-away ``gym`` /
                 ``matplotlib`` import shim.  Only usable in the builder container; used to generate the
                 frozen golden traces under ``tests/golden/`` and for live differential tests.
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
``tests/test_oracle_live_reference.py`` re-runs the differential live whenever ``/root/reference`` exists.
"""
