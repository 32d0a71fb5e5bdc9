"""CPU oracle for the select -> env step -> TD update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU arm), never as a fallback for the CUDA path.

The oracle restates the reference's algorithm (``/root/reference``; citations
are ``file:line`` relative to that tree, abbreviations as in ``SURVEY.md``):

* ``oracle.rng``        -- the shared integer hash, the pre-drawn uniform stream
                           ``U[t, i, k]`` and the duck-typed RNG shims that map the
                           stream onto ``random.Random``/``numpy`` call sites.
* ``oracle.qlearning``  -- ``choose_actions`` / ``learn`` / ``learn_vec`` semantics of
                           ``OptimalQLearningBase`` (QLO:263-934).
* ``oracle.envs``       -- TicTacToe (TTT:67-237) + flatten wrapper (FLT:139-161,
                           UTL:12-48) + gymnasium ``SyncVectorEnv(SAME_STEP)``
                           restatement, the rigged bandit, the integer-hash MDP.
* ``oracle.runtime``    -- ``BaseRuntime.run_single_step`` / ``run_steps`` loop
                           (BRT:184-263, STR:28-76).
* ``oracle/c``          -- the same loop in plain C (fast enough for 2^20 agents).

Parity pinning: ``oracle/make_golden.py`` runs the *real* reference classes
(imported from ``/root/reference/src`` on top of ``oracle/gym_stub``) on seeded
inputs and pre-drawn uniforms and stores inputs + outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
them, and against the golden values held by the reference's own unit tests.
The gymnasium ``SyncVectorEnv(SAME_STEP)`` semantics are third-party and not
exercised by any reference test: parity is *unpinned* at that boundary (see
DESIGN.md).
"""
