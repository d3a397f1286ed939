"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Melissa rollout hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / reported CPU baseline.
``melissa_b200/`` never imports this package.

Contents
--------
ref_loader.py   imports the UNMODIFIED reference (``/root/reference``) under tiny stub
                modules for the packages that are not installed here (gymnasium,
                pettingzoo, matplotlib).  Exists only in the builder container; used to
                pin the restatements below and to generate ``tests/golden/*``.
env_oracle.py   numpy restatement of World/GraphEnv round semantics
                (reference graph_env/env/utils/core.py, graph_env/env/graph.py).
net_oracle.py   pure-torch fp32 restatement of DGN-R / L-DGN / HL-DGN incl. the
                third-party PyG / torch_cluster / tianshou pieces they call.
reset_chain.py  the reference's reset RNG chain (PCG64 -> RandomState) restated.

Parity status: env_oracle is PINNED (checked against the unmodified reference run in
this container and against the reference's own unit-test known answers, see
tests/test_oracle_env_vs_reference.py and tests/golden/).  net_oracle is "parity
unpinned" at the PyG / tianshou boundary: those packages are not installable offline,
the reference has no network test, so the restatement follows their published
semantics and hand-derived small cases only.
"""
