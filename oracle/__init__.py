"""CPU oracle for the ifcb_classifier hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline -- never on the shipped GPU path (which fails loudly when
the CUDA extension is missing).

Parity status: the reference (WHOIGit/ifcb_classifier v0.3.1) ships no tests,
golden vectors or fixtures, so this oracle is pinned
against outputs of the reference's own code executed in the build container:
``tests/golden/make_golden.py`` imports ``IfcbBinDataset`` unmodified from
``/root/reference/neuston_data.py`` (with stub ``ifcb`` modules) and records its
outputs; ``tests/test_oracle_golden.py`` replays them.  The ``.roi``/``.adc``
decoding step (pyifcb, absent and unpinned upstream) is "parity unpinned".
"""
