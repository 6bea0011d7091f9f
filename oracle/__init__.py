"""CPU oracle for the DMHomo homography-warp hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the CPU arm being timed), never on the CUDA product path.

Contents
--------
``port.py``        torch-CPU / numpy restatement of the reference algorithm
                   (every function cites the reference ``file:line`` it follows).
``hsv.py``         restatement of ``matplotlib.colors.hsv_to_rgb`` (third-party,
                   absent from /root/reference, version unpinned by the reference:
                   **parity unpinned** at that one boundary).
``ref_loader.py``  imports the *real* reference in place from /root/reference
                   when it is mounted (this container only; never on the GPU
                   box) - used to pin ``port.py`` and to generate
                   ``tests/golden/*.npz``.

Pinning status: the reference ships no tests / golden vectors,
so ``port.py`` is pinned against outputs of the reference itself executed here
(``tests/golden/make_golden.py`` -> committed fixtures, re-checked by
``tests/test_oracle_golden.py``; and live, when /root/reference is mounted, by
``tests/test_oracle_vs_reference.py``).
"""
